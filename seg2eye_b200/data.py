"""Device-side data layer (beyond the reference API).

The reference prepares every sample on the host, one at a time, in the (single-threaded, `nThreads=0`) loader:
`data/openeds_dataset.py:82-119` resizes the uint8 640x400 mask with cv2 (nearest) and each style / target image with PIL
(bicubic), flips, converts to float and normalises (`data/base_dataset.py:50-80`, preprocess_mode 'fixed').  At hundreds of
images per second per GPU that loop is the limiter, so here the RAW uint8 frames are copied to the GPU and the same
arithmetic runs in four small kernels -- bit-exact: integer label maps and Pillow's fixed-point bicubic are reproduced
exactly (tests/test_gpu_data.py checks every bit against the reference's own get_transform output).

    prep = DevicePreprocessor(opt)                       # crop_size / aspect_ratio / no_flip / isTrain as in the reference
    data_i = prep({'label': u8 (B,640,400), 'style_image': u8 (B,ns,640,400), 'target': u8 (B,640,400)})   # host or CUDA
    trainer.run_generator_one_step(data_i)               # label (B,1,h,w) int64, style_image (B,ns,1,h,w), target (B,1,h,w),
                                                         # target_original (B,1,640,400) int32 -- the reference's dict
"""
import math
import random

import numpy as np
import torch

from . import _lib as L

_PRECISION_BITS = 32 - 8 - 2      # Pillow, src/libImaging/Resample.c


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_table(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for BICUBIC (host side, double precision, same operation order):
    taps int32 [out_size][ksize], bounds int32 [out_size][2] = (first source index, tap count)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * inv) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(n):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << _PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return kk, bounds


class DevicePreprocessor:
    """OpenEDSDataset.__getitem__ + get_transform ('fixed' mode) for a whole batch of raw frames, on the device."""

    def __init__(self, opt, device=None):
        if getattr(opt, 'preprocess_mode', 'fixed') != 'fixed':
            raise ValueError("DevicePreprocessor implements --preprocess_mode fixed (the reference's default)")
        self.w = opt.crop_size
        self.h = round(opt.crop_size / opt.aspect_ratio)
        self.flip_enabled = bool(getattr(opt, 'isTrain', True)) and not getattr(opt, 'no_flip', False)
        self.device = torch.device(device if device is not None else 'cuda')
        self._tables = {}

    def _table(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._tables:
            kk, bounds = pil_bicubic_table(in_size, out_size)
            self._tables[key] = (torch.from_numpy(kk).to(self.device), torch.from_numpy(bounds).to(self.device), kk.shape[1])
        return self._tables[key]

    def resize_images(self, frames):
        """(N,H0,W0) uint8 on the device -> (N,h,w) uint8: Image.resize((w,h), BICUBIC), horizontal pass then vertical."""
        N, H0, W0 = frames.shape
        cur, st = frames, L.stream()
        if W0 != self.w:
            kk, bounds, ks = self._table(W0, self.w)
            out = torch.empty(N, H0, self.w, dtype=torch.uint8, device=self.device)
            L.call("s2e_pil_resample_u8", L.ptr(cur), N, H0, W0, self.w, 1, L.ptr(kk), L.ptr(bounds), ks, L.ptr(out), st)
            cur = out
        if H0 != self.h:
            kk, bounds, ks = self._table(H0, self.h)
            out = torch.empty(N, self.h, self.w, dtype=torch.uint8, device=self.device)
            L.call("s2e_pil_resample_u8", L.ptr(cur), N, H0, self.w, self.h, 0, L.ptr(kk), L.ptr(bounds), ks, L.ptr(out), st)
            cur = out
        return cur

    def __call__(self, raw, flip=None):
        """raw: 'label' uint8 (B,H0,W0); optional 'style_image' uint8 (B,ns,H0,W0), 'target' uint8 (B,H0,W0); other keys pass
        through.  flip: per-sample booleans (default: drawn like base_dataset.get_params, random.random() > 0.5)."""
        dev, st = self.device, L.stream()
        mask = raw['label'].to(dev, non_blocking=True).contiguous()
        assert mask.dtype == torch.uint8 and mask.dim() == 3
        B, H0, W0 = mask.shape
        if flip is None:
            flip = [self.flip_enabled and random.random() > 0.5 for _ in range(B)]
        flags = torch.tensor([1 if f else 0 for f in flip], dtype=torch.uint8).to(dev, non_blocking=True)
        out = {k: v for k, v in raw.items() if k not in ('label', 'style_image', 'target')}
        label = torch.empty(B, 1, self.h, self.w, dtype=torch.int64, device=dev)
        L.call("s2e_label_nearest_flip", L.ptr(mask), B, H0, W0, self.h, self.w, L.ptr(flags), L.ptr(label), st)
        out['label'] = label
        if 'style_image' in raw:
            sty = raw['style_image'].to(dev, non_blocking=True).contiguous()
            ns = sty.shape[1]
            small = self.resize_images(sty.view(B * ns, H0, W0))
            t = torch.empty(B, ns, 1, self.h, self.w, dtype=torch.float32, device=dev)
            L.call("s2e_u8_flip_normalize", L.ptr(small), B * ns, ns, self.h, self.w, L.ptr(flags), L.ptr(t), st)
            out['style_image'] = t
        if 'target' in raw:
            tgt = raw['target'].to(dev, non_blocking=True).contiguous()
            small = self.resize_images(tgt)
            t = torch.empty(B, 1, self.h, self.w, dtype=torch.float32, device=dev)
            L.call("s2e_u8_flip_normalize", L.ptr(small), B, 1, self.h, self.w, L.ptr(flags), L.ptr(t), st)
            out['target'] = t
            orig = torch.empty(B, 1, H0, W0, dtype=torch.int32, device=dev)
            L.call("s2e_u8_flip_to_i32", L.ptr(tgt), B, H0, W0, L.ptr(flags), L.ptr(orig), st)
            out['target_original'] = orig
        return out
