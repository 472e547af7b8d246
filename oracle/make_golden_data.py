"""Generate tests/golden/ref_data.npz: the reference's own per-sample preprocessing (data/base_dataset.py:50-80 get_transform
in 'fixed' mode, as data/openeds_dataset.py:82-119 applies it: cv2 nearest for the mask, PIL bicubic + ToTensor + Normalize for
the images, optional horizontal flip) on synthetic uint8 OpenEDS-shaped frames.  Runs only in the build container.

    python oracle/make_golden_data.py

Per case the fixture stores SHA-256 digests of the full results (bit-exact checks) and strided subsamples."""
import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import import_reference  # noqa: E402

CASES = {"R1": (256, 0.8, False), "R1_flip": (256, 0.8, True), "R2_flip": (384, 0.6, True), "wide": (512, 0.8, False)}


def data_inputs(name):
    rng = np.random.Generator(np.random.PCG64(300 + sorted(CASES).index(name)))
    yy, xx = np.mgrid[0:640, 0:400]
    mask = ((((yy - 320) / 200.0) ** 2 + ((xx - 200) / 150.0) ** 2 <= 1).astype(np.uint8)
            + (((yy - 320) ** 2 + (xx - 210) ** 2) <= 90 ** 2) + (((yy - 320) ** 2 + (xx - 210) ** 2) <= 40 ** 2)).astype(np.uint8)
    mask[rng.integers(0, 640, 200), rng.integers(0, 400, 200)] = rng.integers(0, 4, 200).astype(np.uint8)   # speckle: exercises indexing
    images = rng.integers(0, 256, size=(3, 640, 400)).astype(np.uint8)
    images[0] = np.clip(128 + 100 * np.sin(yy / 7.0) * np.cos(xx / 5.0) + rng.normal(0, 20, (640, 400)), 0, 255).astype(np.uint8)
    return mask, images


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def main():
    import_reference()
    import cv2
    from PIL import Image
    from data.base_dataset import get_transform
    out = {}
    for name, (crop, ar, flip) in CASES.items():
        opt = SimpleNamespace(preprocess_mode="fixed", crop_size=crop, aspect_ratio=ar, load_size=crop, isTrain=True, no_flip=False)
        params = {"crop_pos": (0, 0), "flip": flip}
        mask, images = data_inputs(name)
        t_mask = get_transform(opt, params, method=cv2.INTER_NEAREST, normalize=False, toTensor=False)
        t_img = get_transform(opt, params)
        lab = torch.from_numpy(np.ascontiguousarray(t_mask(mask)))
        ims = torch.stack([t_img(Image.fromarray(im, mode="L")) for im in images])
        out[name + "|label_sha"] = digest(lab.numpy().astype(np.uint8))
        out[name + "|label_sub"] = lab.numpy().reshape(-1)[::397].astype(np.uint8)
        out[name + "|images_sha"] = digest(ims.numpy().astype(np.float32))
        out[name + "|images_sub"] = ims.numpy().reshape(-1)[::997].astype(np.float32)
        out[name + "|shape"] = np.array(ims.shape)
        print(name, tuple(lab.shape), tuple(ims.shape), float(ims.mean()))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "ref_data.npz"), **out)


if __name__ == "__main__":
    main()
