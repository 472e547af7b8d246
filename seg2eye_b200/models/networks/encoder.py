"""ConvEncoder (style encoder) mirror (reference models/networks/encoder.py:13-73)."""
import numpy as np
import torch.nn as nn

from ... import _lib as L
from ... import ops
from .base_network import BaseNetwork
from .layers import Conv2d, Linear
from .normalization import get_nonspade_norm_layer


class ConvEncoder(BaseNetwork):
    def __init__(self, opt):
        super().__init__()
        kw = 3
        pw = int(np.ceil((kw - 1.0) / 2))
        ndf = opt.ngf
        norm_layer = get_nonspade_norm_layer(opt, opt.norm_E)
        chans = [1, ndf, ndf * 2, ndf * 4, ndf * 8, ndf * 8]
        if opt.crop_size >= 256:
            chans.append(ndf * 8)
        self.len_sequence = len(chans) - 1
        for n in range(self.len_sequence):
            self.add_module('layer' + str(n), norm_layer(Conv2d(chans[n], chans[n + 1], kw, stride=2, padding=pw)))
        self.so = s0 = 4
        self.fc_mu = Linear(ndf * 8 * s0 * s0, opt.w_dim)
        self.fc_var = Linear(ndf * 8 * s0 * s0, opt.w_dim)
        self.actvn = nn.LeakyReLU(0.2, False)
        self.opt = opt

    def forward_samples(self, x5):
        """Equivalent of `[self(x5[b]) for b in range(B)]` (pix2pix_model.py:285) in one batched pass.

        x5: (B, ns, 1, H, W).  Every spectral-normed layer still advances its (u, v) B times and sample b is
        normalised with its own sigma_b: the convolutions run unscaled over all B*ns images, 1/sigma_b enters the
        InstanceNorm statistics of sample b, and the sigma chain-rule term is added in backward.
        Returns mu, logvar of shape (B, ns, w_dim) and the 6 feature maps, each (B*ns, C, h, w)."""
        B, ns = x5.shape[0], x5.shape[1]
        x = x5.reshape(B * ns, *x5.shape[2:])
        if x.size(2) != 256 or x.size(3) != 256:
            h = ops.BilinearFn.apply(x, (256, 256))
        else:
            h = ops.as_nhwc(x)
        features = []
        ops.prepare_spectral([getattr(self, 'layer' + str(n))[0] for n in range(self.len_sequence)], self.training, B, True)
        for n in range(self.len_sequence):
            layer = getattr(self, 'layer' + str(n))
            h = layer[1].forward_nhwc_spectral(layer[0].forward_nhwc_unscaled(h), layer[0], B)
            features.append(ops.as_nchw_view(h))
        hw = h.shape[1] * h.shape[2]
        mu = ops.LinearFn.apply(h, self.fc_mu.weight, self.fc_mu.bias, L.ACT_NONE, hw)
        logvar = ops.LinearFn.apply(h, self.fc_var.weight, self.fc_var.bias, L.ACT_NONE, hw)
        return mu.view(B, ns, -1), logvar.view(B, ns, -1), features

    def forward(self, x, get_intermediate_features=False):
        # (ns,1,H,W) fp32 -> bilinear 256x256 (encoder.py:54-55) -> NHWC bf16
        if x.size(2) != 256 or x.size(3) != 256:
            h = ops.BilinearFn.apply(x, (256, 256))
        else:
            h = ops.as_nhwc(x)
        features = []
        for n in range(self.len_sequence):
            layer = getattr(self, 'layer' + str(n))
            h = layer[1].forward_nhwc(layer[0].forward_nhwc(h))
            features.append(ops.as_nchw_view(h))
        hw = h.shape[1] * h.shape[2]
        # LeakyReLU(0.2) + NCHW flatten are folded into the head kernel
        mu = ops.LinearFn.apply(h, self.fc_mu.weight, self.fc_mu.bias, L.ACT_NONE, hw)
        logvar = ops.LinearFn.apply(h, self.fc_var.weight, self.fc_var.bias, L.ACT_NONE, hw)
        return mu, logvar, features
