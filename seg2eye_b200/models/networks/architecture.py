"""SPADE_STYLE_ResnetBlock mirror (reference models/networks/architecture.py:13-62)."""
import torch.nn as nn

from ... import _lib as L
from ... import ops
from .layers import Conv2d
from .normalization import SPADE_STYLE_Block


class SPADE_STYLE_ResnetBlock(nn.Module):
    def __init__(self, fin, fout, opt):
        super().__init__()
        self.learned_shortcut = (fin != fout)
        fmiddle = min(fin, fout)
        sn = 'spectral' in opt.norm_G
        self.conv_0 = Conv2d(fin, fmiddle, 3, padding=1, spectral=sn)
        self.conv_1 = Conv2d(fmiddle, fout, 3, padding=1, spectral=sn)
        if self.learned_shortcut:
            self.conv_s = Conv2d(fin, fout, 1, bias=False, spectral=sn)
        self.norm_0 = SPADE_STYLE_Block(fin, opt)
        self.norm_1 = SPADE_STYLE_Block(fmiddle, opt)
        if self.learned_shortcut:
            self.norm_s = SPADE_STYLE_Block(fin, opt)

    def forward_nhwc(self, x, seg, latent_style):
        # same evaluation order as the reference (shortcut first) so BN buffers / SN vectors advance identically;
        # the LeakyReLU(0.2) of actvn() is fused into the modulation kernel
        if self.learned_shortcut:
            x_s = self.conv_s.forward_nhwc(self.norm_s.forward_nhwc(x, seg, latent_style, L.ACT_NONE))
        else:
            x_s = x
        dx = self.conv_0.forward_nhwc(self.norm_0.forward_nhwc(x, seg, latent_style, L.ACT_LRELU))
        dx = self.conv_1.forward_nhwc(self.norm_1.forward_nhwc(dx, seg, latent_style, L.ACT_LRELU))
        return ops.AddFn.apply(x_s, dx)

    def forward(self, x, seg, latent_style):
        return ops.as_nchw_view(self.forward_nhwc(ops.as_nhwc(x), seg, latent_style))

    def shortcut(self, x, seg, latent_style):
        if self.learned_shortcut:
            xn = ops.as_nhwc(x)
            return ops.as_nchw_view(self.conv_s.forward_nhwc(self.norm_s.forward_nhwc(xn, seg, latent_style, L.ACT_NONE)))
        return x

    def actvn(self, x):
        return ops.as_nchw_view(ops.ActFn.apply(ops.as_nhwc(x), L.ACT_LRELU))
