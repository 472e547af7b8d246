"""SPADE+Style ResNet block with the reference's interface (reference models/networks/architecture.py:13-62):

    out = shortcut(x) + conv_1(lrelu(norm_1(conv_0(lrelu(norm_0(x))))))      shortcut = conv_s(norm_s(x)) if fin != fout

The LeakyReLU(0.2) is fused into the normalisation kernel, the convolutions are spectral-normed tap convolutions."""
import torch.nn as nn

from ... import _lib as L
from ... import ops
from .layers import Conv2d
from .normalization import SPADE_Block, SPADE_STYLE_Block


class SPADE_STYLE_ResnetBlock(nn.Module):
    norm_block = SPADE_STYLE_Block

    def __init__(self, fin, fout, opt):
        super().__init__()
        mid = min(fin, fout)
        sn = 'spectral' in opt.norm_G
        self.learned_shortcut = fin != fout
        # registration order fixes the state_dict order: conv_0, conv_1, [conv_s], norm_0, norm_1, [norm_s]
        self.conv_0 = Conv2d(fin, mid, 3, padding=1, spectral=sn)
        self.conv_1 = Conv2d(mid, fout, 3, padding=1, spectral=sn)
        if self.learned_shortcut:
            self.conv_s = Conv2d(fin, fout, 1, bias=False, spectral=sn)
        self.norm_0 = self.norm_block(fin, opt)
        self.norm_1 = self.norm_block(mid, opt)
        if self.learned_shortcut:
            self.norm_s = self.norm_block(fin, opt)

    def _shortcut_nhwc(self, x, seg, w, up=False, sink=None):
        if not self.learned_shortcut:
            return x
        return self.conv_s.forward_nhwc(self.norm_s.forward_nhwc(x, seg, w, L.ACT_NONE, up, sink))

    def forward_nhwc(self, x, seg, latent_style, up=False):
        """up: the block input is the nearest-2x up-sampling of x (generator.py:86-93).  With a learned shortcut only
        norm_0 / norm_s read the input, so the up-sampled tensor is never written: both read x through an index map.
        Otherwise (identity shortcut: the input itself is added to the output) it is materialised first."""
        if up and not self.learned_shortcut:
            x, up = ops.Upsample2xFn.apply(x), False
        # norm_s and norm_0 both read x: their backward passes write ONE gradient buffer (ops.GradSink)
        sink = ops.GradSink() if (self.learned_shortcut and x.requires_grad) else None
        # the shortcut runs first, as in the reference, so BN buffers / spectral vectors advance in the same order
        skip = self._shortcut_nhwc(x, seg, latent_style, up, sink)
        h = self.conv_0.forward_nhwc(self.norm_0.forward_nhwc(x, seg, latent_style, L.ACT_LRELU, up, sink))
        # out = x_s + dx: the residual add rides in conv_1's epilogue
        return self.conv_1.forward_nhwc(self.norm_1.forward_nhwc(h, seg, latent_style, L.ACT_LRELU), residual=skip)

    def forward(self, x, seg, latent_style):
        return ops.as_nchw_view(self.forward_nhwc(ops.as_nhwc(x), seg, latent_style))

    def shortcut(self, x, seg, latent_style):
        return ops.as_nchw_view(self._shortcut_nhwc(ops.as_nhwc(x), seg, latent_style)) if self.learned_shortcut else x

    def actvn(self, x):
        return ops.as_nchw_view(ops.ActFn.apply(ops.as_nhwc(x), L.ACT_LRELU))


class SPADEResnetBlock(SPADE_STYLE_ResnetBlock):
    """architecture.py:13-62 with plain SPADE normalisation (no style branch): the ResNet block of the original SPADE
    generator (BASELINE config 5).  forward(x, seg) -- a third argument is accepted and ignored."""
    norm_block = SPADE_Block

    def forward_nhwc(self, x, seg, latent_style=None, up=False):
        return super().forward_nhwc(x, seg, None, up)

    def forward(self, x, seg, latent_style=None):
        return ops.as_nchw_view(self.forward_nhwc(ops.as_nhwc(x), seg))
