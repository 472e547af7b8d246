"""Multiscale PatchGAN discriminator with the reference's interface and state_dict layout
(reference models/networks/discriminator.py:14-116): `num_D` copies of a 4x4-conv stack, each fed a 3x3/s2 average-pooled
version of the previous input; every sub-block's output is returned for the feature-matching loss."""
import math

import torch.nn as nn

from ... import _lib as L
from ... import ops
from .base_network import BaseNetwork
from .layers import Conv2d
from .normalization import get_nonspade_norm_layer


class NLayerDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--n_layers_D', type=int, default=4, help='number of conv layers per discriminator')
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        k, pad = 4, math.ceil((4 - 1.0) / 2)
        with_norm = get_nonspade_norm_layer(opt, opt.norm_D)
        widths = [opt.ndf]
        for _ in range(1, opt.n_layers_D):
            widths.append(min(widths[-1] * 2, 512))
        # model0: conv + LeakyReLU (fused epilogue); model1..n-1: spectral conv -> InstanceNorm(+LeakyReLU, fused);
        # model<n>: 1-channel logit map.  nn.Sequential containers only carry the reference's parameter names.
        blocks = [[Conv2d(self.compute_D_input_nc(opt), widths[0], k, stride=2, padding=pad, act=L.ACT_LRELU), nn.LeakyReLU(0.2, False)]]
        for n in range(1, opt.n_layers_D):
            stride = 2 if n < opt.n_layers_D - 1 else 1
            blk = with_norm(Conv2d(widths[n - 1], widths[n], k, stride=stride, padding=pad))
            if isinstance(blk, nn.Sequential):
                blk[1].act = L.ACT_LRELU
            blocks.append([blk, nn.LeakyReLU(0.2, False)])
        blocks.append([Conv2d(widths[-1], 1, k, stride=1, padding=pad)])
        for n, mods in enumerate(blocks):
            self.add_module('model%d' % n, nn.Sequential(*mods))

    def compute_D_input_nc(self, opt):
        return opt.label_nc + opt.output_nc

    def forward_nhwc(self, x):
        outs = []
        for _, sub in self.named_children():
            head = sub[0]
            if isinstance(head, nn.Sequential):          # spectral conv -> InstanceNorm + LeakyReLU
                x = head[1].forward_nhwc(head[0].forward_nhwc(x))
            elif len(sub) > 1:                            # conv whose LeakyReLU lives in the conv epilogue
                x = head.forward_nhwc(x, act=L.ACT_LRELU)
            else:                                         # final prediction
                x = head.forward_nhwc(x)
            outs.append(x)
        return outs

    def forward(self, input):
        outs = [ops.as_nchw_view(t) for t in self.forward_nhwc(ops.as_nhwc(input))]
        return outs[-1] if self.opt.no_ganFeat_loss else outs


class MultiscaleDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--netD_subarch', type=str, default='n_layer', help='architecture of each discriminator')
        parser.add_argument('--num_D', type=int, default=2, help='number of discriminators (scales)')
        opt, _ = parser.parse_known_args()
        if opt.netD_subarch != 'n_layer':
            raise ValueError('unrecognized discriminator subarchitecture %s' % opt.netD_subarch)
        return NLayerDiscriminator.modify_commandline_options(parser, is_train)

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, self.create_single_discriminator(opt))

    def create_single_discriminator(self, opt):
        if opt.netD_subarch != 'n_layer':
            raise ValueError('unrecognized discriminator subarchitecture %s' % opt.netD_subarch)
        return NLayerDiscriminator(opt)

    def downsample(self, input):
        """F.avg_pool2d(3, stride 2, padding 1, count_include_pad=False)."""
        return ops.as_nchw_view(ops.AvgPool3s2Fn.apply(ops.as_nhwc(input)))

    def forward_nhwc(self, x):
        keep_all = not self.opt.no_ganFeat_loss
        ops.prepare_spectral([m for m in self.modules() if isinstance(m, Conv2d)], self.training)
        per_scale = []
        scales = list(self.children())
        for i, d in enumerate(scales):
            outs = d.forward_nhwc(x)
            per_scale.append([ops.as_nchw_view(t) for t in (outs if keep_all else outs[-1:])])
            if i + 1 < len(scales):      # the reference also pools after the last scale and drops the result
                x = ops.AvgPool3s2Fn.apply(x)
        return per_scale

    def forward(self, input):
        return self.forward_nhwc(ops.as_nhwc(input))
