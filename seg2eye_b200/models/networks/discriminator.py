"""MultiscaleDiscriminator / NLayerDiscriminator mirrors (reference models/networks/discriminator.py:14-116)."""
import numpy as np
import torch.nn as nn

from ... import _lib as L
from ... import ops
from .base_network import BaseNetwork
from .layers import Conv2d
from .normalization import get_nonspade_norm_layer


class MultiscaleDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--netD_subarch', type=str, default='n_layer', help='architecture of each discriminator')
        parser.add_argument('--num_D', type=int, default=2, help='number of discriminators to be used in multiscale')
        opt, _ = parser.parse_known_args()
        if opt.netD_subarch != 'n_layer':
            raise ValueError('unrecognized discriminator subarchitecture %s' % opt.netD_subarch)
        NLayerDiscriminator.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        for i in range(opt.num_D):
            self.add_module('discriminator_%d' % i, self.create_single_discriminator(opt))

    def create_single_discriminator(self, opt):
        if opt.netD_subarch == 'n_layer':
            return NLayerDiscriminator(opt)
        raise ValueError('unrecognized discriminator subarchitecture %s' % opt.netD_subarch)

    def downsample(self, input):
        return ops.as_nchw_view(ops.AvgPool3s2Fn.apply(ops.as_nhwc(input)))

    def forward_nhwc(self, x):
        result = []
        get_intermediate_features = not self.opt.no_ganFeat_loss
        for name, D in self.named_children():
            out = D.forward_nhwc(x)
            result.append(out if get_intermediate_features else [out[-1]])
            x = ops.AvgPool3s2Fn.apply(x)
        return [[ops.as_nchw_view(t) for t in d] for d in result]

    def forward(self, input):
        return self.forward_nhwc(ops.as_nhwc(input))


class NLayerDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--n_layers_D', type=int, default=4, help='# layers in each discriminator')
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        kw = 4
        padw = int(np.ceil((kw - 1.0) / 2))
        nf = opt.ndf
        input_nc = self.compute_D_input_nc(opt)
        norm_layer = get_nonspade_norm_layer(opt, opt.norm_D)
        sequence = [[Conv2d(input_nc, nf, kw, stride=2, padding=padw, act=L.ACT_LRELU), nn.LeakyReLU(0.2, False)]]
        for n in range(1, opt.n_layers_D):
            nf_prev = nf
            nf = min(nf * 2, 512)
            stride = 1 if n == opt.n_layers_D - 1 else 2
            blk = norm_layer(Conv2d(nf_prev, nf, kw, stride=stride, padding=padw))
            if isinstance(blk, nn.Sequential):
                blk[1].act = L.ACT_LRELU  # the LeakyReLU that follows is fused into the InstanceNorm kernel
            sequence += [[blk, nn.LeakyReLU(0.2, False)]]
        sequence += [[Conv2d(nf, 1, kw, stride=1, padding=padw)]]
        for n in range(len(sequence)):
            self.add_module('model' + str(n), nn.Sequential(*sequence[n]))

    def compute_D_input_nc(self, opt):
        return opt.label_nc + opt.output_nc

    def forward_nhwc(self, x):
        results = []
        for name, sub in self.named_children():
            first = sub[0]
            if isinstance(first, nn.Sequential):       # spectral conv -> InstanceNorm(+LeakyReLU)
                x = first[1].forward_nhwc(first[0].forward_nhwc(x))
            elif len(sub) > 1:                          # conv + LeakyReLU (fused epilogue) or SN conv w/o norm
                x = first.forward_nhwc(x, act=L.ACT_LRELU)
            else:                                       # final 1-channel prediction
                x = first.forward_nhwc(x)
            results.append(x)
        return results

    def forward(self, input):
        outs = [ops.as_nchw_view(t) for t in self.forward_nhwc(ops.as_nhwc(input))]
        return outs if not self.opt.no_ganFeat_loss else outs[-1]
