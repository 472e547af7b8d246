"""SPADE+Style generator with the reference's interface and state_dict layout.

Mirrors `SPADESTYLEGenerator` (reference models/networks/generator.py:13-102): a 3x3 conv on the down-sampled
segmap, seven SPADE+Style ResNet blocks separated by nearest 2x up-sampling, LeakyReLU, 3x3 conv to one channel, tanh.
Here the whole forward runs on NHWC bf16 activations through seg2eye_b200.ops; only the result is (B,1,H,W) fp32.
"""
from ... import _lib as L
from ... import ops
from .architecture import SPADE_STYLE_ResnetBlock, SPADEResnetBlock
from .base_network import BaseNetwork
from .layers import Conv2d
from .normalization import clear_seg_cache

# (attribute name, input width, output width) in units of ngf, and whether a 2x up-sampling precedes the block
_TRUNK = (("head_0", 16, 16), ("G_middle_0", 16, 16), ("G_middle_1", 16, 16),
          ("up_0", 16, 8), ("up_1", 8, 4), ("up_2", 4, 2), ("up_3", 2, 1))
_UPSAMPLINGS = {"normal": 5, "more": 6, "most": 7}


class SPADESTYLEGenerator(BaseNetwork):
    block = SPADE_STYLE_ResnetBlock

    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--num_upsampling_layers', choices=tuple(_UPSAMPLINGS), default='normal',
                            help="'more': one extra 2x up-sampling between the two middle blocks; "
                                 "'most': additionally one more up-sampling + block before the output conv")
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        nf = opt.ngf
        self.sw, self.sh = self.compute_latent_vector_size(opt)
        self.fc = Conv2d(opt.semantic_nc, 16 * nf, 3, padding=1)
        for name, cin, cout in _TRUNK:
            setattr(self, name, self.block(cin * nf, cout * nf, opt))
        last = nf
        if opt.num_upsampling_layers == 'most':
            # generator.py:45 refers to a helper that does not exist; this is the block it meant to create
            self.up_4 = self.block(nf, nf // 2, opt)
            last = nf // 2
        self.conv_img = Conv2d(last, opt.output_nc, 3, padding=1)

    def compute_latent_vector_size(self, opt):
        """(sw, sh) of the coarsest map: width crop_size / 2^n_up, height from the aspect ratio (generator.py:52-67)."""
        if opt.num_upsampling_layers not in _UPSAMPLINGS:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' % opt.num_upsampling_layers)
        sw = opt.crop_size // (2 ** _UPSAMPLINGS[opt.num_upsampling_layers])
        return sw, round(sw / opt.aspect_ratio)

    def up(self, x):
        return ops.Upsample2xFn.apply(x)

    def _schedule(self):
        """Block names in execution order, each with a flag 'up-sample first'."""
        extra = self.opt.num_upsampling_layers in ('more', 'most')
        plan = [("head_0", False), ("G_middle_0", True), ("G_middle_1", extra)]
        plan += [(n, True) for n in ("up_0", "up_1", "up_2", "up_3")]
        if self.opt.num_upsampling_layers == 'most':
            plan.append(("up_4", True))
        return plan

    def forward(self, input, w=None):
        """`self.loss_target` (optional, consumed by this call): the target image of the L1 / L2 losses, (B,1,H,W) fp32 on the
        device -- when set, the image-head kernel also reduces the two loss sums (see ops.ImageHeadFn)."""
        if self.opt.output_nc != 1:
            raise ValueError('output_nc != 1 is not supported by the B200 path (OpenEDS images are single channel)')
        clear_seg_cache()   # the im2col'd segmaps are shared by the SPADE blocks of this forward only
        ops.prepare_spectral([m for m in self.modules() if isinstance(m, Conv2d)], self.training)
        # wide label maps are zero-padded to 64 channels so that fc runs on the tensor cores like mlp_shared does
        cpad = -(-self.fc.in_channels // 64) * 64 if self.fc.in_channels >= 16 else 0
        x = self.fc.forward_nhwc(ops.seg_nearest(input, self.sh, self.sw, cpad))
        for name, upsample_first in self._schedule():
            x = getattr(self, name).forward_nhwc(x, input, w, up=upsample_first)
        clear_seg_cache()
        target, self.loss_target = getattr(self, 'loss_target', None), None
        if x.shape[-1] == 64 and self.conv_img.in_channels == 64 and ops._state["force_impl"] is None:
            # leaky_relu -> conv_img -> tanh (-> partial sums of the L1 / L2 image losses) in one kernel
            cfg = self.conv_img.cfg._replace(in_act=L.ACT_LRELU)
            img, sums = ops.ImageHeadFn.apply(x, cfg, self.conv_img.weight, self.conv_img.bias, target)
            if sums is not None:
                img._s2e_img_sums = (sums, target)
            return img
        x = self.conv_img.forward_nhwc(x, in_act=L.ACT_LRELU)     # leaky_relu(x, 0.2) -> conv_img, one kernel
        return ops.TanhFn.apply(x)


class SPADEGenerator(SPADESTYLEGenerator):
    """The original SPADE generator (no style branch) -- BASELINE config 5: plain SPADE normalisation
    (normalization.py:63-105) in the block / trunk structure of architecture.py:13-62 and generator.py:22-102, with the
    ApplyStyle term and the division by two removed.  The reference defines no such class (only SPADESTYLEGenerator,
    networks/__init__.py:9); state_dict keys = SPADESTYLEGenerator's without the `.adain.*` entries.
    forward(input, w=None): `w` is ignored."""
    block = SPADEResnetBlock

    def forward(self, input, w=None):
        return super().forward(input, None)
