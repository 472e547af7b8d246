"""SPADESTYLEGenerator mirror (reference models/networks/generator.py:13-102)."""
from ... import _lib as L
from ... import ops
from .architecture import SPADE_STYLE_ResnetBlock
from .base_network import BaseNetwork
from .layers import Conv2d


class SPADESTYLEGenerator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument('--num_upsampling_layers', choices=('normal', 'more', 'most'), default='normal',
                            help="If 'more', adds upsampling layer between the two middle resnet blocks. "
                                 "If 'most', also add one more upsampling + resnet layer at the end of the generator")
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        nf = opt.ngf
        self.sw, self.sh = self.compute_latent_vector_size(opt)
        self.fc = Conv2d(opt.semantic_nc, 16 * nf, 3, padding=1)
        self.head_0 = SPADE_STYLE_ResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_0 = SPADE_STYLE_ResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_1 = SPADE_STYLE_ResnetBlock(16 * nf, 16 * nf, opt)
        self.up_0 = SPADE_STYLE_ResnetBlock(16 * nf, 8 * nf, opt)
        self.up_1 = SPADE_STYLE_ResnetBlock(8 * nf, 4 * nf, opt)
        self.up_2 = SPADE_STYLE_ResnetBlock(4 * nf, 2 * nf, opt)
        self.up_3 = SPADE_STYLE_ResnetBlock(2 * nf, 1 * nf, opt)
        final_nc = nf
        if opt.num_upsampling_layers == 'most':
            # the reference calls an undefined helper here (generator.py:45); we build the block it intended
            self.up_4 = SPADE_STYLE_ResnetBlock(1 * nf, nf // 2, opt)
            final_nc = nf // 2
        self.conv_img = Conv2d(final_nc, opt.output_nc, 3, padding=1)

    def compute_latent_vector_size(self, opt):
        if opt.num_upsampling_layers == 'normal':
            num_up_layers = 5
        elif opt.num_upsampling_layers == 'more':
            num_up_layers = 6
        elif opt.num_upsampling_layers == 'most':
            num_up_layers = 7
        else:
            raise ValueError('opt.num_upsampling_layers [%s] not recognized' % opt.num_upsampling_layers)
        sw = opt.crop_size // (2 ** num_up_layers)
        sh = round(sw / opt.aspect_ratio)
        return sw, sh

    def up(self, x):
        return ops.Upsample2xFn.apply(x)

    def forward(self, input, w=None):
        from .normalization import clear_seg_cache
        clear_seg_cache()   # im2col'd segmaps are shared by the SPADE blocks of one forward only
        seg = input
        x = self.fc.forward_nhwc(ops.seg_nearest(seg, self.sh, self.sw))
        x = self.head_0.forward_nhwc(x, seg, w)
        x = self.up(x)
        x = self.G_middle_0.forward_nhwc(x, seg, w)
        if self.opt.num_upsampling_layers in ('more', 'most'):
            x = self.up(x)
        x = self.G_middle_1.forward_nhwc(x, seg, w)
        x = self.up(x)
        x = self.up_0.forward_nhwc(x, seg, w)
        x = self.up(x)
        x = self.up_1.forward_nhwc(x, seg, w)
        x = self.up(x)
        x = self.up_2.forward_nhwc(x, seg, w)
        x = self.up(x)
        x = self.up_3.forward_nhwc(x, seg, w)
        if self.opt.num_upsampling_layers == 'most':
            x = self.up(x)
            x = self.up_4.forward_nhwc(x, seg, w)
        x = self.conv_img.forward_nhwc(ops.ActFn.apply(x, L.ACT_LRELU))
        clear_seg_cache()
        if self.opt.output_nc == 1:
            return ops.TanhFn.apply(x)
        raise ValueError('output_nc != 1 is not supported by the B200 path (OpenEDS images are single channel)')
