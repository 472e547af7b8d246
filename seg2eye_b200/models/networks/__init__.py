"""Drop-in for the reference's `models.networks` package (models/networks/__init__.py:6-60): network factory,
option hooks and the loss re-exports `Pix2PixModel` / `util.tester` import from here."""
import torch

from .architecture import SPADE_STYLE_ResnetBlock, SPADEResnetBlock
from .base_network import BaseNetwork
from .discriminator import MultiscaleDiscriminator, NLayerDiscriminator
from .encoder import ConvEncoder
from .generator import SPADEGenerator, SPADESTYLEGenerator
from .loss import GANLoss, MSECalculator, StyleLoss, gram_matrix, l1_loss, mse_loss, openEDSaccuracy
from .normalization import FC, SPADE, ApplyStyle, SPADE_STYLE_Block, get_nonspade_norm_layer

# the reference resolves '<name><kind>' by scanning a module for a class of that lower-cased name; the hot path has
# exactly these four
_NETWORKS = {'spadestylegenerator': SPADESTYLEGenerator, 'spadegenerator': SPADEGenerator, 'multiscalediscriminator': MultiscaleDiscriminator,
             'nlayerdiscriminator': NLayerDiscriminator, 'convencoder': ConvEncoder}


def find_network_using_name(target_network_name, filename):
    cls = _NETWORKS.get((target_network_name + filename).replace('_', '').lower())
    if cls is None:
        raise ValueError('models.networks.%s has no network called %r' % (filename, target_network_name))
    assert issubclass(cls, BaseNetwork), "Class %s should be a subclass of BaseNetwork" % cls
    return cls


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    wanted = [(opt.netG, 'generator')] + ([(opt.netD, 'discriminator')] if is_train else []) + [('conv', 'encoder')]
    for name, kind in wanted:
        parser = find_network_using_name(name, kind).modify_commandline_options(parser, is_train)
    return parser


def create_network(cls, opt):
    """Instantiate, report size, move to the process's GPU, initialise.  One process drives one GPU (data parallelism
    is process-level, seg2eye_b200.parallel), so there is no nn.DataParallel wrapper as in the reference."""
    net = cls(opt)
    net.print_network()
    if opt.gpu_ids:
        assert torch.cuda.is_available()
        net.cuda()
    net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_G(opt):
    # --netG spadestyle (the reference's only generator) | spade (the style-less original, BASELINE config 5)
    return create_network(find_network_using_name(getattr(opt, 'netG', 'spadestyle'), 'generator'), opt)


def define_D(opt):
    return create_network(MultiscaleDiscriminator, opt)


def define_E(opt):
    return create_network(ConvEncoder, opt)
