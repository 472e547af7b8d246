"""Drop-in for the reference's `models.networks` package (models/networks/__init__.py:6-60)."""
import torch
import torch.nn as nn

from .base_network import BaseNetwork
from .discriminator import MultiscaleDiscriminator, NLayerDiscriminator
from .encoder import ConvEncoder
from .generator import SPADESTYLEGenerator
from .loss import GANLoss, MSECalculator, StyleLoss, gram_matrix, openEDSaccuracy, l1_loss, mse_loss
from .normalization import SPADE, SPADE_STYLE_Block, ApplyStyle, FC, get_nonspade_norm_layer
from .architecture import SPADE_STYLE_ResnetBlock

_REGISTRY = {
    ('spadestyle', 'generator'): SPADESTYLEGenerator,
    ('multiscale', 'discriminator'): MultiscaleDiscriminator,
    ('nlayer', 'discriminator'): NLayerDiscriminator,
    ('conv', 'encoder'): ConvEncoder,
}


def find_network_using_name(target_network_name, filename):
    key = (target_network_name.replace('_', '').lower(), filename)
    if key not in _REGISTRY:
        raise ValueError('In models.networks.%s there is no class matching %s%s' % (filename, target_network_name, filename))
    network = _REGISTRY[key]
    assert issubclass(network, BaseNetwork), "Class %s should be a subclass of BaseNetwork" % network
    return network


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    netG_cls = find_network_using_name(opt.netG, 'generator')
    parser = netG_cls.modify_commandline_options(parser, is_train)
    if is_train:
        netD_cls = find_network_using_name(opt.netD, 'discriminator')
        parser = netD_cls.modify_commandline_options(parser, is_train)
    netE_cls = find_network_using_name('conv', 'encoder')
    parser = netE_cls.modify_commandline_options(parser, is_train)
    return parser


def create_network(cls, opt):
    """models/networks/__init__.py:39-48.  One process drives one GPU (data parallelism is process-level,
    see seg2eye_b200.parallel), so several gpu_ids are not wrapped in nn.DataParallel."""
    net = cls(opt)
    net.print_network()
    if len(opt.gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.cuda()
    net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_G(opt):
    return create_network(SPADESTYLEGenerator, opt)


def define_D(opt):
    return create_network(MultiscaleDiscriminator, opt)


def define_E(opt):
    return create_network(ConvEncoder, opt)
