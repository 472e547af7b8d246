"""BaseNetwork mirror (reference models/networks/base_network.py:10-59): same init_weights / print_network
contract, so `create_network` and the option plumbing behave as in the reference."""
import torch.nn as nn
from torch.nn import init


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        n = sum(p.numel() for p in self.parameters())
        print('Network [%s] was created. Total number of parameters: %.1f million. '
              'To see the architecture, do print(network).' % (type(self).__name__, n / 1000000))

    def init_weights(self, init_type='normal', gain=0.02):
        """Every module whose class name contains Conv/Linear gets its `weight` initialised (for spectral-norm
        convs `weight` aliases `weight_orig`, as in torch's spectral_norm), bias <- 0 (base_network.py:28-59)."""
        def init_func(m):
            cn = m.__class__.__name__
            if cn.find('BatchNorm2d') != -1:
                if getattr(m, 'weight', None) is not None:
                    init.normal_(m.weight.data, 1.0, gain)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)
            elif hasattr(m, 'weight') and (cn.find('Conv') != -1 or cn.find('Linear') != -1):
                w = m.weight.data
                if init_type == 'normal':
                    init.normal_(w, 0.0, gain)
                elif init_type == 'xavier':
                    init.xavier_normal_(w, gain=gain)
                elif init_type == 'xavier_uniform':
                    init.xavier_uniform_(w, gain=1.0)
                elif init_type == 'kaiming':
                    init.kaiming_normal_(w, a=0, mode='fan_in')
                elif init_type == 'orthogonal':
                    init.orthogonal_(w, gain=gain)
                elif init_type == 'none':
                    m.reset_parameters()
                else:
                    raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)

        self.apply(init_func)
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights(init_type, gain)
