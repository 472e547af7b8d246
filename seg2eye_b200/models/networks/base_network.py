"""BaseNetwork: parameter count print-out and the reference's weight initialisation contract
(reference models/networks/base_network.py:10-59), which `create_network` relies on."""
import torch.nn as nn
from torch.nn import init


def _init_weight(w, init_type, gain, module):
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
        module.reset_parameters()
    else:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)


def _weights_replaced(module, incompatible_keys):
    from ... import ops
    ops.bump_weights_epoch()


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()
        # the bf16 tap-major weight copies the kernels read are keyed on the master tensors' versions; a wholesale load
        # additionally raises the global epoch so that CUDA-graph replays (no host-side checks inside) re-pack first
        self.register_load_state_dict_post_hook(_weights_replaced)

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        millions = sum(p.numel() for p in self.parameters()) / 1000000
        print('Network [%s] was created. Total number of parameters: %.1f million. '
              'To see the architecture, do print(network).' % (type(self).__name__, millions))

    def init_weights(self, init_type='normal', gain=0.02):
        """Selection is by class NAME, as in the reference: modules whose class name contains 'Conv' or 'Linear' get
        their `weight` re-initialised and their bias zeroed (for spectral-norm convs `weight` aliases `weight_orig`);
        'BatchNorm2d' modules with affine parameters get N(1, gain) / 0.  `FC` (the style dense layer) is untouched."""
        def visit(m):
            name = type(m).__name__
            has_w = getattr(m, 'weight', None) is not None
            if 'BatchNorm2d' in name:
                if has_w:
                    init.normal_(m.weight.data, 1.0, gain)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)
            elif has_w and ('Conv' in name or 'Linear' in name):
                _init_weight(m.weight.data, init_type, gain, m)
                if getattr(m, 'bias', None) is not None:
                    init.constant_(m.bias.data, 0.0)

        self.apply(visit)
        for child in self.children():
            if hasattr(child, 'init_weights'):
                child.init_weights(init_type, gain)
