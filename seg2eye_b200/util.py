"""Checkpoint helpers with the reference's on-disk layout (util/util.py:195-221):
`<checkpoints_dir>/<name>/<epoch>_net_<G|D|E>.pth` = fp32 CPU state_dict with the reference's keys."""
import os

import torch

from . import ops


def save_network(net, label, epoch, opt):
    save_path = os.path.join(opt.checkpoints_dir, opt.name, '%s_net_%s.pth' % (epoch, label))
    os.makedirs(os.path.dirname(save_path), exist_ok=True)
    sd = {k: v.detach().to('cpu') for k, v in net.state_dict().items()}
    torch.save(sd, save_path)


def load_network(net, label, epoch, opt, save_dir=None):
    if save_dir is None:
        save_dir = os.path.join(opt.checkpoints_dir, opt.name)
    save_path = os.path.join(save_dir, '%s_net_%s.pth' % (epoch, label))
    weights = torch.load(save_path, map_location='cpu')
    # checkpoints written by the reference under nn.DataParallel carry a 'module.' prefix
    weights = {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in weights.items()}
    net.load_state_dict(weights)
    ops.bump_weights_epoch()
    print(f"Loaded network from {save_path} for epoch {epoch}")
    return net
