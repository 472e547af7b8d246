"""torch.optim.Adam semantics (reference pix2pix_model.py:92-110) with the update done by our fused kernel."""
import torch

from . import _lib as L
from . import ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=(float(betas[0]), float(betas[1])), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        st = L.stream()
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state['step'] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                assert p.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
                L.call("s2e_adam_step", L.ptr(p), L.ptr(g), L.ptr(state['exp_avg']), L.ptr(state['exp_avg_sq']),
                       p.numel(), group['lr'], b1, b2, group['eps'], group['weight_decay'], state['step'], st)
        ops.bump_weights_epoch()
