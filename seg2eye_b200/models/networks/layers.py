"""Parameter-holding layers with the reference's state_dict names, computing through seg2eye_b200.ops."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib as L
from ... import ops


class Conv2d(nn.Module):
    """nn.Conv2d replacement (weight OIHW fp32 master, optional bias) running as a tap-convolution kernel.
    With spectral=True it carries torch.nn.utils.spectral_norm's state: `weight_orig`, `weight_u`, `weight_v`
    (reference normalization.py:26, architecture.py:31-34) and `weight` aliases weight_orig's storage."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=True, spectral=False, act=L.ACT_NONE):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.cfg = ops.ConvCfg(k, k, stride, padding, act)
        self.spectral = spectral
        w = torch.empty(cout, cin, k, k)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(cin * k * k)
        if not spectral:
            self.weight = nn.Parameter(w)
        if bias:
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)
        if spectral:  # spectral_norm re-registers the weight after the bias
            self.weight_orig = nn.Parameter(w)
            self.register_buffer('weight_u', F.normalize(torch.randn(cout), dim=0, eps=1e-12))
            self.register_buffer('weight_v', F.normalize(torch.randn(cin * k * k), dim=0, eps=1e-12))

    def reset_parameters(self):
        """nn.Conv2d.reset_parameters (what --init_type none leaves in place, base_network.py:45-46)."""
        w = self.master_weight()
        nn.init.kaiming_uniform_(w.data, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(w[0].numel())
            nn.init.uniform_(self.bias.data, -bound, bound)

    def __getattr__(self, name):
        if name == 'weight' and 'weight_orig' in self._parameters:
            return self._parameters['weight_orig'].data
        return super().__getattr__(name)

    def master_weight(self):
        return self.weight_orig if self.spectral else self.weight

    def sn_state(self):
        if not self.spectral:
            return None
        ready, self._sn_ready = getattr(self, '_sn_ready', None), None
        if ready is not None:     # the enclosing network already iterated all of its layers (ops.prepare_spectral)
            inv = ready[0]
        else:
            inv = ops.spectral_inv_sigma(self.weight_orig, self.weight_u, self.weight_v, self.training)
        return (self.weight_u, self.weight_v, inv)

    def forward_nhwc(self, x, act=None, residual=None, in_act=L.ACT_NONE):
        """residual: optional tensor shaped like the output, added in the convolution's epilogue.
        in_act: activation applied to x first; fused into the kernels for the conv_img shape (64 -> 1, 3x3), a separate
        elementwise kernel otherwise."""
        cfg = self.cfg if act is None else self.cfg._replace(act=act)
        if in_act != L.ACT_NONE:
            if (self.out_channels == 1 and self.in_channels == 64 and x.shape[-1] == 64 and cfg.kh == 3 and cfg.stride == 1
                    and cfg.pad == 1):
                cfg = cfg._replace(in_act=in_act)
            else:
                x = ops.ActFn.apply(x, in_act)
        if x.shape[-1] != self.in_channels:   # zero-padded activation channels (e.g. the 16-channel D input)
            assert x.shape[-1] > self.in_channels, "input has fewer channels than the layer expects"
            cfg = cfg._replace(cin_pad=x.shape[-1])
        if ops.HEAD_ON_TC and self.out_channels < 8 and cfg.stride == 1 and self.in_channels * cfg.kh * cfg.kw >= 4096:
            # few output channels but a long reduction (PatchGAN logit head, K = 8192) on the tensor cores with the output
            # channels zero-padded to one 64-wide tile.  Off since the tiled CUDA-core head kernel (conv_thin.cu) exists:
            # the padded GEMM moves 16 taps x the whole input through shared memory for one live output column
            cfg = cfg._replace(cout_pad=64)
        biases = (self.bias,) if self.bias is not None else ()
        if (not ops.HEAD_ON_TC and residual is None and not self.spectral and cfg.cin_pad == 0
                and ops.head_conv_ok(x, cfg, self.weight)):
            return ops.head_conv(x, cfg, self.weight, self.bias)     # PatchGAN logit head: input read once, not once per tap
        return ops.tap_conv(x, cfg, (self.master_weight(),), biases, self.sn_state(), residual)

    def forward_nhwc_unscaled(self, x):
        """conv(x, weight_orig) without the 1/sigma factor and without touching u/v (batched style encoder)."""
        assert self.spectral and self.bias is None
        return ops.tap_conv(x, self.cfg, (self.weight_orig,), (), None)

    def forward(self, x):
        return ops.as_nchw_view(self.forward_nhwc(ops.as_nhwc(x)))


class Linear(nn.Module):
    """nn.Linear replacement in fp32 (encoder.py:48-49 fc_mu / fc_var)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.in_features, self.out_features = cin, cout
        bound = 1 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.empty(cout, cin).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))

    def reset_parameters(self):
        bound = 1 / math.sqrt(self.in_features)
        nn.init.uniform_(self.weight.data, -bound, bound)
        nn.init.uniform_(self.bias.data, -bound, bound)

    def forward(self, x):
        return ops.LinearFn.apply(x.float(), self.weight, self.bias, L.ACT_NONE, 0)


class BatchNorm2dStats(nn.Module):
    """Buffers of nn.BatchNorm2d(affine=False): running_mean / running_var / num_batches_tracked
    (normalization.py:75).  The arithmetic lives in the fused SPADE+Style kernel."""

    def __init__(self, c):
        super().__init__()
        self.num_features = c
        self.eps, self.momentum = 1e-5, 0.1
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))


class InstanceNorm2d(nn.Module):
    """nn.InstanceNorm2d(affine=False) (+ optionally the LeakyReLU that follows it) as one kernel."""

    def __init__(self, c, act=L.ACT_NONE):
        super().__init__()
        self.num_features, self.act = c, act

    def forward_nhwc(self, x):
        if ops._state.get("fm_pairs") and x.shape[0] % 2 == 0 and x.is_cuda:
            # [fake ; real] feature of the discriminator inside the generator step: the feature-matching L1 between the two halves
            # is reduced by the apply kernel (pix2pix_model.compute_generator_loss picks `_s2e_fm_sum` up)
            fm = torch.zeros(1, dtype=torch.float32, device=x.device)
            y = ops.InstNormFn.apply(x, self.act, None, None, None, 1, None, fm)
            y._s2e_fm_sum = (fm, y._version)
            return y
        return ops.InstNormFn.apply(x, self.act)

    def forward_nhwc_spectral(self, z, conv, n_samples):
        """z = UNSCALED output of the spectral conv `conv` over n_samples groups of images; equals what n_samples
        successive conv->norm calls (one per group) would produce, buffers of `conv` advanced n_samples times."""
        ready, conv._sn_ready = getattr(conv, '_sn_ready', None), None
        if ready is not None and ready[1] is not None and ready[0].numel() == n_samples:
            inv, U, V = ready
        else:
            inv, U, V = ops.spectral_multi(conv.weight_orig, conv.weight_u, conv.weight_v, conv.training, n_samples)
        return ops.InstNormFn.apply(z, self.act, inv, U, V, z.shape[0] // n_samples, conv.weight_orig)

    def forward(self, x):
        return ops.as_nchw_view(self.forward_nhwc(ops.as_nhwc(x)))
