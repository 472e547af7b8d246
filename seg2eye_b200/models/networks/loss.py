"""Loss mirrors (reference models/networks/loss.py): GANLoss, openEDSaccuracy, MSECalculator, gram_matrix,
StyleLoss.  Reductions run in our loss-reduction kernels; results keep the reference's shapes."""
import numpy as np
import torch
import torch.nn as nn

from ... import _lib as L
from ... import ops


def _flat(x):
    """Contiguous storage view of a prediction (our NHWC-backed views or plain tensors)."""
    if x.dim() == 4 and x.dtype == torch.bfloat16 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1)
    return x.contiguous()


class GANLoss(nn.Module):
    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        self.real_label, self.fake_label = target_real_label, target_fake_label
        self.gan_mode, self.opt = gan_mode, opt
        if gan_mode not in ('ls', 'original', 'w', 'hinge'):
            raise ValueError('Unexpected gan_mode {}'.format(gan_mode))

    def loss(self, input, target_is_real, for_discriminator=True):
        """loss.py:58-83: 'original' = BCE with logits against the real / fake label, 'ls' = MSE against the label,
        'hinge', else WGAN.  Each is one reduction kernel (and one elementwise kernel in backward)."""
        x = _flat(input)
        n = x.numel()
        label = self.real_label if target_is_real else self.fake_label
        if self.gan_mode == 'original':
            return ops.reduce_loss(x, None, L.RED_BCE, 1.0 / n, label).view(())
        if self.gan_mode == 'ls':
            return ops.reduce_loss(x, None, L.RED_LS, 1.0 / n, label).view(())
        if self.gan_mode == 'hinge':
            if for_discriminator:
                kind = L.RED_HINGE_REAL if target_is_real else L.RED_HINGE_FAKE
                return ops.reduce_loss(x, None, kind, -1.0 / n).view(())
            assert target_is_real, "The generator's hinge loss must be aiming for real"
            return ops.reduce_loss(x, None, L.RED_SUM, -1.0 / n).view(())
        # wgan
        return ops.reduce_loss(x, None, L.RED_SUM, (-1.0 if target_is_real else 1.0) / n).view(())

    def __call__(self, input, target_is_real, for_discriminator=True):
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss_tensor = self.loss(pred_i, target_is_real, for_discriminator)
                loss = loss + loss_tensor.view(1)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator)


def l1_loss(a, b):
    """nn.L1Loss() (mean) between two same-shaped tensors; gradient flows to `a` only when b is detached."""
    x, y = _flat(a), _flat(b)
    return ops.reduce_loss(x, y, L.RED_L1, 1.0 / x.numel()).view(())


def mse_loss(a, b):
    x, y = _flat(a), _flat(b)
    return ops.reduce_loss(x, y, L.RED_L2, 1.0 / x.numel()).view(())


def openEDSaccuracy(produced, target):
    """loss.py:102-111: sqrt(sum (p-t)^2) / (h*w)."""
    produced, target = produced.float().contiguous(), target.float().contiguous()
    h, w = produced.shape[-2:]
    if produced.is_cuda:
        s = ops.reduce_loss(produced, target, L.RED_L2, 1.0).view(())
    else:
        s = torch.sum((produced - target) ** 2)
    return torch.sqrt(s) / (h * w)


class MSECalculator:
    @classmethod
    def calculate_mse_for_images(cls, produced, target):
        assert produced.shape == target.shape
        assert torch.min(produced) >= 0 and torch.max(produced) <= 255
        assert torch.min(target) >= 0 and torch.max(target) <= 255
        assert produced.shape[-2:] == (640, 400), f"Invalid shape: {produced.shape}"
        assert len(produced.shape) == 4, "Please feed 4D tensors"
        return torch.stack([openEDSaccuracy(produced[i], target[i]) for i in range(produced.shape[0])])

    @classmethod
    def calculate_error_statistics(cls, all_errors, mode, dataset_key):
        rel = np.sum(all_errors) / len(all_errors) * 1471
        return {f'mse/{dataset_key}/{mode}/relative': rel}


def gram_matrix(input):
    a, b, c, d = input.size()
    features = input.float().reshape(a * b, c * d)
    return torch.mm(features, features.t()).div(a * b * c * d)


class StyleLoss(nn.Module):
    def forward(self, predicted_feature, target_feature):
        return torch.nn.functional.mse_loss(gram_matrix(predicted_feature), gram_matrix(target_feature).detach())
