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
                kind, coef = (L.RED_HINGE_REAL if target_is_real else L.RED_HINGE_FAKE), -1.0 / n
            else:
                assert target_is_real, "The generator's hinge loss must be aiming for real"
                kind, coef = L.RED_SUM, -1.0 / n
        else:   # wgan
            kind, coef = L.RED_SUM, (-1.0 if target_is_real else 1.0) / n
        # logits straight out of the discriminator's head: their sums were reduced in the kernel that produced them
        pre = ops.gan_presummed(input, kind, coef) if input is x else None
        if pre is not None:
            return pre.view(())
        return ops.reduce_loss(x, None, kind, coef).view(())

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


def _presummed(a, b, idx, kind):
    """The generator's image head already reduced sum f(a - b) for this very (a, b) pair (ops.ImageHeadFn)."""
    pre = getattr(a, '_s2e_img_sums', None)
    if pre is None or pre[1] is None or b is None or pre[1].data_ptr() != b.data_ptr() or pre[1].shape != b.shape:
        return None
    return ops.PrecomputedLossFn.apply(a, pre[1], pre[0], idx, kind, 1.0 / a.numel()).view(())


def l1_loss(a, b):
    """nn.L1Loss() (mean) between two same-shaped tensors; gradient flows to `a` only when b is detached."""
    pre = _presummed(a, b, 0, L.RED_L1)
    if pre is not None:
        return pre
    x, y = _flat(a), _flat(b)
    return ops.reduce_loss(x, y, L.RED_L1, 1.0 / x.numel()).view(())


def mse_loss(a, b):
    pre = _presummed(a, b, 1, L.RED_L2)
    if pre is not None:
        return pre
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
        """loss.py:113-133: per-image OpenEDS error of two 0..255 batches shaped (B,1,640,400).  CUDA integer batches take
        the exact-integer kernel (ops.openeds_score); anything else follows the reference's fp32 arithmetic."""
        assert produced.shape == target.shape
        assert torch.min(produced) >= 0 and torch.max(produced) <= 255, f"Min: {torch.min(produced)}, max: {torch.max(produced)}"
        assert torch.min(target) >= 0 and torch.max(target) <= 255
        assert produced.shape[-2:] == (640, 400), f"Invalid shape: {produced.shape}"
        assert len(produced.shape) == 4, "Please feed 4D tensors"
        if produced.is_cuda and not produced.is_floating_point() and not target.is_floating_point():
            return ops.openeds_score(produced, target)
        return torch.stack([openEDSaccuracy(produced[i], target[i]) for i in range(produced.shape[0])])

    @classmethod
    def calculate_mse_for_tensors(cls, produced, target):
        """loss.py:136-157 (bound as criterionOpenEDS, pix2pix_model.py:36): both tensors in [-1,1] -> 0..255 integers
        (ImageProcessor.to_255imagebatch, ending in .int(): the result carries no gradient) -> per-image error."""
        assert produced.shape == target.shape
        if not (produced.is_cuda and torch.cuda.is_current_stream_capturing()):   # range checks synchronise with the host
            assert torch.min(produced) >= -1 and torch.max(produced) <= 1, f"Min: {torch.min(produced)}, max: {torch.max(produced)}"
            assert torch.min(target) >= -1 and torch.max(target) <= 1
        assert len(produced.shape) == 4, "Please feed 4D tensors"
        return ops.openeds_score(ops.to255(produced), ops.to255(target))

    @classmethod
    def calculate_error_statistics(cls, all_errors, mode, dataset_key):
        rel = np.sum(all_errors) / len(all_errors) * 1471
        return {f'mse/{dataset_key}/{mode}/relative': rel}


def gram_matrix(input):
    """loss.py:177-189: F F^T / (a*b*c*d) with F = input viewed as (a*b, c*d); on the device through the tensor-core Gram
    GEMM (ops.gram_matrix, value only -- the differentiable path is StyleLoss)."""
    if input.is_cuda:
        return ops.gram_matrix(input)
    a, b, c, d = input.size()
    features = input.float().reshape(a * b, c * d)
    return torch.mm(features, features.t()).div(a * b * c * d)


class StyleLoss(nn.Module):
    """loss.py:192-200: mse(gram(predicted), gram(target).detach()).  Accepts the reference's NCHW tensors or our
    aggregated (B, h, w, C) fp32 feature batches (`nhwc=True`, what Pix2PixModel passes)."""

    def forward(self, predicted_feature, target_feature, nhwc=False):
        if not nhwc:
            predicted_feature = predicted_feature.float().permute(0, 2, 3, 1)
            target_feature = target_feature.float().permute(0, 2, 3, 1)
        return ops.gram_loss(predicted_feature, target_feature)
