"""Device-side inference / validation tail.

The reference finishes every validation / test batch on the host (util/tester.py:44-47,96-100): `.detach().cpu()`, one
`cv2.resize(INTER_LINEAR)` per image in float64 to the raw OpenEDS frame 640x400, `(x + 1) * 255 / 2`, `.int()`
(data/postprocessor.py:57-114), then the per-image OpenEDS score (models/networks/loss.py:102-133).  Here the same
integers and scores are produced on the GPU by one kernel pass (s2e_to255_resize), so that a batch-64 inference sweep
(BASELINE config 4) never waits for the host.

    ImageProcessor.to_255resized_imagebatch(fake)             # same name / defaults as the reference's class
    errors, fake, fake_resized = validation_tail(model, data_i)  # Tester.run_batch for a CUDA model
"""
import torch

from . import ops


class ImageProcessor:
    """The two ImageProcessor entry points of the reference that sit on the hot inference tail, for CUDA tensors."""

    @classmethod
    def to_255resized_imagebatch(cls, image, w=400, h=640, as_tensor=True):
        if image.dim() == 3:
            image = image.unsqueeze(0)
        out = ops.to255_resize(image, (h, w))
        return out if as_tensor else out.cpu().numpy()

    @classmethod
    def to_255imagebatch(cls, image, as_tensor=True):
        out = ops.to255(image)
        return out if as_tensor else out.cpu().numpy()


def validation_tail(model, data_i, size=(640, 400)):
    """Tester.run_batch (util/tester.py:96-100) without leaving the device: inference, resize to the raw frame, 0..255
    integers, per-image OpenEDS error against data_i['target_original'] (int, (B,1,640,400) or (B,640,400)).
    Returns (errors (B,) fp32, fake (B,1,h,w) fp32, fake_resized (B,1,640,400) int32), all CUDA tensors."""
    fake = model.forward(data_i, mode="inference").detach()
    target = data_i["target_original"]
    if target.dim() == 3:
        target = target.unsqueeze(1)
    fake_resized, errors = ops.to255_resize(fake, size, target=target)
    return errors, fake, fake_resized
