"""Generate tests/golden/ref_tail.npz: the inference / validation tail of the UNMODIFIED reference on synthetic images
(util/tester.py:44-47,96-100 -> data/postprocessor.py:57-114 ImageProcessor.to_255resized_imagebatch, and
models/networks/loss.py:102-157 MSECalculator).  Runs only in the build container (needs /root/reference and cv2).

    python oracle/make_golden_tail.py

Inputs are rebuilt by the tests from `tail_inputs(seed)` below (portable numpy PCG64); the fixture stores, per case, the
SHA-256 of the full int32 result (bit-exact check over every pixel), a strided subsample of it and the per-image errors.
One shim beyond make_golden.py's: `np.float` (removed from numpy 1.24+, used at postprocessor.py:110)."""
import hashlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import import_reference  # noqa: E402

CASES = {"R1": (2, 320, 256), "R2": (1, 640, 384), "odd": (1, 37, 29), "wide": (1, 640, 512)}


def tail_inputs(name):
    """fake images in [-1,1] with saturated regions (tanh of a wide normal), an int target in [0,255]."""
    n, h, w = CASES[name]
    rng = np.random.Generator(np.random.PCG64(900 + sorted(CASES).index(name)))
    fake = np.tanh(rng.standard_normal((n, 1, h, w)) * 3).astype(np.float32)
    fake[:, :, :3, :3] = 1.0
    fake[:, :, -3:, -3:] = -1.0
    target = rng.integers(0, 256, size=(n, 1, 640, 400)).astype(np.int32)
    return torch.from_numpy(fake), torch.from_numpy(target)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).digest(), dtype=np.uint8).copy()


def main():
    import_reference()
    if not hasattr(np, "float"):
        np.float = float
    from data.postprocessor import ImageProcessor
    from models.networks.loss import MSECalculator
    out = {}
    for name in CASES:
        fake, target = tail_inputs(name)
        resized = ImageProcessor.to_255resized_imagebatch(fake, as_tensor=True)
        assert resized.shape == (fake.shape[0], 1, 640, 400) and resized.dtype == torch.int32
        errs = MSECalculator.calculate_mse_for_images(resized, target)
        out[name + "|sha256"] = digest(resized.numpy())
        out[name + "|sub"] = resized.numpy().reshape(-1)[::997].astype(np.uint8)
        out[name + "|errors"] = errs.numpy().astype(np.float64)
        print(name, tuple(resized.shape), errs.numpy())
    # the loss-side variant (pix2pix_model.py:36,209-213): no resize, fp32 tensors in [-1,1]
    rng = np.random.Generator(np.random.PCG64(77))
    a = torch.from_numpy(rng.uniform(-1, 1, size=(3, 1, 64, 48)).astype(np.float32))
    b = torch.from_numpy(rng.uniform(-1, 1, size=(3, 1, 64, 48)).astype(np.float32))
    out["tensors|errors"] = MSECalculator.calculate_mse_for_tensors(a, b).numpy().astype(np.float64)
    out["tensors|sha256"] = digest(ImageProcessor.to_255imagebatch(a).numpy())
    print("tensors", out["tensors|errors"])
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "ref_tail.npz"), **out)


if __name__ == "__main__":
    main()
