"""Generate tests/golden/ref_ganloss.npz: the UNMODIFIED reference GANLoss (models/networks/loss.py:17-99) in its four
modes on fixed multiscale predictions.  Runs only in the build container (needs /root/reference).

    python oracle/make_golden_ganloss.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import import_reference  # noqa: E402
from oracle import seg2eye_oracle as O  # noqa: E402


def preds(seed=77):
    """Two discriminators x five outputs (only the last of each enters the GAN loss), portable PCG64 values."""
    shapes = {"d%d_%d" % (i, j): (2, 1 + 3 * (j < 4), 9 - i, 7 - i) for i in range(2) for j in range(5)}
    st = O.synth_state(shapes, seed, scale=2.0)
    return [[st["d%d_%d" % (i, j)] for j in range(5)] for i in range(2)]


def main():
    import_reference()
    from models.networks.loss import GANLoss
    out = {}
    p = preds()
    for mode in ("hinge", "ls", "original", "w"):
        crit = GANLoss(mode, tensor=torch.FloatTensor)
        for real in (True, False):
            for for_d in (True, False):
                if mode == "hinge" and not for_d and not real:
                    continue      # the reference asserts here (loss.py:75)
                out["%s_%d_%d" % (mode, real, for_d)] = crit(p, real, for_discriminator=for_d).detach().numpy()
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "ref_ganloss.npz"), **out)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
