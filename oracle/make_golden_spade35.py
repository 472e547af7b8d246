"""Generate tests/golden/ref_spade35.npz: the reference's SPADE layer (models/networks/normalization.py:63-105) with a
35-class label map -- the per-layer pin of BASELINE config 5 (the original, style-less SPADE generator; SURVEY 8(c) last
bullet: the reference defines no such network class, so the model level is pinned through its layers).
Runs only in the build container (needs /root/reference).

    python oracle/make_golden_spade35.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import import_reference, sub  # noqa: E402
from oracle import seg2eye_oracle as O  # noqa: E402

CASES = {"batch_c64": ("spadebatch3x3", 64, 35, (2, 24, 20)), "instance_c32": ("spadeinstance3x3", 32, 35, (2, 17, 13))}


def inputs(name):
    cfg, c, nc, (b, h, w) = CASES[name]
    rng = np.random.Generator(np.random.PCG64(5000 + sorted(CASES).index(name)))
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32) * 1.5 + 0.3)
    label = torch.from_numpy(rng.integers(0, nc, size=(b, 1, 4 * h, 4 * w)).astype(np.int64))
    shapes = {"param_free_norm.running_mean": (c,), "param_free_norm.running_var": (c,), "param_free_norm.num_batches_tracked": (),
              "mlp_shared.0.weight": (128, nc, 3, 3), "mlp_shared.0.bias": (128,), "mlp_gamma.weight": (c, 128, 3, 3),
              "mlp_gamma.bias": (c,), "mlp_beta.weight": (c, 128, 3, 3), "mlp_beta.bias": (c,)}
    if "instance" in cfg:
        shapes = {k: v for k, v in shapes.items() if "param_free_norm" not in k}
    return x, O.one_hot(label, nc), O.synth_state(shapes, 6000 + sorted(CASES).index(name))


def main():
    import_reference()
    from models.networks.normalization import SPADE
    out = {}
    for name, (cfg, c, nc, _) in CASES.items():
        x, seg, sd = inputs(name)
        m = SPADE(cfg, c, nc)
        m.load_state_dict({k: v.clone() for k, v in sd.items()})
        m.train()
        xr = x.clone().requires_grad_()
        y = m(xr, seg)
        y.square().mean().backward()
        out[name + "|out_sub"], out[name + "|out_stat"] = sub(y)
        out[name + "|dx_sub"], out[name + "|dx_stat"] = sub(xr.grad)
        out[name + "|dw_shared_sub"], out[name + "|dw_shared_stat"] = sub(m.mlp_shared[0].weight.grad)
        if "batch" in cfg:
            out[name + "|running_var"] = m.param_free_norm.running_var.numpy().copy()
        print(name, tuple(y.shape), out[name + "|out_stat"])
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "ref_spade35.npz"), **out)


if __name__ == "__main__":
    main()
