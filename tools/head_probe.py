"""Time the PatchGAN logit head (Cin -> 1, 4x4, stride 1, pad 2) forward and backward in tap-channel form (ops.HeadConvFn)
and, for comparison, zero-padded onto the tcgen05 kernels (the route used until round 2d).

  python tools/head_probe.py            # both D scales of the R2 / batch-16 step
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seg2eye_b200 import _lib as L, ops

SHAPES = [(32, 82, 50, 512), (32, 42, 26, 512)]


def run(B, H, W, Cin, pad_tc, n=7):
    torch.manual_seed(0)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16).requires_grad_()
    w = (torch.randn(1, Cin, 4, 4, device="cuda") / (Cin * 16) ** 0.5).requires_grad_()
    b = torch.randn(1, device="cuda").requires_grad_()
    cfg = ops.ConvCfg(4, 4, 1, 2, 0)
    if pad_tc:
        cfg = cfg._replace(cout_pad=64)
    dy = torch.randn(B, H + 1, W + 1, 1, device="cuda").to(torch.bfloat16)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for i in range(n + 2):
        ev[0].record()
        y = ops.tap_conv(x, cfg, (w,), (b,)) if pad_tc else ops.head_conv(x, cfg, w, b)
        ev[1].record()
        y.backward(dy)
        ev[2].record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(ev[0].elapsed_time(ev[1]))
            tb.append(ev[1].elapsed_time(ev[2]))
        out = (y.detach().float().clone(), x.grad.float().clone(), w.grad.clone(), b.grad.clone())
        x.grad = w.grad = b.grad = None
    tf.sort()
    tb.sort()
    print("B%d %dx%d %d->1 k4  %-18s fwd %.3f ms   bwd (dgrad + wgrad + glue) %.3f ms" % (
        B, H, W, Cin, "padded tcgen05" if pad_tc else "tap-channel form", tf[len(tf) // 2], tb[len(tb) // 2]), flush=True)
    return out


if __name__ == "__main__":
    for s in SHAPES:
        a = run(*s, pad_tc=False)
        r = run(*s, pad_tc=True)
        for name, u, v in zip(("y", "dx", "dw", "db"), a, r):
            print("    %s: rel diff between the two routes %.2e" % (name, float((u - v).norm() / v.norm())))
