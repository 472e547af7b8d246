"""Time (CUDA events) the forward / data-gradient / weight-gradient tcgen05 kernels of single convolution shapes of the
bench workload, or run each once for an `ncu --set full` capture (--once).

  python tools/conv_probe.py [--once] [B H W Cin Cout k ...]      default: the dominant shapes of the R2 / batch-16 step
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seg2eye_b200 import _lib as L, ops

DEFAULT = [(16, 640, 384, 128, 256, 3), (16, 640, 384, 128, 128, 3), (16, 640, 384, 128, 64, 3), (16, 640, 384, 64, 64, 3),
           (16, 320, 192, 128, 512, 3), (16, 320, 192, 256, 128, 3), (16, 40, 24, 1024, 1024, 3), (16, 640, 384, 64, 128, 1)]


def run(B, H, W, Cin, Cout, k, once, n=5):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16).requires_grad_()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).requires_grad_()
    b = torch.randn(Cout, device="cuda").requires_grad_()
    dy = torch.randn(B, H, W, Cout, device="cuda").to(torch.bfloat16)
    cfg = ops.ConvCfg(k, k, 1, k // 2, 0)
    reps = 1 if once else n + 2
    res = {}
    for i in range(reps):
        if i == reps - n and not once:
            ops.profile_begin()
        y = ops.tap_conv(x, cfg, (w,), (b,))
        y.backward(dy)
        x.grad = w.grad = b.grad = None
    if once:
        torch.cuda.synchronize()
        return
    prof = ops._prof
    torch.cuda.synchronize()
    for a, e, fl, tag in prof["tc"]:
        res.setdefault(tag.split()[0], []).append((a.elapsed_time(e), fl))
    ops.profile_end()
    line = "%-34s" % ("B%d %dx%d %d->%d k%d" % (B, H, W, Cin, Cout, k))
    for kind in ("fwd", "dgrad", "wgrad"):
        ms = sorted(t for t, _ in res[kind])[len(res[kind]) // 2]
        line += "  %s %.3f ms %5.0f TF/s" % (kind, ms, res[kind][0][1] / ms / 1e9)
    print(line, flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    shapes = [tuple(int(v) for v in args[i:i + 6]) for i in range(0, len(args), 6)] or DEFAULT
    for s in shapes:
        run(*s, once)
