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
    Ho, Wo = H + 2 * (k // 2) - k + 1, W + 2 * (k // 2) - k + 1
    dy = torch.randn(B, Ho, Wo, Cout, device="cuda").to(torch.bfloat16)
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


def run_seg(B, H, W, once, n=5):
    """mlp_shared as the K=64 GEMM on the im2col'd segmap (SegConvFn) + ReLU."""
    seg = torch.nn.functional.one_hot(torch.randint(0, 4, (B, H, W), device="cuda"), 4).permute(0, 3, 1, 2).float().contiguous()
    col = ops.seg_im2col(seg, H, W)
    w = (torch.randn(128, 4, 3, 3, device="cuda") / 6).requires_grad_()
    b = torch.zeros(128, device="cuda").requires_grad_()
    dy = torch.randn(B, H, W, 128, device="cuda").to(torch.bfloat16)
    reps = 1 if once else n + 2
    for i in range(reps):
        if i == reps - n and not once:
            ops.profile_begin()
        y = ops.SegConvFn.apply(col, w, b, L.ACT_RELU, False)
        y.backward(dy)
        w.grad = b.grad = None
    torch.cuda.synchronize()
    if once:
        return
    res = {}
    for a, e, fl, tag in ops._prof["tc"]:
        res.setdefault(tag.split()[0], []).append((a.elapsed_time(e), fl))
    ops.profile_end()
    line = "%-34s" % ("seg B%d %dx%d 4->128 (K=64 GEMM)" % (B, H, W))
    for kind in res:
        ms = sorted(t for t, _ in res[kind])[len(res[kind]) // 2]
        line += "  %s %.3f ms %5.0f TF/s" % (kind, ms, res[kind][0][1] / ms / 1e9)
    print(line, flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    for a in sys.argv[1:]:
        if a.startswith("--dbg"):       # --dbg5=3: experimental multi-tap weight-gradient kernel (see include/seg2eye_b200.h)
            key, val = a[5:].split("=")
            L.call("s2e_debug_set", int(key), int(val))
    shapes = [tuple(int(v) for v in args[i:i + 6]) for i in range(0, len(args), 6)] or DEFAULT
    if "--seg" in sys.argv:
        run_seg(16, 640, 384, once)
    else:
        for s in shapes:
            run(*s, once)
