"""Time (CUDA events) or run once (--once, for `ncu --set full`) the SPADE+Style normalisation kernels of the bench
workload: statistics, forward modulation, backward (reduce + fold + apply), with and without the nearest-2x index map,
plus the fused gamma|beta-conv + modulation kernel in training mode (SpadeConvFn).

  python tools/norm_probe.py [--once] [--fused] [B H W C up ...]        default: the full-resolution shapes of the R2 / batch-16 step
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seg2eye_b200 import _lib as L, ops

DEFAULT = [(16, 640, 384, 128, 1), (16, 640, 384, 64, 0), (16, 320, 192, 256, 1), (16, 160, 96, 512, 1)]


def run(B, H, W, C, up, once, fused, n=5):
    hx, wx = (H // 2, W // 2) if up else (H, W)
    x = torch.randn(B, hx, wx, C, device="cuda").to(torch.bfloat16).requires_grad_()
    style = (torch.randn(B, 2 * C, device="cuda") * 0.3).requires_grad_()
    dout = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    cfg = ops.NormCfg(False, L.ACT_LRELU, True, 0.1, 1e-5)
    rm, rv, nbt = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda"), torch.tensor(0, device="cuda")
    if fused:
        actv = torch.relu(torch.randn(B, H, W, 128, device="cuda")).to(torch.bfloat16).requires_grad_()
        wg = (torch.randn(C, 128, 3, 3, device="cuda") / 34).requires_grad_()
        wb = (torch.randn(C, 128, 3, 3, device="cuda") / 34).requires_grad_()
        bg, bb = torch.zeros(C, device="cuda").requires_grad_(), torch.zeros(C, device="cuda").requires_grad_()
        ccfg = ops.ConvCfg(3, 3, 1, 1, L.ACT_NONE)
        fn = lambda: ops.SpadeConvFn.apply(actv, x, style, wg, wb, bg, bb, ccfg, cfg, rm, rv, nbt, bool(up), None)
    else:
        gb = torch.randn(B, H, W, 2 * C, device="cuda").to(torch.bfloat16).requires_grad_()
        fn = lambda: ops.SpadeStyleFn.apply(x, gb, style, cfg, rm, rv, nbt, bool(up), None)
    reps = 1 if once else n + 2
    ev = []
    for i in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        y = fn()
        e[1].record()
        y.backward(dout)
        e[2].record()
        ev.append(e)
        for t in (x, style):
            t.grad = None
    torch.cuda.synchronize()
    if once:
        return
    f = sorted(a.elapsed_time(b) for a, b, _ in ev[2:])[n // 2]
    bw = sorted(b.elapsed_time(c) for _, b, c in ev[2:])[n // 2]
    el = B * H * W * C
    print("%-28s%s fwd(+stats) %.3f ms  %5.0f GB/s (8 B/el)   bwd %.3f ms  %5.0f GB/s (12 B/el)" % (
        "B%d %dx%d C%d%s" % (B, H, W, C, " up" if up else ""), " fused-conv" if fused else "", f, 8 * el / f / 1e6, bw, 12 * el / bw / 1e6), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once, fused = "--once" in sys.argv, "--fused" in sys.argv
    for a in sys.argv[1:]:
        if a.startswith("--dbg"):
            key, val = a[5:].split("=")
            L.call("s2e_debug_set", int(key), int(val))
    shapes = [tuple(int(v) for v in args[i:i + 5]) for i in range(0, len(args), 5)] or DEFAULT
    for s in shapes:
        run(*s, once, fused)
