"""Micro-benchmark of the forward tap-convolution kernel's epilogue (timing experiments with debug key 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seg2eye_b200 import _lib as L, ops
def run(B, H, W, Cin, Cout, k, dbg, bias=True, n=10):
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5)
    b = torch.randn(Cout, device="cuda") if bias else None
    cfg = ops.ConvCfg(k, k, 1, k // 2, 0)
    L.call("s2e_debug_set", 3, dbg)
    with torch.no_grad():
        for _ in range(3):
            ops.tap_conv(x, cfg, (w,), (b,) if bias else ())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.tap_conv(x, cfg, (w,), (b,) if bias else ())
        e1.record(); torch.cuda.synchronize()
    L.call("s2e_debug_set", 3, 0)
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    tiles = (B * H * W + 127) // 128 * ((Cout + (256 if Cout >= 256 else 128 if Cout >= 128 else 64) - 1) // (256 if Cout >= 256 else 128 if Cout >= 128 else 64))
    return ms, fl / ms / 1e9, tiles
for shape in [(16, 640, 384, 64, 128, 1), (16, 640, 384, 128, 256, 3), (16, 160, 96, 512, 256, 3)]:
    for dbg, name in [(0, "full"), (1, "no-store"), (2, "no-math"), (3, "no-store,no-math"), (7, "no-store,no-math,no-tmemld")]:
        for bias in (True, False):
            ms, tf, tiles = run(*shape, dbg, bias)
            print("%-28s %-28s bias=%d  %.3f ms  %.0f TFLOP/s  %.2f us/tile/SM" % (shape, name, bias, ms, tf, ms * 1e3 / (tiles / 148)), flush=True)
