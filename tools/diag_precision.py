"""GPU diagnostic: error levels of the tap-convolution implementations (tcgen05 vs SIMT vs fp32 torch) and of the
generator blocks fed with oracle inputs.  Writes gpurun_out/diag_precision.txt"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.nn.functional as F
from seg2eye_b200 import _lib as L, ops
from oracle import seg2eye_oracle as O

out = open(os.path.join(REPO, "gpurun_out", "diag_precision.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()
bf = lambda x: x.to(torch.bfloat16).float()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-12))
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
nchw = lambda y: y.float().permute(0, 3, 1, 2).cpu()

cases = [(2, 128, 128, 10, 8, 3, 1, 1), (2, 128, 256, 20, 16, 3, 1, 1), (1, 256, 128, 40, 32, 3, 1, 1), (1, 128, 64, 40, 32, 1, 1, 0),
         (3, 64, 128, 21, 17, 4, 2, 2), (2, 64, 128, 11, 9, 4, 1, 2), (2, 64, 128, 32, 32, 3, 2, 1), (4, 512, 512, 20, 16, 3, 1, 1),
         (2, 256, 256, 80, 64, 3, 1, 1), (2, 128, 512, 160, 128, 3, 1, 1)]
for (B, Cin, Cout, H, W, k, s, p) in cases:
    g = torch.Generator().manual_seed(0)
    x = bf(torch.randn(B, Cin, H, W, generator=g)); w = bf(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    yr = F.conv2d(xr, wr, None, stride=s, padding=p); dy = bf(torch.randn(yr.shape, generator=g)); yr.backward(dy)
    res = {}
    for name, impl in (("tc", L.IMPL_TC), ("tc2", L.IMPL_TC), ("simt", L.IMPL_SIMT)):
        xc, wc = nhwc(x).requires_grad_(), w.cuda().requires_grad_()
        with ops.force_impl(impl):
            y = ops.tap_conv(xc, ops.ConvCfg(k, k, s, p, 0), (wc,), ())
            y.backward(nhwc(dy))
        torch.cuda.synchronize()
        res[name] = (nchw(y), nchw(xc.grad), wc.grad.cpu())
    P("case", (B, Cin, Cout, H, W, k, s, p))
    for i, nm in enumerate(("fwd", "dgrad", "wgrad")):
        ref = (yr.detach(), xr.grad, wr.grad)[i]
        a, b, c = res["tc"][i], res["tc2"][i], res["simt"][i]
        P("   %-6s tc-vs-ref %.5f simt-vs-ref %.5f tc-vs-simt %.6f maxabs(tc-simt) %.4g  bitwise-det %s  refmax %.3g" % (
            nm, rel(a, ref), rel(c, ref), rel(a, c), float((a - c).abs().max()), bool(torch.equal(a, b)), float(ref.abs().max())))

# ---- generator blocks, each fed with the ORACLE's input (module-level parity)
from types import SimpleNamespace
from seg2eye_b200.models import networks
oopt = O.make_opt(ngf=16, ndf=16)
d = vars(oopt).copy(); d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer")
opt = SimpleNamespace(**d)
sd = O.synth_state(O.generator_shapes(oopt), 101)
batch = O.synth_batch(oopt, 2, 404); seg = O.one_hot(batch["label"], 4)
w = O.synth_state({"w": (2, 16)}, 5, scale=4.0)["w"]
taps = {}
with torch.no_grad():
    O.generator_forward({k: v.clone() for k, v in sd.items()}, seg, w, oopt, taps=taps)
names = ["head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"]
for impl in (None, L.IMPL_SIMT):
    G = networks.SPADESTYLEGenerator(opt); G.load_state_dict({k: v.clone() for k, v in sd.items()}); G.cuda().train()
    segc, wc = seg.cuda(), w.cuda()
    errs_chain, errs_fed = {}, {}
    with torch.no_grad(), ops.force_impl(impl):
        x = G.fc.forward_nhwc(ops.seg_nearest(segc, G.sh, G.sw))
        prev = "fc"
        for nme in names:
            if nme not in ("head_0", "G_middle_1"):
                x = G.up(x)
            # module-level: feed oracle input
            xin = taps[prev]
            if nme not in ("head_0", "G_middle_1"):
                xin = xin.repeat_interleave(2, 2).repeat_interleave(2, 3)
            G2 = networks.SPADESTYLEGenerator(opt); G2.load_state_dict({k: v.clone() for k, v in sd.items()}); G2.cuda().train()
            yfed = getattr(G2, nme).forward_nhwc(nhwc(xin), segc, wc)
            errs_fed[nme] = round(rel(nchw(yfed), taps[nme]), 5)
            x = getattr(G, nme).forward_nhwc(x, segc, wc)
            errs_chain[nme] = round(rel(nchw(x), taps[nme]), 5)
            prev = nme
    P("impl", impl, "chained", errs_chain)
    P("impl", impl, "oracle-fed", errs_fed)
