"""Determinism / path-agreement diagnostic for the full generator and encoder on the GPU."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from types import SimpleNamespace
from seg2eye_b200 import _lib as L, ops
from seg2eye_b200.models import networks
from oracle import seg2eye_oracle as O
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))
oopt = O.make_opt(ngf=16, ndf=16)
d = vars(oopt).copy(); d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer")
opt = SimpleNamespace(**d)
sdG = O.synth_state(O.generator_shapes(oopt), 101); sdE = O.synth_state(O.encoder_shapes(oopt), 303)
batch = O.synth_batch(oopt, 2, 404); seg = O.one_hot(batch["label"], 4).cuda(); style = batch["style_image"].cuda()
names = ["head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"]
def run(impl):
    G = networks.SPADESTYLEGenerator(opt); G.load_state_dict({k: v.clone() for k, v in sdG.items()}); G.cuda().train()
    E = networks.ConvEncoder(opt); E.load_state_dict({k: v.clone() for k, v in sdE.items()}); E.cuda().train()
    outs = {}
    with torch.no_grad(), ops.force_impl(impl):
        mus = []
        for b in range(2):
            mu, lv, feats = E(style[b])
            mus.append(mu)
            for i, f in enumerate(feats): outs["E%d_f%d" % (b, i)] = f.float().clone()
        w = torch.stack(mus).mean(1); outs["w"] = w.clone()
        x = G.fc.forward_nhwc(ops.seg_nearest(seg, G.sh, G.sw)); outs["fc"] = x.float().clone()
        for nme in names:
            if nme not in ("head_0", "G_middle_1"): x = G.up(x)
            blk = getattr(G, nme)
            x = blk.forward_nhwc(x, seg, w); outs[nme] = x.float().clone()
        y = G.conv_img.forward_nhwc(ops.ActFn.apply(x, L.ACT_LRELU)); outs["conv_img"] = y.float().clone()
        outs["img"] = ops.TanhFn.apply(y).clone()
    return outs
a, b, c = run(None), run(None), run(L.IMPL_SIMT)
for k in a:
    print("%-22s tc-vs-tc %.2e (bitwise %s)   tc-vs-simt %.2e" % (k, rel(a[k], b[k]), bool(torch.equal(a[k], b[k])), rel(a[k], c[k])), flush=True)
