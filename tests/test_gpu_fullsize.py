"""Parity at the BENCHMARKED width and resolutions (ngf = ndf = 64; R1 = 320x256 and R2 = 640x384): the shapes the
small-model tests never reach -- N = 256 / 512 / 1024 / 2048 weight tiles, K = 9216 reductions, the full-resolution
fused SPADE C = 128 (+ up-sampling) path, persistent loops over thousands of tiles.

Reference = the CPU oracle (pinned to the reference by tests/test_oracle_golden.py) on the same inputs and weights, once
with O(1)-gain synthetic weights (every gamma / beta alive) and once with the reference's own initialisation.

Tolerances:
  * module level (each E level, D levels 1-3, each generator block of the chain up to G_middle_1): 1e-2 relative L2;
  * deep ends of the chains -- discriminator levels 4-5 (measured 0.9-1.14e-2 after five bf16 convolutions), generator
    trunk after up_0 .. up_3 (measured 0.84-1.03e-2 at up_0, 1.2e-2 at up_3) and the image: the bf16-operand policy alone (bf16 inputs and weights of
    every contraction, everything else exact) already gives 0.8e-2 at up_3 and 1.3e-2 on the image (CPU emulation,
    tools/precision_floor.py, asserted in tests/test_host.py) -- bound 1.6e-2 on the trunk (measured up to 1.43e-2 at up_3), 2e-2 on the image (measured 1.08-1.19e-2);
  * losses of a full G + D iteration: 2e-2 relative (2e-2 absolute floor for the hinge-G mean of signed logits).
The measured errors are written to gpurun_out/fullsize_parity.json."""
import json
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import seg2eye_oracle as O

pytestmark = pytest.mark.gpu
TOL_ACT, TOL_TRUNK, TOL_IMAGE, TOL_LOSS = 1e-2, 1.6e-2, 2e-2, 2e-2
RES = {"R1": (256, 0.8), "R2": (384, 0.6)}
_report = {}


def rel(a, b):
    a = a.detach().float().cpu().double()
    b = b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def make_opts(res, **kw):
    crop, ar = RES[res]
    o = O.make_opt(crop_size=crop, aspect_ratio=ar, lambda_l1=10.0, **kw)
    d = vars(o).copy()
    d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer", continue_train=False,
             which_epoch="latest", checkpoints_dir="/tmp/s2e_ckpt", name="t64", no_vgg_loss=True, lambda_openeds=0.0,
             lambda_style_w=0.0, lambda_style_feat=0.0, lambda_gram=0.0, netG="spadestyle", netD="multiscale")
    return o, SimpleNamespace(**d)


def states(oopt, init):
    f = O.synth_state if init == "synth" else O.init_state
    return dict(G=f(O.generator_shapes(oopt), 21), D=f(O.discriminator_shapes(oopt), 22), E=f(O.encoder_shapes(oopt), 23))


def load(net, sd):
    net.load_state_dict({k: v.clone() for k, v in sd.items()})
    return net.cuda()


def _dump():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(_report, open(os.path.join(out, "fullsize_parity.json"), "w"), indent=1, sort_keys=True)


@pytest.mark.parametrize("res,init", [("R1", "synth"), ("R1", "ref"), ("R2", "synth"), ("R2", "ref")])
def test_forward_E_G_D_at_bench_width(res, init):
    from seg2eye_b200.models import networks
    torch.set_num_threads(os.cpu_count() or 1)
    oopt, opt = make_opts(res)
    sd = states(oopt, init)
    batch = O.synth_batch(oopt, 1, 31)
    seg = O.one_hot(batch["label"], 4)
    errs = {}
    # ---- style encoder (4 style images of the sample)
    with torch.no_grad():
        mu_o, _, feats_o = O.encoder_forward({k: v.clone() for k, v in sd["E"].items()}, batch["style_image"][0], oopt)
    E = load(networks.ConvEncoder(opt), sd["E"]).train()
    with torch.no_grad():
        mu, _, feats = E(batch["style_image"][0].cuda())
    for i, (a, b) in enumerate(zip(feats, feats_o)):
        errs["E.layer%d" % i] = rel(a, b)
    errs["E.mu"] = rel(mu, mu_o)
    # ---- generator, chained, with the oracle's per-block activations
    w = mu_o.mean(0, keepdim=True)
    taps = {}
    with torch.no_grad():
        fake_o = O.generator_forward({k: v.clone() for k, v in sd["G"].items()}, seg, w, oopt, taps=taps)
    G = load(networks.SPADESTYLEGenerator(opt), sd["G"]).train()
    from seg2eye_b200 import ops
    segc, wc = seg.cuda(), w.cuda()
    with torch.no_grad():
        ops.prepare_spectral([m for m in G.modules() if isinstance(m, networks.layers.Conv2d)], True)
        x = G.fc.forward_nhwc(ops.seg_nearest(segc, G.sh, G.sw))
        errs["G.fc"] = rel(x.permute(0, 3, 1, 2), taps["fc"])
        for name, up in G._schedule():
            x = getattr(G, name).forward_nhwc(x, segc, wc, up=up)
            errs["G." + name] = rel(x.permute(0, 3, 1, 2), taps[name])
    G2 = load(networks.SPADESTYLEGenerator(opt), sd["G"]).train()
    with torch.no_grad():
        fake = G2(segc, wc)
    errs["G.image"] = rel(fake, fake_o)
    # ---- discriminator on [fake ; real]
    both = torch.cat([torch.cat([seg, fake_o], 1), torch.cat([seg, batch["target"]], 1)], 0)
    with torch.no_grad():
        outs_o = O.discriminator_forward({k: v.clone() for k, v in sd["D"].items()}, both, oopt)
    D = load(networks.MultiscaleDiscriminator(opt), sd["D"]).train()
    with torch.no_grad():
        outs = D(both.cuda())
    for i in range(2):
        for j in range(5):
            assert outs[i][j].shape == outs_o[i][j].shape
            errs["D.%d.%d" % (i, j)] = rel(outs[i][j], outs_o[i][j])
    _report["forward %s %s" % (res, init)] = {k: round(v, 5) for k, v in errs.items()}
    _dump()
    for k, v in errs.items():
        if k == "G.image":
            tol = TOL_IMAGE
        elif k in ("G.up_0", "G.up_1", "G.up_2", "G.up_3") or k.endswith(".3") or k.endswith(".4"):
            tol = TOL_TRUNK     # deep end of a chain: generator trunk after 4+ blocks, discriminator levels 4 and 5
        else:
            tol = TOL_ACT
        assert v < tol, (k, v, errs)


@pytest.mark.parametrize("res,bs,init", [("R1", 2, "synth"), ("R1", 1, "ref"), ("R2", 1, "ref")])
def test_trainer_iteration_at_bench_width(res, bs, init):
    """One full Pix2PixTrainer iteration (G step + D step, Adam) at ngf = ndf = 64 against OracleTrainer."""
    from seg2eye_b200.trainers.pix2pix_trainer import Pix2PixTrainer
    torch.set_num_threads(os.cpu_count() or 1)
    oopt, opt = make_opts(res)
    sd = states(oopt, init)
    batch = O.synth_batch(oopt, bs, 41)
    tr = Pix2PixTrainer(opt)
    m = tr.pix2pix_model
    for net, k in ((m.netG, "G"), (m.netD, "D"), (m.netE, "E")):
        load(net, sd[k])
    data = {k: v.clone() for k, v in batch.items()}
    tr.run_generator_one_step(data)
    tr.run_discriminator_one_step(data)
    ours = {k: float(v.reshape(-1)[0]) for k, v in tr.get_latest_losses().items()}
    ot = O.OracleTrainer(sd["G"], sd["D"], sd["E"], oopt)
    ot.run_generator_one_step(batch)
    ot.run_discriminator_one_step(batch)
    ref = {k: float(v.reshape(-1)[0]) for k, v in {**ot.g_losses, **ot.d_losses}.items()}
    img_err = rel(tr.generated, ot.generated)
    _report["iteration %s B%d %s" % (res, bs, init)] = {"ours": ours, "reference": ref, "image": round(img_err, 5)}
    _dump()
    assert set(ours) == set(ref)
    for k in ref:
        assert abs(ours[k] - ref[k]) <= TOL_LOSS * abs(ref[k]) + (2e-2 if k == "GAN" else 0.0), (k, ours[k], ref[k])
    assert img_err < TOL_IMAGE, img_err
    # post-step buffers.  They were last touched by the D step's generator pass, i.e. on the weights AFTER the generator's
    # Adam step; Adam(beta1 = 0) moves every weight by lr * sign(g), so weights whose tiny gradient flips sign under bf16
    # noise move the other way.  With O(1)-gain weights that is invisible (lr 1e-4 against weights ~3e-2); with the
    # reference initialisation (weights ~1e-3) the second forward differs by 1e-2 .. 1e-1 in u / v and the running
    # statistics (measured 1.4e-2, 2.0e-2, 1.3e-1), so there only the step counters are compared
    post = m.netG.state_dict()
    if init == "synth":
        for k in ("up_3.norm_1.spade.param_free_norm.running_var", "head_0.norm_0.spade.param_free_norm.running_mean",
                  "up_3.conv_0.weight_u", "up_1.conv_s.weight_v"):
            assert rel(post[k], ot.sdG[k]) < TOL_ACT, k
    assert int(post["up_2.norm_0.spade.param_free_norm.num_batches_tracked"]) == int(
        ot.sdG["up_2.norm_0.spade.param_free_norm.num_batches_tracked"])
