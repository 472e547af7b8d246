"""BASELINE config 5 on the GPU: the original (style-less) SPADE generator with a 35-class label map.
  * per layer: our SPADE module (35-channel one-hot zero-padded to 64 channels -> mlp_shared as a 64-channel tap
    convolution on the tensor cores; plain-SPADE mode of the fused normalisation kernels) against the reference-recorded
    outputs of tests/golden/ref_spade35.npz and against the oracle, forward and backward;
  * model level (the reference defines no such network: oracle restatement per SURVEY 8(c)): SPADEGenerator forward and one
    full G + D trainer iteration with --netG spade against OracleTrainer."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import seg2eye_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_ACT, TOL_CHAIN, TOL_LOSS = 1e-2, 2e-2, 2e-2


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def sub(t, n=4096):
    f = t.detach().float().cpu().reshape(-1).double()
    step = max(1, f.numel() // n)
    return f[::step][:n].float()


@pytest.mark.parametrize("name", ["batch_c64", "instance_c32"])
def test_spade_layer_35_classes(name):
    from oracle.make_golden_spade35 import CASES, inputs
    from seg2eye_b200.models import networks
    ref = np.load(os.path.join(GOLD, "ref_spade35.npz"))
    cfg, c, nc, _ = CASES[name]
    x, seg, sd = inputs(name)
    m = networks.SPADE(cfg, c, nc)
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m.cuda().train()
    xb = x.to(torch.bfloat16).float()          # what the bf16 activation path sees
    xc = xb.cuda().requires_grad_()
    y = m(xc, seg.cuda())
    y.float().square().mean().backward()
    # oracle on the same (bf16-rounded) input
    sdo = {"L." + k: v.clone() for k, v in sd.items()}
    for k, v in sdo.items():
        if v.dtype == torch.float32 and "running" not in k:
            v.requires_grad_(True)
    xr = xb.clone().requires_grad_()
    yo = O.spade(sdo, "L", xr, seg, instance="instance" in cfg)
    yo.square().mean().backward()
    assert rel(y, yo) < TOL_ACT, rel(y, yo)
    assert rel(sub(y), torch.from_numpy(ref[name + "|out_sub"])) < TOL_ACT      # the reference's own output (fp32 input)
    assert rel(xc.grad, xr.grad) < 2e-2, rel(xc.grad, xr.grad)
    g = dict(m.named_parameters())
    for k in ("mlp_shared.0.weight", "mlp_shared.0.bias", "mlp_gamma.weight", "mlp_beta.bias"):
        assert rel(g[k].grad, sdo["L." + k].grad) < 2e-2, (k, rel(g[k].grad, sdo["L." + k].grad))
    if "batch" in cfg:
        assert rel(m.param_free_norm.running_var, torch.from_numpy(ref[name + "|running_var"])) < 1e-3


def _opts(**kw):
    o = O.make_opt(ngf=16, ndf=16, label_nc=35, crop_size=128, aspect_ratio=1.0, netG="spade", **kw)
    d = vars(o).copy()
    d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer", continue_train=False,
             which_epoch="latest", checkpoints_dir="/tmp/s2e_ckpt", name="t35", no_vgg_loss=True, lambda_openeds=0.0,
             lambda_style_w=0.0, lambda_style_feat=0.0, lambda_gram=0.0, netD="multiscale")
    return o, SimpleNamespace(**d)


def test_plain_spade_generator_forward_and_state_dict():
    from seg2eye_b200.models import networks
    oopt, opt = _opts()
    shapes = O.generator_shapes(oopt)
    sd = O.synth_state(shapes, 51)
    G = networks.define_G(opt)
    assert type(G).__name__ == "SPADEGenerator"
    assert list(G.state_dict().keys()) == list(shapes.keys())           # SPADESTYLEGenerator's layout without .adain.*
    assert not any("adain" in k for k in G.state_dict())
    G.load_state_dict({k: v.clone() for k, v in sd.items()})
    G.cuda().train()
    batch = O.synth_batch(oopt, 2, 52)
    seg = O.one_hot(batch["label"], 35)
    taps = {}
    with torch.no_grad():
        ref = O.generator_forward({k: v.clone() for k, v in sd.items()}, seg, None, oopt, taps=taps)
        out = G(seg.cuda())
    assert out.shape == ref.shape == (2, 1, 128, 128)
    assert rel(out, ref) < TOL_CHAIN, rel(out, ref)


def test_plain_spade_training_iteration_vs_oracle():
    from seg2eye_b200.trainers.pix2pix_trainer import Pix2PixTrainer
    oopt, opt = _opts(lambda_l1=10.0)
    sds = dict(G=O.synth_state(O.generator_shapes(oopt), 61), D=O.synth_state(O.discriminator_shapes(oopt), 62))
    tr = Pix2PixTrainer(opt)
    m = tr.pix2pix_model
    assert m.netE is None
    m.netG.load_state_dict({k: v.clone() for k, v in sds["G"].items()})
    m.netD.load_state_dict({k: v.clone() for k, v in sds["D"].items()})
    m.netG.cuda(); m.netD.cuda()
    batch = O.synth_batch(oopt, 2, 63)
    batch.pop("style_image")
    data = {k: v.clone() for k, v in batch.items()}
    tr.run_generator_one_step(data)
    tr.run_discriminator_one_step(data)
    ours = {k: float(v.reshape(-1)[0]) for k, v in tr.get_latest_losses().items()}
    ot = O.OracleTrainer(sds["G"], sds["D"], {}, oopt)
    ot.run_generator_one_step(batch)
    ot.run_discriminator_one_step(batch)
    ref = {k: float(v.reshape(-1)[0]) for k, v in {**ot.g_losses, **ot.d_losses}.items()}
    assert set(ours) == set(ref) == {"GAN", "L1/weighted", "GAN_Feat", "D/Fake", "D/real"}
    for k in ref:
        assert abs(ours[k] - ref[k]) <= TOL_LOSS * abs(ref[k]) + (2e-2 if k == "GAN" else 0.0), (k, ours[k], ref[k])
    assert rel(tr.generated, ot.generated) < TOL_CHAIN
    # the 36-channel discriminator input is padded to 48 channels so that its first 4x4-s2 convolution is tensor-core shaped
    assert rel(m.netG.state_dict()["up_3.norm_1.spade.param_free_norm.running_var"],
               ot.sdG["up_3.norm_1.spade.param_free_norm.running_var"]) < TOL_ACT
