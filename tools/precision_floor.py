"""CPU emulation of the precision policy on the chained SPADE+Style generator: where does the
relative error of the chained forward come from, and what is the floor that bf16 OPERANDS alone impose?

The fp32 oracle generator is re-evaluated with bf16 rounding switched on at chosen points:
    x     conv inputs (the SPADE+Style outputs)        w      weights of every contraction
    actv  ReLU(mlp_shared) (input of gamma|beta)       gb     gamma | beta outputs
    dx    conv_0 outputs        skip  conv_s outputs   trunk  block outputs x_s + conv_1(.) and G.fc
    ximg / wimg   input / weight of conv_img
and compared with the unrounded evaluation (relative L2 per block of the chain and on the image).

    python tools/precision_floor.py [--ngf 16] [--init synth|ref]

Result (ngf 16, 320x256, batch 2; the same for reference-initialised weights): all points 1.1e-2 at up_3 / 1.8e-2 on the
image (the device measures 1.2e-2 / 1.7-1.9e-2); operands only {x, w, actv} -- every stored tensor fp32, only the tensor
-core inputs bf16 -- 0.82e-2 / 1.36e-2.  The image error is ~1.55x the up_3 error because conv_img contracts 576 noisy
inputs into one channel.  A chained-image tolerance of 1e-2 is therefore below what "bf16 inputs to every tensor-core
contraction" can deliver for this 7-block network; module-level activations (each block fed with the reference's input)
sit at 3.5-4.4e-3.  tests/test_host.py::test_bf16_operand_floor_of_the_chained_generator pins these numbers."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import seg2eye_oracle as O  # noqa: E402

ALL = frozenset({"x", "w", "gb", "actv", "dx", "trunk", "skip", "ximg", "wimg"})
OPERANDS = frozenset({"x", "w", "actv", "ximg", "wimg"})


def _r(t, key, fl):
    return t.to(torch.bfloat16).float() if key in fl else t


def _spade(sd, p, x, seg, fl):
    normalized = O.batch_norm_train(x, sd, p + ".param_free_norm")
    seg_r = O.nearest_resize(seg, x.shape[2:])
    actv = _r(F.relu(F.conv2d(seg_r, _r(sd[p + ".mlp_shared.0.weight"], "w", fl), sd[p + ".mlp_shared.0.bias"], padding=1)), "actv", fl)
    gamma = _r(F.conv2d(actv, _r(sd[p + ".mlp_gamma.weight"], "w", fl), sd[p + ".mlp_gamma.bias"], padding=1), "gb", fl)
    beta = _r(F.conv2d(actv, _r(sd[p + ".mlp_beta.weight"], "w", fl), sd[p + ".mlp_beta.bias"], padding=1), "gb", fl)
    return normalized * (1 + gamma) + beta


def _block(sd, p, x, seg, w, fl):
    return (_spade(sd, p + ".spade", x, seg, fl) + O.apply_style(sd, p + ".adain", x, w)) / 2


def _resblock(sd, p, x, seg, w, fin, fout, fl, taps):
    def conv(name, t, pad):
        wt = O.spectral_weight(sd, "%s.%s" % (p, name), True)
        return F.conv2d(t, _r(wt, "w", fl), sd.get("%s.%s.bias" % (p, name)), padding=pad)
    x_s = _r(conv("conv_s", _r(_block(sd, p + ".norm_s", x, seg, w, fl), "x", fl), 0), "skip", fl) if fin != fout else x
    n0 = _r(F.leaky_relu(_block(sd, p + ".norm_0", x, seg, w, fl), 0.2), "x", fl)
    dx = _r(conv("conv_0", n0, 1), "dx", fl)
    n1 = _r(F.leaky_relu(_block(sd, p + ".norm_1", dx, seg, w, fl), 0.2), "x", fl)
    taps[p] = out = _r(x_s + conv("conv_1", n1, 1), "trunk", fl)
    return out


def generator(sd, seg, w, opt, fl, taps):
    """oracle.generator_forward ('normal' up-sampling) with bf16 rounding at the points named in `fl`."""
    sw, sh = O.latent_size(opt)
    x = _r(F.conv2d(O.nearest_resize(seg, (sh, sw)), _r(sd["fc.weight"], "w", fl), sd["fc.bias"], padding=1), "trunk", fl)
    blocks = O.generator_blocks(opt)
    up = lambda t: t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    for i, (name, fin, fout) in enumerate(blocks):
        if i not in (0, 2):
            x = up(x)
        x = _resblock(sd, name, x, seg, w, fin, fout, fl, taps)
    x = F.conv2d(_r(F.leaky_relu(x, 0.2), "ximg", fl), _r(sd["conv_img.weight"], "wimg", fl), sd["conv_img.bias"], padding=1)
    return torch.tanh(x)


def measure(flags, ngf=16, init="synth", batch=2):
    """-> ({block: rel err}, image rel err) of the rounded evaluation against the unrounded one."""
    oopt = O.make_opt(ngf=ngf, ndf=ngf)
    shapes = O.generator_shapes(oopt)
    sd = O.synth_state(shapes, 101) if init == "synth" else O.init_state(shapes, 101)
    seg = O.one_hot(O.synth_batch(oopt, batch, 404)["label"], 4)
    w = O.synth_state({"w": (batch, 16)}, 5, scale=4.0)["w"]
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    t0, t1 = {}, {}
    with torch.no_grad():
        ref = generator({k: v.clone() for k, v in sd.items()}, seg, w, oopt, frozenset(), t0)
        out = generator({k: v.clone() for k, v in sd.items()}, seg, w, oopt, frozenset(flags), t1)
    return {n: rel(t1[n], t0[n]) for n in t0}, rel(out, ref)


if __name__ == "__main__":
    ngf = int(sys.argv[sys.argv.index("--ngf") + 1]) if "--ngf" in sys.argv else 16
    init = sys.argv[sys.argv.index("--init") + 1] if "--init" in sys.argv else "synth"
    for name, fl in (("all rounding points", ALL), ("fp32 trunk", ALL - {"trunk"}), ("fp32 conv outputs", ALL - {"trunk", "skip", "dx", "gb"}),
                     ("operands only (floor)", OPERANDS), ("inputs only", {"x", "actv", "ximg"}), ("weights only", {"w", "wimg"})):
        blocks, img = measure(fl, ngf, init)
        print("%-24s up_0 %.2e  up_3 %.2e  image %.2e" % (name, blocks["up_0"], blocks["up_3"], img), flush=True)
