"""Generate tests/golden/ref_variants.npz: ONE full trainer iteration (G step + D step) of the UNMODIFIED reference for
option variants beyond the default configuration, on the same portable weights / batch as
ref_small.npz.  Runs only in the build container (needs /root/reference).

    python oracle/make_golden_variants.py [--second]     # --second: ref_variants2.npz (style / Gram / openEDS losses)
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle.make_golden import SEEDS, SMALL, import_reference, ref_opt, sub  # noqa: E402
from oracle import seg2eye_oracle as O  # noqa: E402

# name -> (extra reference command-line flags, the matching oracle option overrides)
VARIANTS = {
    "spadeinstance": (["--norm_G", "spectralspadeinstance3x3"], dict(norm_G="spectralspadeinstance3x3")),
    "more_upsampling": (["--num_upsampling_layers", "more"], dict(num_upsampling_layers="more")),
    "aggr_max": (["--style_aggr_method", "max"], dict(style_aggr_method="max")),
    "l2_loss": (["--lambda_l2", "15", "--lambda_l1", "0"], dict(lambda_l2=15.0, lambda_l1=0.0)),
    "gan_ls": (["--gan_mode", "ls"], dict(gan_mode="ls")),
    "gan_original": (["--gan_mode", "original"], dict(gan_mode="original")),
    "gan_w": (["--gan_mode", "w"], dict(gan_mode="w")),
    "no_feat": (["--no_ganFeat_loss"], dict(no_ganFeat_loss=True)),
    "no_ttur": (["--no_TTUR"], dict(no_TTUR=True)),
}


# second fixture (ref_variants2.npz): the optional style losses of the authors' own runs (scripts/current_runs_spadestyle.sh:
# L2 15 + style_w 0.5 + style_feat + max aggregation) plus the Gram loss, and the (gradient-free) openEDS loss term
VARIANTS2 = {
    "style_losses": (["--lambda_l2", "15", "--lambda_l1", "0", "--lambda_style_w", "0.5", "--lambda_style_feat", "0.001",
                      "--lambda_gram", "1.0", "--style_aggr_method", "max"],
                     dict(lambda_l2=15.0, lambda_l1=0.0, lambda_style_w=0.5, lambda_style_feat=0.001, lambda_gram=1.0,
                          style_aggr_method="max")),
    "style_mean": (["--lambda_style_w", "0.5", "--lambda_style_feat", "0.01", "--lambda_gram", "10.0"],
                   dict(lambda_style_w=0.5, lambda_style_feat=0.01, lambda_gram=10.0)),
    "openeds": (["--lambda_openeds", "0.5"], dict(lambda_openeds=0.5)),
}


def main(variants=None, fname="ref_variants.npz"):
    variants = variants or VARIANTS
    import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from trainers.pix2pix_trainer import Pix2PixTrainer
    out = {}
    base = ["--ngf", str(SMALL["ngf"]), "--ndf", str(SMALL["ndf"]), "--lambda_l1", str(SMALL["lambda_l1"]), "--batchSize", "2"]
    for name, (flags, over) in variants.items():
        tmp = tempfile.mkdtemp()
        opt = ref_opt(tmp, base + flags)       # later flags win (lambda_l1 override)
        oopt = O.make_opt(**{**SMALL, **over})
        with contextlib.redirect_stdout(io.StringIO()):
            trainer = Pix2PixTrainer(opt)
        m = trainer.pix2pix_model
        shapes = dict(G=O.generator_shapes(oopt), D=O.discriminator_shapes(oopt), E=O.encoder_shapes(oopt))
        for k, net in dict(G=m.netG, D=m.netD, E=m.netE).items():
            ref_sd = net.state_dict()
            assert list(ref_sd.keys()) == list(shapes[k].keys()), (name, k, set(ref_sd) ^ set(shapes[k]))
            net.load_state_dict(O.synth_state(shapes[k], SEEDS[k]))
        batch = O.synth_batch(oopt, 2, SEEDS["batch"])
        data = {k: v.clone() for k, v in batch.items()}
        trainer.run_generator_one_step(data)
        trainer.run_discriminator_one_step(data)
        for k, v in trainer.get_latest_losses(include_log_losses=name in VARIANTS2).items():
            out["%s|loss|%s" % (name, k)] = v.detach().reshape(-1).numpy().astype(np.float64)
        out["%s|generated_sub" % name], out["%s|generated_stat" % name] = sub(trainer.get_latest_generated())
        sdG = m.netG.state_dict()
        out["%s|post_G_conv_img.weight" % name] = sdG["conv_img.weight"].numpy().copy()
        out["%s|post_G_up_1.conv_0.weight_u" % name] = sdG["up_1.conv_0.weight_u"].numpy().copy()
        out["%s|post_D_model4_bias" % name] = m.netD.state_dict()["discriminator_1.model4.0.bias"].numpy().copy()
        print(name, {k.split("|")[-1]: float(v[0]) for k, v in out.items() if k.startswith(name + "|loss|")}, flush=True)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", fname), **out)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    if "--second" in sys.argv:
        main(VARIANTS2, "ref_variants2.npz")
    else:
        main()
