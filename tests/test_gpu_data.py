"""Device-side data layer against the reference's own preprocessing: every label and every float bit of
seg2eye_b200.data.DevicePreprocessor must equal what data/base_dataset.py:get_transform produced for the same raw frames
(tests/golden/ref_data.npz, SHA-256 over the full results) and the oracle's restatement."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import seg2eye_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_device_preprocessing_bit_exact_vs_reference_fixture():
    from oracle.make_golden_data import CASES, data_inputs, digest
    from seg2eye_b200.data import DevicePreprocessor
    ref = np.load(os.path.join(GOLD, "ref_data.npz"))
    for name, (crop, ar, flip) in CASES.items():
        mask, images = data_inputs(name)
        prep = DevicePreprocessor(SimpleNamespace(crop_size=crop, aspect_ratio=ar, isTrain=True, no_flip=False, preprocess_mode="fixed"))
        raw = {"label": torch.from_numpy(mask)[None], "style_image": torch.from_numpy(images[:2])[None],
               "target": torch.from_numpy(images[2])[None], "filename": ["x"]}
        out = prep(raw, flip=[flip])
        w, h = crop, round(crop / ar)
        assert out["label"].shape == (1, 1, h, w) and out["label"].dtype == torch.int64
        assert out["style_image"].shape == (1, 2, 1, h, w) and out["target"].shape == (1, 1, h, w)
        assert out["target_original"].shape == (1, 1, 640, 400) and out["target_original"].dtype == torch.int32 and out["filename"] == ["x"]
        lab = out["label"][0, 0].cpu().numpy()
        assert np.array_equal(digest(lab.astype(np.uint8)), ref[name + "|label_sha"]), name
        ims = torch.cat([out["style_image"][0], out["target"]], 0).cpu()          # (3,1,h,w), the fixture's order
        assert np.array_equal(ims.numpy().reshape(-1)[::997], ref[name + "|images_sub"]), name
        assert np.array_equal(digest(ims.numpy().astype(np.float32)), ref[name + "|images_sha"]), name
        lab_o, ims_o = O.preprocess_sample(mask, images, w, h, flip)
        assert torch.equal(out["label"][0, 0].cpu(), lab_o.long()) and torch.equal(ims, torch.stack(ims_o))
        want = images[2][:, ::-1] if flip else images[2]
        assert np.array_equal(out["target_original"][0, 0].cpu().numpy(), want.astype(np.int32))


def test_preprocessed_batch_feeds_the_trainer():
    """A batch of raw frames -> DevicePreprocessor -> one trainer iteration: the dict has the reference's keys / shapes / dtypes
    (label (B,1,h,w) integer, style_image (B,ns,1,h,w), target (B,1,h,w)) and per-sample flips are applied consistently."""
    from seg2eye_b200.data import DevicePreprocessor
    from seg2eye_b200.trainers.pix2pix_trainer import Pix2PixTrainer
    oopt = O.make_opt(ngf=16, ndf=16, lambda_l1=10.0)
    d = vars(oopt).copy()
    d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer", continue_train=False, which_epoch="latest",
             checkpoints_dir="/tmp/s2e_ckpt", name="td", no_vgg_loss=True, lambda_openeds=0.0, lambda_style_w=0.0, lambda_style_feat=0.0,
             lambda_gram=0.0, netG="spadestyle", netD="multiscale", preprocess_mode="fixed", no_flip=False)
    opt = SimpleNamespace(**d)
    rng = np.random.Generator(np.random.PCG64(5))
    raw = {"label": torch.from_numpy(rng.integers(0, 4, size=(2, 640, 400)).astype(np.uint8)),
           "style_image": torch.from_numpy(rng.integers(0, 256, size=(2, 4, 640, 400)).astype(np.uint8)),
           "target": torch.from_numpy(rng.integers(0, 256, size=(2, 640, 400)).astype(np.uint8))}
    prep = DevicePreprocessor(opt)
    a, b = prep(raw, flip=[False, True]), prep(raw, flip=[False, False])
    assert torch.equal(a["label"][0], b["label"][0]) and torch.equal(a["label"][1], b["label"][1].flip(-1))
    assert torch.equal(a["style_image"][1], b["style_image"][1].flip(-1)) and torch.equal(a["target"][0], b["target"][0])
    assert float(a["target"].min()) >= -1 and float(a["target"].max()) <= 1
    tr = Pix2PixTrainer(opt)
    tr.run_generator_one_step(dict(a))
    tr.run_discriminator_one_step(dict(a))
    losses = tr.get_latest_losses()
    assert set(losses) == {"GAN", "L1/weighted", "GAN_Feat", "D/Fake", "D/real"} and all(torch.isfinite(v).all() for v in losses.values())
