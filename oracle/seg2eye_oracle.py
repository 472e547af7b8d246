"""CPU oracle for the Seg2Eye SPADE+Style G/D training step.

TEST INFRASTRUCTURE ONLY.  This file is a plain fp32 PyTorch-CPU restatement of the
reference algorithm, written functionally over flat ``state_dict``-style dictionaries.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import it; the product package ``seg2eye_b200`` never does.

Parity pin: the reference ships no tests / golden vectors, so the
oracle is pinned against the *reference itself*, executed in the build container by
``oracle/make_golden.py`` (imports /root/reference with three import shims) which writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Each function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# options (options/base_options.py:21-64, options/train_options.py:25-51)
# --------------------------------------------------------------------------------------
def make_opt(**kw):
    o = dict(
        ngf=64, ndf=64, w_dim=16, input_ns=4, label_nc=4, semantic_nc=4, output_nc=1,
        crop_size=256, aspect_ratio=0.8, num_upsampling_layers="normal",
        norm_G="spectralspadebatch3x3", norm_D="spectralinstance", norm_E="spectralinstance",
        num_D=2, n_layers_D=4, no_ganFeat_loss=False, gan_mode="hinge",
        lambda_feat=10.0, lambda_l1=0.0, lambda_l2=0.0, style_aggr_method="mean",
        lr=2e-4, beta1=0.5, beta2=0.999, no_TTUR=False, weight_decay=0.0,
        niter=14, niter_decay=7, isTrain=True,
    )
    o.update(kw)
    o["semantic_nc"] = o["label_nc"]
    return SimpleNamespace(**o)


def latent_size(opt):
    """generator.py:52-67 -> (sw, sh)."""
    n_up = {"normal": 5, "more": 6, "most": 7}[opt.num_upsampling_layers]
    sw = opt.crop_size // (2 ** n_up)
    sh = round(sw / opt.aspect_ratio)
    return sw, sh


# --------------------------------------------------------------------------------------
# state-dict shapes (SURVEY 8(b) state_dict layout; probed from the reference classes)
# --------------------------------------------------------------------------------------
def _sn_conv_shapes(sd, p, cout, cin, k, bias):
    # spectral_norm re-registers weight as weight_orig after bias; u/v are buffers
    if bias:
        sd[p + ".bias"] = (cout,)
    sd[p + ".weight_orig"] = (cout, cin, k, k)
    sd[p + ".weight_u"] = (cout,)
    sd[p + ".weight_v"] = (cin * k * k,)


def _spade_style_shapes(sd, p, c, opt):
    sd[p + ".spade.param_free_norm.running_mean"] = (c,)
    sd[p + ".spade.param_free_norm.running_var"] = (c,)
    sd[p + ".spade.param_free_norm.num_batches_tracked"] = ()
    sd[p + ".spade.mlp_shared.0.weight"] = (128, opt.semantic_nc, 3, 3)
    sd[p + ".spade.mlp_shared.0.bias"] = (128,)
    for g in ("mlp_gamma", "mlp_beta"):
        sd[p + ".spade.%s.weight" % g] = (c, 128, 3, 3)
        sd[p + ".spade.%s.bias" % g] = (c,)
    if getattr(opt, "netG", "spadestyle") != "spade":     # the style-less SPADE generator (BASELINE config 5) has no ApplyStyle
        sd[p + ".adain.linear.weight"] = (2 * c, opt.w_dim)
        sd[p + ".adain.linear.bias"] = (2 * c,)


def generator_blocks(opt):
    nf = opt.ngf
    return [("head_0", 16 * nf, 16 * nf), ("G_middle_0", 16 * nf, 16 * nf),
            ("G_middle_1", 16 * nf, 16 * nf), ("up_0", 16 * nf, 8 * nf),
            ("up_1", 8 * nf, 4 * nf), ("up_2", 4 * nf, 2 * nf), ("up_3", 2 * nf, nf)]


def generator_shapes(opt):
    """Key order follows module registration order in generator.py:22-50 /
    architecture.py:17-42 / normalization.py:63-89,172-182."""
    sd = {}
    nf = opt.ngf
    instance = "instance" in opt.norm_G
    sd["fc.weight"] = (16 * nf, opt.semantic_nc, 3, 3)
    sd["fc.bias"] = (16 * nf,)
    for name, fin, fout in generator_blocks(opt):
        fmid = min(fin, fout)
        _sn_conv_shapes(sd, name + ".conv_0", fmid, fin, 3, True)
        _sn_conv_shapes(sd, name + ".conv_1", fout, fmid, 3, True)
        if fin != fout:
            _sn_conv_shapes(sd, name + ".conv_s", fout, fin, 1, False)
        norms = [("norm_0", fin), ("norm_1", fmid)] + ([("norm_s", fin)] if fin != fout else [])
        for nn_, c in norms:
            _spade_style_shapes(sd, "%s.%s" % (name, nn_), c, opt)
    sd["conv_img.weight"] = (opt.output_nc, nf, 3, 3)
    sd["conv_img.bias"] = (opt.output_nc,)
    if instance:
        sd = {k: v for k, v in sd.items() if "param_free_norm" not in k}
    return sd


def discriminator_shapes(opt):
    """discriminator.py:70-100."""
    sd = {}
    for i in range(opt.num_D):
        p = "discriminator_%d" % i
        nf = opt.ndf
        sd[p + ".model0.0.weight"] = (nf, opt.label_nc + opt.output_nc, 4, 4)
        sd[p + ".model0.0.bias"] = (nf,)
        for n in range(1, opt.n_layers_D):
            nf_prev, nf = nf, min(nf * 2, 512)
            _sn_conv_shapes(sd, p + ".model%d.0.0" % n, nf, nf_prev, 4, False)
        sd[p + ".model%d.0.weight" % opt.n_layers_D] = (1, nf, 4, 4)
        sd[p + ".model%d.0.bias" % opt.n_layers_D] = (1,)
    return sd


def encoder_shapes(opt):
    """encoder.py:16-51."""
    sd = {}
    ndf = opt.ngf
    chans = [1, ndf, ndf * 2, ndf * 4, ndf * 8, ndf * 8] + ([ndf * 8] if opt.crop_size >= 256 else [])
    for i in range(len(chans) - 1):
        _sn_conv_shapes(sd, "layer%d.0" % i, chans[i + 1], chans[i], 3, False)
    for n in ("fc_mu", "fc_var"):
        sd[n + ".weight"] = (opt.w_dim, ndf * 8 * 16)
        sd[n + ".bias"] = (opt.w_dim,)
    return sd


def synth_state(shapes, seed, scale=None):
    """Portable deterministic weights (numpy PCG64, stable across platforms).

    Scales are O(0.05-1) rather than the reference's xavier(0.02) so that the gamma/beta
    branch is numerically alive (SURVEY section 7 'hard parts').  u/v are unit vectors,
    running_var positive, num_batches_tracked int64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(int(rng.integers(0, 5)), dtype=torch.int64)
            continue
        a = rng.standard_normal(shp).astype(np.float32)
        if k.endswith("weight_u") or k.endswith("weight_v"):
            a = a / max(float(np.linalg.norm(a)), 1e-12)
        elif k.endswith("running_var"):
            a = (0.5 + np.abs(a)).astype(np.float32)
        elif k.endswith("running_mean"):
            a = 0.1 * a
        elif k.endswith("bias"):
            a = 0.1 * a
        elif a.ndim >= 2:
            fan_in = int(np.prod(shp[1:]))
            s = scale if scale is not None else 1.0
            a = a * np.float32(s / math.sqrt(fan_in))
        out[k] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def init_state(shapes, seed):
    """The reference's own initialisation (base_network.py:28-59 with --init_type xavier --init_variance 0.02;
    FC keeps randn * in^-0.5, normalization.py:117-127; BN buffers 0/1; u, v = normalised randn)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(0, dtype=torch.int64)
            continue
        a = rng.standard_normal(shp).astype(np.float32)
        if k.endswith("weight_u") or k.endswith("weight_v"):
            a = a / max(float(np.linalg.norm(a)), 1e-12)
        elif k.endswith("running_var"):
            a = np.ones(shp, np.float32)
        elif k.endswith("running_mean") or k.endswith("bias"):
            a = np.zeros(shp, np.float32)
        elif "adain.linear.weight" in k:
            a = a * np.float32(shp[1] ** -0.5)
        else:
            rf = int(np.prod(shp[2:])) if len(shp) > 2 else 1
            a = a * np.float32(0.02 * math.sqrt(2.0 / (shp[1] * rf + shp[0] * rf)))
        out[k] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out


def synth_batch(opt, batch, seed, hw=None):
    """eye-shaped 4-class label ellipses, images/targets U(-1,1), 4-D label.  For label_nc > 4 (config 5:
    35 classes) the four regions are subdivided into vertical bands so that every class id occurs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if hw is None:
        w = opt.crop_size
        h = round(opt.crop_size / opt.aspect_ratio)
    else:
        h, w = hw
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    label = np.zeros((batch, 1, h, w), dtype=np.uint8)
    for b in range(batch):
        cy = h * (0.4 + 0.2 * rng.random())
        cx = w * (0.4 + 0.2 * rng.random())
        ry = h * (0.15 + 0.1 * rng.random())
        rx = w * (0.3 + 0.15 * rng.random())
        r2 = min(ry, rx) * (0.6 + 0.2 * rng.random())
        r3 = r2 * (0.3 + 0.3 * rng.random())
        d_scl = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        d_cir = (yy - cy) ** 2 + (xx - cx) ** 2
        lab = np.zeros((h, w), dtype=np.uint8)
        lab[d_scl <= 1.0] = 1
        lab[(d_cir <= r2 * r2) & (d_scl <= 1.0)] = 2
        lab[(d_cir <= r3 * r3) & (d_scl <= 1.0)] = 3
        if opt.label_nc > 4:
            bands = (xx.astype(np.int64) * ((opt.label_nc + 3) // 4) // w).astype(np.uint8)
            lab = np.minimum(lab + 4 * bands, opt.label_nc - 1).astype(np.uint8)
        label[b, 0] = lab
    style = rng.uniform(-1, 1, size=(batch, opt.input_ns, 1, h, w)).astype(np.float32)
    target = rng.uniform(-1, 1, size=(batch, 1, h, w)).astype(np.float32)
    return {"label": torch.from_numpy(label), "style_image": torch.from_numpy(style),
            "target": torch.from_numpy(target)}


# --------------------------------------------------------------------------------------
# integer path: one-hot and nearest indices (bit-exact)
# --------------------------------------------------------------------------------------
def one_hot(label, nc):
    """pix2pix_model.py:138-152: zeros(B,nc,H,W).scatter_(1, label.long(), 1.0)."""
    lab = label.long()
    if lab.dim() == 3:
        lab = lab.unsqueeze(0)
    b, _, h, w = lab.shape
    return torch.zeros(b, nc, h, w, device=lab.device).scatter_(1, lab, 1.0)


def nearest_src_index(dst_len, src_len):
    """ATen 'nearest' (legacy) index: floor(dst * (src/dst)) computed in fp32, clamped.
    Used by F.interpolate at normalization.py:97 and generator.py:72."""
    scale = np.float32(src_len) / np.float32(dst_len)
    idx = np.floor(np.arange(dst_len, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, src_len - 1)


def nearest_resize(x, size):
    hi = torch.from_numpy(nearest_src_index(size[0], x.shape[2])).to(x.device)
    wi = torch.from_numpy(nearest_src_index(size[1], x.shape[3])).to(x.device)
    return x[:, :, hi][:, :, :, wi]


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def spectral_weight(sd, p, training=True):
    """torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:62-117) as applied at
    normalization.py:26 and architecture.py:31-34: one power iteration per forward in
    training mode, in-place on the u/v buffers, sigma = u.(W v) with u,v constants."""
    w = sd[p + ".weight_orig"]
    u, v = sd[p + ".weight_u"], sd[p + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
    uc, vc = u.clone(), v.clone()
    sigma = torch.dot(uc, torch.mv(wm, vc))
    return w / sigma


def batch_norm_train(x, sd, p):
    """nn.BatchNorm2d(affine=False) in training mode (normalization.py:75): biased variance
    for normalisation, unbiased for running_var, momentum 0.1, eps 1e-5."""
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    with torch.no_grad():
        sd[p + ".running_mean"].mul_(0.9).add_(0.1 * mean.detach())
        sd[p + ".running_var"].mul_(0.9).add_(0.1 * var.detach() * (n / max(n - 1, 1)))
        sd[p + ".num_batches_tracked"].add_(1)
    return (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + 1e-5)


def instance_norm(x):
    """nn.InstanceNorm2d(affine=False, eps=1e-5) (normalization.py:41,73)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + 1e-5)


def batch_norm_eval(x, sd, p):
    """nn.BatchNorm2d(affine=False) in eval mode (test.py: model.eval()): running statistics, nothing is updated."""
    m, v = sd[p + ".running_mean"], sd[p + ".running_var"]
    return (x - m[None, :, None, None]) * torch.rsqrt(v[None, :, None, None] + 1e-5)


def spade(sd, p, x, seg, instance=False, training=True):
    """SPADE.forward, normalization.py:91-105."""
    if instance:
        normalized = instance_norm(x)
    elif training:
        normalized = batch_norm_train(x, sd, p + ".param_free_norm")
    else:
        normalized = batch_norm_eval(x, sd, p + ".param_free_norm")
    seg_r = nearest_resize(seg, x.shape[2:])
    actv = F.relu(F.conv2d(seg_r, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def apply_style(sd, p, x, w):
    """ApplyStyle.forward + FC.forward, normalization.py:134-169 (w_lrmul = b_lrmul = 1,
    LeakyReLU(0.2) on the style vector, no normalisation of x)."""
    style = F.leaky_relu(F.linear(w, sd[p + ".linear.weight"], sd[p + ".linear.bias"]), 0.2)
    style = style.view(-1, 2, x.shape[1], 1, 1)
    return x * (style[:, 0] + 1.0) + style[:, 1]


def spade_style_block(sd, p, x, seg, w, opt, training=True):
    """SPADE_STYLE_Block.forward, normalization.py:184-192.  w None: the block of the ORIGINAL SPADE generator
    (BASELINE config 5; SURVEY 8(c) last bullet) -- SPADE.forward (normalization.py:91-105) alone, i.e. the same block
    with the ApplyStyle term and the division by two removed."""
    s = spade(sd, p + ".spade", x, seg, instance="instance" in opt.norm_G, training=training)
    if w is None:
        return s
    a = apply_style(sd, p + ".adain", x, w)
    return (s + a) / 2


def resblock(sd, p, x, seg, w, fin, fout, opt, training=True, taps=None):
    """SPADE_STYLE_ResnetBlock.forward, architecture.py:44-62.  Evaluation order matters for
    the BN buffers: shortcut (norm_s) first, then norm_0, then norm_1."""
    sn = "spectral" in opt.norm_G

    def conv(name, t, pad):
        if sn:
            wt = spectral_weight(sd, "%s.%s" % (p, name), training)
        else:
            wt = sd["%s.%s.weight" % (p, name)]
        return F.conv2d(t, wt, sd.get("%s.%s.bias" % (p, name)), padding=pad)

    if fin != fout:
        ns = spade_style_block(sd, p + ".norm_s", x, seg, w, opt, training)
        x_s = conv("conv_s", ns, 0)
    else:
        x_s = x
    n0 = F.leaky_relu(spade_style_block(sd, p + ".norm_0", x, seg, w, opt, training), 0.2)
    dx = conv("conv_0", n0, 1)
    n1 = F.leaky_relu(spade_style_block(sd, p + ".norm_1", dx, seg, w, opt, training), 0.2)
    dx = conv("conv_1", n1, 1)
    out = x_s + dx
    if taps is not None:
        taps[p + ".norm_0"] = n0
        taps[p] = out
    return out


def generator_forward(sd, seg, w, opt, training=True, taps=None):
    """SPADESTYLEGenerator.forward, generator.py:69-102 ('normal'/'more' upsampling)."""
    sw, sh = latent_size(opt)
    x = nearest_resize(seg, (sh, sw))
    x = F.conv2d(x, sd["fc.weight"], sd["fc.bias"], padding=1)
    if taps is not None:
        taps["fc"] = x
    blocks = generator_blocks(opt)

    def up(t):  # nn.Upsample(scale_factor=2), nearest
        return t.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    x = resblock(sd, blocks[0][0], x, seg, w, blocks[0][1], blocks[0][2], opt, training, taps)
    x = up(x)
    x = resblock(sd, blocks[1][0], x, seg, w, blocks[1][1], blocks[1][2], opt, training, taps)
    if opt.num_upsampling_layers in ("more", "most"):
        x = up(x)
    x = resblock(sd, blocks[2][0], x, seg, w, blocks[2][1], blocks[2][2], opt, training, taps)
    for name, fin, fout in blocks[3:]:
        x = up(x)
        x = resblock(sd, name, x, seg, w, fin, fout, opt, training, taps)
    x = F.conv2d(F.leaky_relu(x, 0.2), sd["conv_img.weight"], sd["conv_img.bias"], padding=1)
    return torch.tanh(x)


def encoder_forward(sd, x, opt, training=True):
    """ConvEncoder.forward, encoder.py:53-73: bilinear to 256x256 (align_corners=False),
    [SN conv3x3 s2 p1 (no bias) -> InstanceNorm] x n with no activation in between,
    LeakyReLU(0.2), flatten, fc_mu / fc_var."""
    if x.shape[2] != 256 or x.shape[3] != 256:
        x = F.interpolate(x, size=(256, 256), mode="bilinear", align_corners=False)
    feats = []
    n_layers = 6 if opt.crop_size >= 256 else 5
    for i in range(n_layers):
        wt = spectral_weight(sd, "layer%d.0" % i, training)
        x = instance_norm(F.conv2d(x, wt, None, stride=2, padding=1))
        feats.append(x)
    out = F.leaky_relu(x, 0.2).reshape(x.shape[0], -1)
    mu = F.linear(out, sd["fc_mu.weight"], sd["fc_mu.bias"])
    logvar = F.linear(out, sd["fc_var.weight"], sd["fc_var.bias"])
    return mu, logvar, feats


def aggregate(t, dim, opt):
    """Pix2PixModel._aggregate_tensor, pix2pix_model.py:271-278."""
    if opt.style_aggr_method == "mean":
        return t.mean(dim=dim)
    return t.max(dim=dim).values


def encode_w(sdE, style_image, opt, training=True, with_features=False):
    """pix2pix_model.py:280-314: one netE call per batch sample over its ns style images,
    mu stacked to (B, ns, w_dim) and aggregated (mean | max) over ns.  with_features: also the per-sample lists of
    feature maps aggregated over the ns images (pix2pix_model.py:297-302)."""
    assert style_image.dim() == 5
    res = [encoder_forward(sdE, style_image[b], opt, training) for b in range(style_image.shape[0])]
    w = aggregate(torch.stack([r[0] for r in res], dim=0), 1, opt)
    if not with_features:
        return w
    return w, [[aggregate(f, 0, opt) for f in r[2]] for r in res]


def gram_matrix(x):
    """loss.py:177-189."""
    a, b, c, d = x.shape
    f = x.reshape(a * b, c * d)
    return torch.mm(f, f.t()).div(a * b * c * d)


def style_feature_loss(feats_fake, feats_real, kind):
    """_compute_style_feature_loss / _compute_gram_loss, pix2pix_model.py:162-184: per encoder level, the per-sample
    aggregated maps stacked over the batch, MSE (kind 'feat': both sides live) or StyleLoss (kind 'gram': loss.py:192-200,
    Gram matrix of the real side detached), summed over the levels."""
    total = []
    for lvl in range(len(feats_fake[0])):
        ff = torch.stack([f[lvl] for f in feats_fake])
        fr = torch.stack([f[lvl] for f in feats_real])
        if kind == "feat":
            total.append(F.mse_loss(ff, fr))
        else:
            total.append(F.mse_loss(gram_matrix(ff), gram_matrix(fr).detach()))
    return torch.sum(torch.stack(total))


def to_255(image):
    """ImageProcessor.to_255imagebatch on a tensor in [-1,1] (postprocessor.py:57-72,92-96): add 1, mul 255, div 2 in the
    tensor's own dtype, then .int() (truncation)."""
    if image.dim() == 3:
        image = image.unsqueeze(0)
    return torch.div(torch.mul(torch.add(image, 1), 255), 2).int()


def openeds_accuracy(produced, target):
    """openEDSaccuracy, loss.py:102-111 (one image)."""
    diff = produced.float() - target.float()
    h, w = diff.shape[-2:]
    return torch.sqrt(torch.sum(diff * diff).float()) / (h * w)


def mse_for_images(produced, target):
    """MSECalculator.calculate_mse_for_images / _for_tensors, loss.py:113-157 (after the 0..255 conversion)."""
    return torch.stack([openeds_accuracy(produced[i], target[i]) for i in range(produced.shape[0])])


def _fma(a, b, c):
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def cv_linear_table(dst, src):
    """Coefficient table of cv2.resize(INTER_LINEAR) for one axis (modules/imgproc/src/resize.cpp): source index and
    fractional part of fx = (d + 0.5) * scale - 0.5 with scale = 1 / (dst / src).  The OpenCV build the reference runs
    on (4.x wheels, AVX2 dispatch) contracts the multiply-add, so the product is not rounded before the subtraction."""
    scale = 1.0 / (float(dst) / float(src))
    idx, frac = np.empty(dst, np.int64), np.empty(dst, np.float64)
    for d in range(dst):
        c = _fma(d + 0.5, scale, -0.5)
        idx[d] = math.floor(c)
        frac[d] = c - math.floor(c)
    return idx, frac


def resize_linear_f64(img, w, h):
    """cv2.resize(img.astype(float64), (w, h), interpolation=cv2.INTER_LINEAR) for one (H, W) image."""
    H, W = img.shape
    sx, fx = cv_linear_table(w, W)
    sy, fy = cv_linear_table(h, H)
    lo, hi = sx < 0, sx >= W - 1
    fx[lo], sx[lo] = 0.0, 0
    fx[hi], sx[hi] = 0.0, W - 1
    s = img.astype(np.float64)
    sx1 = np.minimum(sx + 1, W - 1)
    rows = s[:, sx] * (1.0 - fx)[None] + s[:, sx1] * fx[None]
    single = sx + 1 >= W
    rows[:, single] = s[:, sx[single]]
    r0, r1 = np.clip(sy, 0, H - 1), np.clip(sy + 1, 0, H - 1)
    return rows[r0] * (1.0 - fy)[:, None] + rows[r1] * fy[:, None]


def to_255_resized(fake, w=400, h=640):
    """ImageProcessor.to_255resized_imagebatch (postprocessor.py:98-114), the tail of util/tester.py:44-47: per image
    cv2.resize in float64, then the 0..255 mapping in float64 and .int().  (B,1,h0,w0) fp32 -> (B,1,h,w) int32."""
    out = np.stack([resize_linear_f64(im[0], w, h)[None] for im in fake.detach().cpu().numpy()])
    return to_255(torch.from_numpy(out))


def nlayer_discriminator(sd, p, x, opt, training=True):
    """NLayerDiscriminator.forward, discriminator.py:78-116: 4x4 convs, padding 2."""
    outs = []
    x = F.leaky_relu(F.conv2d(x, sd[p + ".model0.0.weight"], sd[p + ".model0.0.bias"], stride=2, padding=2), 0.2)
    outs.append(x)
    for n in range(1, opt.n_layers_D):
        stride = 1 if n == opt.n_layers_D - 1 else 2
        wt = spectral_weight(sd, p + ".model%d.0.0" % n, training)
        x = F.leaky_relu(instance_norm(F.conv2d(x, wt, None, stride=stride, padding=2)), 0.2)
        outs.append(x)
    n = opt.n_layers_D
    x = F.conv2d(x, sd[p + ".model%d.0.weight" % n], sd[p + ".model%d.0.bias" % n], stride=1, padding=2)
    outs.append(x)
    return outs


def discriminator_forward(sd, x, opt, training=True):
    """MultiscaleDiscriminator.forward, discriminator.py:46-63."""
    result = []
    for i in range(opt.num_D):
        result.append(nlayer_discriminator(sd, "discriminator_%d" % i, x, opt, training))
        x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    return result


def hinge(pred, target_is_real, for_discriminator):
    """GANLoss.loss, loss.py:66-77 (hinge)."""
    if for_discriminator:
        if target_is_real:
            return -torch.mean(torch.clamp(pred - 1, max=0))
        return -torch.mean(torch.clamp(-pred - 1, max=0))
    return -torch.mean(pred)


def gan_term(pred, target_is_real, for_discriminator, mode="hinge", real_label=1.0, fake_label=0.0):
    """GANLoss.loss, loss.py:58-83: 'original' = BCE with logits against the label, 'ls' = MSE against the label,
    'hinge', anything else = WGAN."""
    label = real_label if target_is_real else fake_label
    if mode == "original":
        return F.binary_cross_entropy_with_logits(pred, torch.full_like(pred, label))
    if mode == "ls":
        return F.mse_loss(pred, torch.full_like(pred, label))
    if mode == "hinge":
        return hinge(pred, target_is_real, for_discriminator)
    return -pred.mean() if target_is_real else pred.mean()


def gan_loss(preds, target_is_real, for_discriminator, mode="hinge"):
    """GANLoss.__call__, loss.py:85-99: last tensor of each D, averaged over num_D, shape (1,)."""
    loss = 0
    for p in preds:
        loss = loss + gan_term(p[-1], target_is_real, for_discriminator, mode).view(1)
    return loss / len(preds)


def discriminate(sdD, seg, fake, real, opt, training=True):
    """pix2pix_model.py:328-358."""
    both = torch.cat([torch.cat([seg, fake], 1), torch.cat([seg, real], 1)], 0)
    out = discriminator_forward(sdD, both, opt, training)
    nb = both.shape[0] // 2
    return [[t[:nb] for t in d] for d in out], [[t[nb:] for t in d] for d in out]


def generator_losses(sdG, sdD, sdE, batch, opt):
    """compute_generator_loss, pix2pix_model.py:186-247 (GAN, L1 / L2, openEDS, style_w / style_feat / gram, GAN_Feat)."""
    seg = one_hot(batch["label"], opt.label_nc)
    if getattr(opt, "netG", "spadestyle") == "spade":
        w, feats_real = None, []
    else:
        w, feats_real = encode_w(sdE, batch["style_image"], opt, with_features=True)
    fake = generator_forward(sdG, seg, w, opt)
    pred_fake, pred_real = discriminate(sdD, seg, fake, batch["target"], opt)
    losses = {"GAN": gan_loss(pred_fake, True, False, getattr(opt, "gan_mode", "hinge"))}
    if opt.lambda_l2:
        losses["L2/weighted"] = F.mse_loss(fake, batch["target"]) * opt.lambda_l2
    if opt.lambda_l1:
        losses["L1/weighted"] = F.l1_loss(fake, batch["target"]) * opt.lambda_l1
    if getattr(opt, "lambda_openeds", 0):     # pix2pix_model.py:209-213: the .int() inside makes it a constant
        losses["openeds/weighted"] = mse_for_images(to_255(fake), to_255(batch["target"])) * opt.lambda_openeds
    lw, lf, lg = (getattr(opt, k, 0) for k in ("lambda_style_w", "lambda_style_feat", "lambda_gram"))
    if lw or lf or lg:                         # pix2pix_model.py:215-231
        w_fake, feats_fake = encode_w(sdE, fake.unsqueeze(1), opt, with_features=True)
        if lw > 0:
            losses["style_w/weighted"] = F.mse_loss(w_fake, w) * lw
        if lf > 0:
            losses["style_feat/weighted"] = style_feature_loss(feats_fake, feats_real, "feat") * lf
        if lg > 0:
            losses["gram/weighted"] = style_feature_loss(feats_fake, feats_real, "gram") * lg
    if not opt.no_ganFeat_loss:
        fm = torch.zeros(1, device=fake.device)
        for i in range(len(pred_fake)):
            for j in range(len(pred_fake[i]) - 1):
                fm = fm + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) * opt.lambda_feat / len(pred_fake)
        losses["GAN_Feat"] = fm
    return losses, fake


def discriminator_losses(sdG, sdD, sdE, batch, opt):
    """compute_discriminator_loss, pix2pix_model.py:249-264."""
    seg = one_hot(batch["label"], opt.label_nc)
    with torch.no_grad():
        w = None if getattr(opt, "netG", "spadestyle") == "spade" else encode_w(sdE, batch["style_image"], opt)
        fake = generator_forward(sdG, seg, w, opt)
    fake = fake.detach().requires_grad_()
    pred_fake, pred_real = discriminate(sdD, seg, fake, batch["target"], opt)
    mode = getattr(opt, "gan_mode", "hinge")
    return {"D/Fake": gan_loss(pred_fake, False, True, mode), "D/real": gan_loss(pred_real, True, True, mode)}


# --------------------------------------------------------------------------------------
# trainer (trainers/pix2pix_trainer.py:26-45, pix2pix_model.py:92-110)
# --------------------------------------------------------------------------------------
def _is_param(k):
    return not (k.endswith("weight_u") or k.endswith("weight_v") or "running_" in k or k.endswith("num_batches_tracked"))


class OracleTrainer:
    """Adam(betas (0,0.9) under TTUR, lr/2 for G+E and lr*2 for D), one G step then one D step."""

    def __init__(self, sdG, sdD, sdE, opt):
        self.opt = opt
        self.sdG, self.sdD, self.sdE = sdG, sdD, sdE
        for sd in (sdG, sdD, sdE):
            for k, v in sd.items():
                if _is_param(k):
                    v.requires_grad_(True)
        if opt.no_TTUR:
            b1, b2, glr, dlr = opt.beta1, opt.beta2, opt.lr, opt.lr
        else:
            b1, b2, glr, dlr = 0.0, 0.9, opt.lr / 2, opt.lr * 2
        gp = [v for sd in (sdG, sdE) for k, v in sd.items() if _is_param(k)]
        dp = [v for k, v in sdD.items() if _is_param(k)]
        self.opt_G = torch.optim.Adam(gp, lr=glr, betas=(float(b1), float(b2)), weight_decay=opt.weight_decay)
        self.opt_D = torch.optim.Adam(dp, lr=dlr, betas=(float(b1), float(b2)), weight_decay=opt.weight_decay)

    def run_generator_one_step(self, batch):
        self.opt_G.zero_grad()
        self.g_losses, self.generated = generator_losses(self.sdG, self.sdD, self.sdE, batch, self.opt)
        sum(self.g_losses.values()).mean().backward()
        self.opt_G.step()

    def run_discriminator_one_step(self, batch):
        self.opt_D.zero_grad()
        self.d_losses = discriminator_losses(self.sdG, self.sdD, self.sdE, batch, self.opt)
        sum(self.d_losses.values()).mean().backward()
        self.opt_D.step()


# --------------------------------------------------------------------------------------
# data layer (data/base_dataset.py:50-80, data/openeds_dataset.py:82-119), preprocess_mode 'fixed'
# --------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2     # Pillow, src/libImaging/Resample.c


def _pil_bicubic(x):
    """Pillow's bicubic kernel (a = -0.5), Resample.c bicubic_filter."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_table(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for Image.resize(..., BICUBIC) on 8-bit images: per output
    index the first source index, the tap count and the taps as 22-bit fixed-point integers.  (The reference resizes
    every style / target image with PIL, base_dataset.py:88-91 via :69-72.)"""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_pil_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PIL_PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _pil_resample_axis(img, kk, bounds, axis):
    a = img.astype(np.int64)
    if axis == 0:
        a = a.T
    out = np.zeros((a.shape[0], kk.shape[0]), np.int64)
    for xx in range(kk.shape[0]):
        xmin, n = bounds[xx]
        out[:, xx] = (1 << (PIL_PRECISION_BITS - 1)) + (a[:, xmin:xmin + n] * kk[xx, :n].astype(np.int64)[None]).sum(1)
    out = np.clip(out >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
    return out.T if axis == 0 else out


def pil_resize_bicubic(img, w, h):
    """Image.fromarray(img, 'L').resize((w, h), Image.BICUBIC) for a uint8 (H, W) array: horizontal pass, then vertical,
    8-bit intermediate (Resample.c ImagingResampleInner)."""
    out = img
    if img.shape[1] != w:
        out = _pil_resample_axis(out, *pil_bicubic_table(img.shape[1], w), 1)
    if img.shape[0] != h:
        out = _pil_resample_axis(out, *pil_bicubic_table(img.shape[0], h), 0)
    return out


def cv_nearest_index(dst, src):
    """cv2.resize(INTER_NEAREST): source index min(floor(d * (1 / (dst / src))), src - 1) (resize.cpp resizeNN)."""
    ifx = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst) * ifx).astype(np.int64), src - 1)


def preprocess_sample(mask, images, w, h, flip):
    """OpenEDSDataset.__getitem__ (openeds_dataset.py:82-119) with get_transform in 'fixed' mode (base_dataset.py:69-80):
    mask uint8 (H0, W0) -> cv2 nearest -> flip -> integer label (h, w); every image uint8 (H0, W0) -> PIL bicubic -> flip ->
    ToTensor (/255) -> Normalize(0.5, 0.5) -> fp32 (1, h, w)."""
    lab = mask[cv_nearest_index(h, mask.shape[0])][:, cv_nearest_index(w, mask.shape[1])]
    if flip:
        lab = lab[:, ::-1]
    outs = []
    for im in images:
        r = pil_resize_bicubic(im, w, h)
        if flip:
            r = r[:, ::-1]
        t = torch.from_numpy(np.ascontiguousarray(r)).to(torch.float32).div(255)
        outs.append(t.sub(0.5).div(0.5).unsqueeze(0))
    return torch.from_numpy(np.ascontiguousarray(lab)), outs
