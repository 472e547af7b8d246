"""Pix2PixModel mirror (reference models/pix2pix_model.py:13-374): same constructor, forward(data, mode)
contract, loss dictionary shapes, optimizers and checkpoint layout; the arithmetic runs in seg2eye_b200 kernels."""
import torch
from torch import nn

from .. import _lib as L
from .. import ops, optim, util
from . import networks


class Pix2PixModel(torch.nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        networks.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.netG, self.netD, self.netE = self.initialize_networks(opt)
        if opt.isTrain:
            self.criterionGAN = networks.GANLoss(opt.gan_mode, opt=self.opt)
            self.criterionFeat = networks.l1_loss
            self.criterionL1 = networks.l1_loss
            self.criterionL2 = networks.mse_loss
            if not getattr(opt, 'no_vgg_loss', True):
                raise ValueError('VGG loss does not exist in the reference either (networks.VGGLoss is undefined)')
            if getattr(opt, 'lambda_openeds', 0):
                raise ValueError('lambda_openeds carries no gradient in the reference (postprocessor.py:72 .int()); unsupported')
            if getattr(opt, 'lambda_style_feat', 0) > 0:
                self.criterion_style_feat = nn.MSELoss()
            if getattr(opt, 'lambda_style_w', 0) > 0:
                self.criterion_style_w = nn.MSELoss()
            if getattr(opt, 'lambda_gram', 0) > 0:
                self.criterion_gram = networks.StyleLoss()
            self.reset_loss_log()

    # ---- loss log (pix2pix_model.py:49-59)
    def get_loss_log(self):
        return {k: torch.mean(torch.stack(v)) for k, v in self.loss_log.items() if len(v)}

    def add_to_loss_log(self, key, value):
        self.loss_log.setdefault(key, []).append(value)

    def reset_loss_log(self):
        self.loss_log = {}

    # ---- entry point (pix2pix_model.py:65-90)
    def forward(self, data, mode):
        input_semantics, style_image, target_image = self.preprocess_input(data)
        if mode == 'generator':
            return self.compute_generator_loss(input_semantics, style_image, target_image)
        elif mode == 'discriminator':
            return self.compute_discriminator_loss(input_semantics, style_image, target_image)
        elif mode == 'encode_only':
            w, features = self.encode_w(style_image)
            return w
        elif mode == 'inference':
            with torch.no_grad():
                if 'latent_style' in data:
                    fake_image = self.generate_fake_from_stylecode(input_semantics, data['latent_style'].to(input_semantics.device))
                else:
                    fake_image, _, _ = self.generate_fake(input_semantics, style_image)
                self.reset_loss_log()
            return fake_image
        else:
            raise ValueError("|mode| is invalid")

    def create_optimizers(self, opt):
        G_params = list(self.netG.parameters()) + list(self.netE.parameters())
        D_params = list(self.netD.parameters()) if opt.isTrain else []
        if opt.no_TTUR:
            beta1, beta2 = opt.beta1, opt.beta2
            G_lr, D_lr = opt.lr, opt.lr
        else:
            beta1, beta2 = 0, 0.9
            G_lr, D_lr = opt.lr / 2, opt.lr * 2
        wd = getattr(opt, 'weight_decay', 0.0)
        optimizer_G = optim.Adam(G_params, lr=G_lr, betas=(beta1, beta2), weight_decay=wd)
        optimizer_D = optim.Adam(D_params, lr=D_lr, betas=(beta1, beta2), weight_decay=wd)
        return optimizer_G, optimizer_D

    def save(self, epoch):
        util.save_network(self.netG, 'G', epoch, self.opt)
        util.save_network(self.netD, 'D', epoch, self.opt)
        util.save_network(self.netE, 'E', epoch, self.opt)

    # ---- helpers
    def initialize_networks(self, opt):
        netG = networks.define_G(opt)
        netD = networks.define_D(opt) if opt.isTrain else None
        netE = networks.define_E(opt)
        if not opt.isTrain or opt.continue_train:
            netG = util.load_network(netG, 'G', opt.which_epoch, opt)
            if opt.isTrain:
                netD = util.load_network(netD, 'D', opt.which_epoch, opt)
                netE = util.load_network(netE, 'E', opt.which_epoch, opt)
        return netG, netD, netE

    def device(self):
        return next(self.netG.parameters()).device

    def preprocess_input(self, data):
        """pix2pix_model.py:138-160: label -> long -> device -> one-hot; mutates `data` in place."""
        dev = self.device()
        data['label'] = data['label'].long().to(dev, non_blocking=True)
        data['style_image'] = data['style_image'].to(dev, non_blocking=True)
        label_map = data['label']
        if label_map.dim() == 3:
            # the reference unsqueezes dim 0 here, which is only right for batch size 1
            label_map = label_map.unsqueeze(0) if label_map.shape[0] == 1 else label_map.unsqueeze(1)
        input_semantics = ops.one_hot(label_map, self.opt.label_nc)
        if "target" in data:
            data['target'] = data['target'].to(dev, non_blocking=True)
            return input_semantics, data['style_image'], data['target']
        return input_semantics, data['style_image'], None

    def compute_generator_loss(self, input_semantics, style_image, target_image):
        G_losses = {}
        fake_image, latent_style_real, style_features_real = self.generate_fake(input_semantics, style_image)
        # D's parameter gradients from this step are discarded by the reference (zero_grad before the D step)
        with ops.skip_weight_grads():
            pred_fake, pred_real = self.discriminate(input_semantics, fake_image, target_image)
        d_full = getattr(self, '_last_d_out', None)   # un-split [fake ; real] features of the same call
        self._last_d_out = None
        G_losses['GAN'] = self.criterionGAN(pred_fake, True, for_discriminator=False)
        if self.opt.lambda_l2:
            l2_loss = self.criterionL2(fake_image, target_image)
            G_losses['L2/weighted'] = l2_loss * self.opt.lambda_l2
            self.add_to_loss_log('L2/raw', l2_loss.detach())
        if self.opt.lambda_l1:
            l1_loss = self.criterionL1(fake_image, target_image)
            G_losses['L1/weighted'] = l1_loss * self.opt.lambda_l1
            self.add_to_loss_log('L1/raw', l1_loss.detach())
        if getattr(self.opt, 'lambda_style_feat', 0) or getattr(self.opt, 'lambda_style_w', 0) or getattr(self.opt, 'lambda_gram', 0):
            latent_style_fake, style_features_fake = self.encode_w(fake_image.unsqueeze(1))
            if self.opt.lambda_style_w > 0:
                raw = self.criterion_style_w(latent_style_fake, latent_style_real)
                G_losses['style_w/weighted'] = raw * self.opt.lambda_style_w
                self.add_to_loss_log('style_w/raw', raw.detach())
            if self.opt.lambda_style_feat > 0:
                raw = self._compute_style_feature_loss(style_features_fake, style_features_real)
                G_losses['style_feat/weighted'] = raw * self.opt.lambda_style_feat
                self.add_to_loss_log('style_feat/raw', raw.detach())
            if self.opt.lambda_gram > 0:
                raw = self._compute_gram_loss(style_features_fake, style_features_real)
                G_losses['gram/weighted'] = raw * self.opt.lambda_gram
                self.add_to_loss_log('gram/raw', raw.detach())
        if not self.opt.no_ganFeat_loss:
            num_D = len(pred_fake)
            GAN_Feat_loss = torch.zeros(1, device=fake_image.device)
            for i in range(num_D):
                for j in range(len(pred_fake[i]) - 1):
                    if d_full is not None:
                        # nn.L1Loss(pred_fake, pred_real.detach()) evaluated on the un-split tensor (no slice copies)
                        t = networks.loss._flat(d_full[i][j])
                        unweighted = ops.HalvesLossFn.apply(t, L.RED_L1, 2.0 / t.numel()).view(())
                    else:
                        unweighted = self.criterionFeat(pred_fake[i][j], pred_real[i][j].detach())
                    GAN_Feat_loss = GAN_Feat_loss + unweighted * self.opt.lambda_feat / num_D
            G_losses['GAN_Feat'] = GAN_Feat_loss
        return G_losses, fake_image

    def compute_discriminator_loss(self, input_semantics, real_image, target_image):
        D_losses = {}
        with torch.no_grad():
            fake_image, _, _ = self.generate_fake(input_semantics, real_image)
            fake_image = fake_image.detach()
        # the reference marks fake_image as requiring grad (pix2pix_model.py:254) but never reads that gradient
        pred_fake, pred_real = self.discriminate(input_semantics, fake_image, target_image)
        self._last_d_out = None   # never keep an autograd graph alive across steps (CUDA-graph capture needs that)
        D_losses['D/Fake'] = self.criterionGAN(pred_fake, False, for_discriminator=True)
        D_losses['D/real'] = self.criterionGAN(pred_real, True, for_discriminator=True)
        return D_losses

    def _feat_stack_loss(self, crit, features_fake, features_real):
        losses = []
        for i in range(len(features_fake[0])):
            ff = torch.stack([f[i].float() for f in features_fake])
            fr = torch.stack([f[i].float() for f in features_real])
            losses.append(crit(ff, fr))
        return torch.sum(torch.stack(losses))

    def _compute_style_feature_loss(self, features_fake, features_real):
        return self._feat_stack_loss(self.criterion_style_feat, features_fake, features_real)

    def _compute_gram_loss(self, features_fake, features_real):
        return self._feat_stack_loss(self.criterion_gram, features_fake, features_real)

    def _aggregate_tensor(self, tensor, dim=1):
        if self.opt.style_aggr_method == 'mean':
            return torch.mean(tensor, dim=dim)
        elif self.opt.style_aggr_method == 'max':
            return torch.max(tensor, dim=dim).values
        raise ValueError(f"Aggregation method not found: {self.opt.style_aggr_method}")

    def _compute_multiple_netE(self, real_image):
        # The reference calls netE once per sample (pix2pix_model.py:285), advancing its spectral-norm vectors once
        # per call.  forward_samples reproduces exactly that in one batched pass.
        B, ns = real_image.shape[0], real_image.shape[1]
        if hasattr(self.netE, 'forward_samples'):
            out, _, feats = self.netE.forward_samples(real_image)
            features = [[f[b * ns:(b + 1) * ns] for f in feats] for b in range(B)]
        else:
            result = [self.netE(real_image[b]) for b in range(B)]
            mu, logvar, features = zip(*result)
            out = torch.stack(mu, dim=0)
        assert out.shape == (*real_image.shape[:2], self.opt.w_dim)
        return out, features

    def _compute_aggregated_w(self, real_image):
        multiple_w, features = self._compute_multiple_netE(real_image)
        w = self._aggregate_tensor(multiple_w)
        need_feats = self.opt.isTrain and (getattr(self.opt, 'lambda_style_feat', 0) or getattr(self.opt, 'lambda_gram', 0))
        features_aggregated = []
        if need_feats:
            for b in range(real_image.shape[0]):
                features_aggregated.append([self._aggregate_tensor(f.float(), dim=0) for f in features[b]])
        assert w.shape == (real_image.shape[0], self.opt.w_dim)
        return w, features_aggregated

    def encode_w(self, real_image):
        if real_image.dim() == 5:
            return self._compute_aggregated_w(real_image)
        raise ValueError("real_image should have 5 dimensions")

    def generate_fake_from_stylecode(self, input_semantics, latent_style):
        return self.netG(input_semantics, latent_style)

    def generate_fake(self, input_semantics, style_image):
        latent_style, features = self.encode_w(style_image)
        fake_image = self.generate_fake_from_stylecode(input_semantics, latent_style)
        return fake_image, latent_style, features

    def discriminate(self, input_semantics, fake_image, real_image):
        # both concatenations of the reference (pix2pix_model.py:328-338) are one layout kernel
        # channels zero-padded to 16 so that D's first 4x4-s2 convolution (5 -> 64) is tensor-core shaped after the
        # space-to-depth step (4*16 = 64 input channels); the padded channels carry zero weights
        fake_and_real = ops.MakeDInputFn.apply(input_semantics, fake_image, real_image, 16)
        discriminator_out = self.netD.forward_nhwc(fake_and_real)
        self._last_d_out = discriminator_out
        return self.divide_pred(discriminator_out)

    def divide_pred(self, pred):
        if type(pred) == list:
            fake = [[t[:t.size(0) // 2] for t in p] for p in pred]
            real = [[t[t.size(0) // 2:] for t in p] for p in pred]
        else:
            fake, real = pred[:pred.size(0) // 2], pred[pred.size(0) // 2:]
        return fake, real

    def use_gpu(self):
        return len(self.opt.gpu_ids) > 0
