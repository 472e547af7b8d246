"""Pix2PixTrainer mirror (reference trainers/pix2pix_trainer.py:9-88).  With torch.distributed initialised the
G+E and D gradients are averaged across ranks by bucketed NCCL all-reduce (seg2eye_b200.parallel)."""
from ..models.pix2pix_model import Pix2PixModel
from .. import parallel


class Pix2PixTrainer():
    def __init__(self, opt):
        self.opt = opt
        self.pix2pix_model = Pix2PixModel(opt)
        self.pix2pix_model_on_one_gpu = self.pix2pix_model
        self.generated = None
        if opt.isTrain:
            self.optimizer_G, self.optimizer_D = self.pix2pix_model_on_one_gpu.create_optimizers(opt)
            self.old_lr = opt.lr
            m = self.pix2pix_model
            self.reducer_G = parallel.GradReducer(list(m.netG.parameters()) + list(m.netE.parameters()))
            self.reducer_D = parallel.GradReducer(list(m.netD.parameters()))

    def run_generator_one_step(self, data):
        self.pix2pix_model.train()
        self.optimizer_G.zero_grad()
        g_losses, generated = self.pix2pix_model(data, mode='generator')
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        self.reducer_G.allreduce()
        self.optimizer_G.step()
        self.g_losses = g_losses
        self.generated = generated

    def run_discriminator_one_step(self, data):
        self.pix2pix_model.train()
        self.optimizer_D.zero_grad()
        d_losses = self.pix2pix_model(data, mode='discriminator')
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        self.reducer_D.allreduce()
        self.optimizer_D.step()
        self.d_losses = d_losses

    def get_latest_losses(self, include_log_losses=False):
        losses = {**self.g_losses, **self.d_losses}
        if include_log_losses:
            losses = {**losses, **self.pix2pix_model_on_one_gpu.get_loss_log()}
            self.pix2pix_model_on_one_gpu.reset_loss_log()
        return losses

    def get_latest_generated(self):
        return self.generated

    def save(self, epoch):
        self.pix2pix_model_on_one_gpu.save(epoch)

    def update_learning_rate(self, epoch):
        if epoch > self.opt.niter:
            lrd = self.opt.lr / self.opt.niter_decay
            new_lr = self.old_lr - lrd
        else:
            new_lr = self.old_lr
        if new_lr != self.old_lr:
            if self.opt.no_TTUR:
                new_lr_G, new_lr_D = new_lr, new_lr
            else:
                new_lr_G, new_lr_D = new_lr / 2, new_lr * 2
            for param_group in self.optimizer_D.param_groups:
                param_group['lr'] = new_lr_D
            for param_group in self.optimizer_G.param_groups:
                param_group['lr'] = new_lr_G
            print('update learning rate: %f -> %f' % (self.old_lr, new_lr))
            self.old_lr = new_lr
