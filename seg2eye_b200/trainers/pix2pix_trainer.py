"""Pix2PixTrainer mirror (reference trainers/pix2pix_trainer.py:9-88).  With torch.distributed initialised the
G+E and D gradients are averaged across ranks by bucketed NCCL all-reduce (seg2eye_b200.parallel).

Beyond the reference: `enable_cuda_graphs(example_batch)` captures the whole generator step and the whole
discriminator step (forward, backward, all-reduce-free single-GPU case, Adam) into two CUDA graphs, so that a
training iteration is two graph launches instead of ~1300 kernel launches driven from Python."""
import torch

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

    # ------------------------------------------------------------------ CUDA-graph fast path
    def _g_step_body(self, data):
        g_losses, generated = self.pix2pix_model(data, mode='generator')
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        self.reducer_G.allreduce()
        self.optimizer_G.step()
        return g_losses, generated

    def _d_step_body(self, data):
        d_losses = self.pix2pix_model(data, mode='discriminator')
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        self.reducer_D.allreduce()
        self.optimizer_D.step()
        return d_losses

    def enable_cuda_graphs(self, example_data, warmup=3):
        """Capture the G step and the D step for batches shaped like `example_data` (label (B,1,H,W),
        style_image (B,ns,1,H,W), target (B,1,H,W)).  Afterwards run_*_one_step copy the batch into static device
        buffers and replay.  Shapes must not change; call disable_cuda_graphs() to return to eager execution."""
        if parallel.world_size() > 1:
            # capturing the NCCL all-reduce inside the step graph deadlocked on the 2-GPU box (round 1); multi-GPU
            # runs use the eager path until the collective is moved outside the captured region
            raise RuntimeError("CUDA-graph steps are single-process for now; multi-GPU runs use the eager path")
        dev = self.pix2pix_model.device()
        self.pix2pix_model.train()
        # autograd graphs of earlier eager steps keep the parameters' AccumulateGrad nodes bound to the legacy stream,
        # which cannot be joined from a capturing stream: drop every reference to them before warming up
        if getattr(self, 'g_losses', None) is not None:
            self.g_losses = {k: v.detach() for k, v in self.g_losses.items()}
        if getattr(self, 'd_losses', None) is not None:
            self.d_losses = {k: v.detach() for k, v in self.d_losses.items()}
        if self.generated is not None:
            self.generated = self.generated.detach()
        self.pix2pix_model.reset_loss_log()
        import gc
        gc.collect()
        self._static = {'label': example_data['label'].long().to(dev).clone(),
                        'style_image': example_data['style_image'].float().to(dev).clone(),
                        'target': example_data['target'].float().to(dev).clone()}
        # the warm-up iterations below are real training steps: snapshot every piece of state they touch (weights,
        # BN / spectral-norm buffers, Adam moments and step counters) and restore it in place after the capture
        m = self.pix2pix_model
        tensors = [t for net in (m.netG, m.netD, m.netE) for t in list(net.parameters()) + list(net.buffers())]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # one throw-away iteration so that optimizer state exists before the snapshot
            self.optimizer_G.zero_grad(set_to_none=True)
            self.optimizer_D.zero_grad(set_to_none=True)
            had_state = len(self.optimizer_G.state) > 0
            if not had_state:
                snap0 = [t.detach().clone() for t in tensors]
                self._g_step_body(dict(self._static))
                self._d_step_body(dict(self._static))
                for opt_ in (self.optimizer_G, self.optimizer_D):
                    for st in opt_.state.values():
                        st['exp_avg'].zero_()
                        st['exp_avg_sq'].zero_()
                    for g in opt_.param_groups:
                        if g.get('_s2e_state') is not None:
                            g['_s2e_state'][0].zero_()
                for t, s0 in zip(tensors, snap0):
                    t.detach().copy_(s0)
        torch.cuda.current_stream().wait_stream(side)
        opt_tensors = [st[k] for opt_ in (self.optimizer_G, self.optimizer_D) for st in opt_.state.values()
                       for k in ('exp_avg', 'exp_avg_sq')]
        opt_tensors += [g['_s2e_state'] for opt_ in (self.optimizer_G, self.optimizer_D) for g in opt_.param_groups
                        if g.get('_s2e_state') is not None]
        snap = [t.detach().clone() for t in tensors + opt_tensors]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):     # allocates optimizer state, fills the packed-weight cache, sets func attributes
                self.optimizer_G.zero_grad(set_to_none=True)
                self._g_step_body(dict(self._static))
                self.optimizer_D.zero_grad(set_to_none=True)
                self._d_step_body(dict(self._static))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()    # the graphs allocate from a private pool; give the eager cache back first
        self._graph_G, self._graph_D = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        self.optimizer_G.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph_G):
            self._g_out = self._g_step_body(dict(self._static))
        self.optimizer_D.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph_D, pool=self._graph_G.pool()):
            self._d_out = self._d_step_body(dict(self._static))
        for t, s0 in zip(tensors + opt_tensors, snap):
            t.detach().copy_(s0)
        from .. import ops
        ops.bump_weights_epoch()
        ops.repack_stale()   # the graphs re-pack after their own Adam; the restored weights need it once, here
        self._graphs = True

    def disable_cuda_graphs(self):
        self._graphs = False
        self._graph_G = self._graph_D = self._g_out = self._d_out = None

    def _load_static(self, data):
        for k, buf in self._static.items():
            src = data[k]
            if src.shape != buf.shape:
                raise RuntimeError("CUDA-graph step captured for %s %s, got %s" % (k, tuple(buf.shape), tuple(src.shape)))
            if src.data_ptr() != buf.data_ptr():
                buf.copy_(src, non_blocking=True)
        data.update(self._static)   # the reference mutates `data` in place too (pix2pix_model.py:140-158)

    def run_generator_one_step(self, data):
        if getattr(self, '_graphs', False):
            self._load_static(data)
            self._graph_G.replay()
            self.g_losses, self.generated = self._g_out
            return
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
        if getattr(self, '_graphs', False):
            self._load_static(data)
            self._graph_D.replay()
            self.d_losses = self._d_out
            return
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
            self.optimizer_D.sync_hyperparams()
            self.optimizer_G.sync_hyperparams()
            print('update learning rate: %f -> %f' % (self.old_lr, new_lr))
            self.old_lr = new_lr
