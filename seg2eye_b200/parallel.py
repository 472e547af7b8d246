"""Data parallelism: one process per GPU, full replicas, batch sharded, per-rank BatchNorm statistics (what the
reference's nn.DataParallel did); the only collective is the gradient average.

GradReducer averages gradients with bucketed all-reduces that OVERLAP the backward pass: every parameter carries a
post-accumulate-grad hook that copies its gradient into its bucket's flat fp32 buffer; when the last gradient of a
bucket has arrived the bucket's all-reduce is launched asynchronously (NCCL runs it on its own stream after the work
already queued on ours, so it proceeds while the remaining backward kernels run).  `allreduce()` after backward waits
for the buckets and scatters the averaged values back into `.grad`.  Buckets are built from the parameters that
actually received a gradient in the first step (fc_var never does; D's parameters get none inside the G step), in
reverse registration order ~ the order in which backward produces them.  NCCL over NVLink on the box, gloo in the CPU
tests."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "pending", "handle")

    def __init__(self, params):
        self.params = params
        self.offsets, n = [], 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        self.pending = len(params)
        self.handle = None


class GradReducer:
    def __init__(self, params, bucket_bytes=32 << 20, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.overlap = overlap
        self.buckets = None
        self._slot = {}
        self._hooks = []
        self.launched_during_backward = 0   # diagnostics: buckets whose all-reduce started from a hook

    # ---- bucket construction (first step: we now know which parameters receive gradients)
    def _build(self):
        live = [p for p in reversed(self.params) if p.grad is not None]
        self.buckets, cur, size = [], [], 0
        for p in live:
            n = p.numel() * 4
            if cur and size + n > self.bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            self.buckets.append(_Bucket(cur))
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._slot[p] = (bi, pi)
        if self.overlap:
            for p in live:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p):
        bi, pi = self._slot[p]
        b = self.buckets[bi]
        off = b.offsets[pi]
        b.flat[off:off + p.numel()].copy_(p.grad.detach().reshape(-1))
        b.pending -= 1
        if b.pending == 0:
            b.handle = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, async_op=True)
            self.launched_during_backward += 1

    # ---- called once after backward
    def allreduce(self):
        ws = world_size()
        if ws == 1:
            return
        first = self.buckets is None
        if first:
            self._build()
        for b in self.buckets:
            if b.handle is None:
                # first step (hooks were not installed yet) or a gradient that did not arrive: gather what exists
                for p, off in zip(b.params, b.offsets):
                    if p.grad is not None:
                        b.flat[off:off + p.numel()].copy_(p.grad.detach().reshape(-1))
                    else:
                        b.flat[off:off + p.numel()].zero_()
                b.handle = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, async_op=True)
        inv = 1.0 / ws
        for b in self.buckets:
            b.handle.wait()
            for p, off in zip(b.params, b.offsets):
                if p.grad is not None:
                    p.grad.copy_(b.flat[off:off + p.numel()].view_as(p.grad)).mul_(inv)
            b.handle = None
            b.pending = len(b.params)


def broadcast_module(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)
