"""Data parallelism: one process per GPU, gradients averaged with bucketed all-reduce (NCCL over NVLink on
the GPU box, gloo in the CPU tests).  The reference has no multi-process path (README.md:58); this follows
SURVEY.md 8(e): full replicas, batch sharded, per-rank BatchNorm statistics (what nn.DataParallel did),
parameters that never receive a gradient (fc_var) are skipped consistently on every rank."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class GradReducer:
    def __init__(self, params, bucket_bytes=64 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self._flat = {}

    def buckets(self):
        """Deterministic buckets over the parameters that currently hold a gradient (same on all ranks)."""
        out, cur, size = [], [], 0
        for p in self.params:
            if p.grad is None:
                continue
            n = p.numel() * 4
            if cur and size + n > self.bucket_bytes:
                out.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            out.append(cur)
        return out

    def allreduce(self):
        ws = world_size()
        if ws == 1:
            return
        handles = []
        for i, bucket in enumerate(self.buckets()):
            n = sum(p.numel() for p in bucket)
            flat = self._flat.get(i)
            if flat is None or flat.numel() != n or flat.device != bucket[0].grad.device:
                flat = torch.empty(n, dtype=torch.float32, device=bucket[0].grad.device)
                self._flat[i] = flat
            off = 0
            for p in bucket:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
            handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
        for h, flat, bucket in handles:
            h.wait()
            off = 0
            for p in bucket:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad)).div_(ws)
                off += p.numel()


def broadcast_module(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)
