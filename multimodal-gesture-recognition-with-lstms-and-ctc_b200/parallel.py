"""Data parallelism: one process per GPU, batch sharded, ONE flat fp32 gradient all-reduce per
step over NCCL/NVLink (no reference counterpart: the reference is single-GPU).
The reduction happens BEFORE clipvalue/Adam/maxnorm, which then run identically on every rank."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_rows(n_rows, rank, world):
    """Rank r owns rows [r*n/world, (r+1)*n/world) of the global batch."""
    return (rank * n_rows) // world, ((rank + 1) * n_rows) // world


class FlatViews(list):
    """Per-parameter views of one flat buffer (`.flat`): an optimiser that sees it can update every tensor in one launch."""
    flat = None


class FlatGradBucket:
    """All trainable gradients in one flat fp32 buffer -> a single all-reduce (sum)."""

    def __init__(self, params):
        self.shapes = [tuple(p.shape) for p in params]
        self.sizes = [int(p.numel()) for p in params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=params[0].device)
        self.views = FlatViews()
        self.views.flat = self.flat
        o = 0
        for n, s in zip(self.sizes, self.shapes):
            self.views.append(self.flat[o:o + n].view(s))
            o += n

    def pack(self, grads):
        if self.flat.is_cuda:
            from . import ops
            ops.pack(list(grads), self.flat)          # one launch for all tensors
        else:
            for v, g in zip(self.views, grads):       # host-side tests (gloo)
                v.copy_(g)
        return self.flat

    def all_reduce(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.views
