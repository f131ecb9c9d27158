"""Multi-GPU plumbing for the geometry path: one process per GPU (torchrun), the batch (or view) axis sharded
contiguously, topology replicated, and ONE all-reduce (SUM) per step on the flattened gradient of the
parameters shared by the batch.  The reference's equivalent is nn.DataParallel's implicit reduce-add
(train_multigpu.py:136-140,273); it has no NCCL call site.

Works with any torch.distributed backend: "nccl" on the B200 box, "gloo" in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` (the first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank: int, world: int):
    """Slice every (B, ...) tensor of a dict to this rank's samples."""
    out = {}
    for k, t in tensors.items():
        lo, hi = shard_range(t.shape[0], rank, world)
        out[k] = t[lo:hi]
    return out


class GradBucket:
    """Flatten a list of gradient tensors into one buffer so that the step issues a single collective."""

    def __init__(self, params):
        self.params = list(params)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce(self, average: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        if len(grads) == 1:
            dist.all_reduce(grads[0])
            if average:
                grads[0].div_(dist.get_world_size())
            self.params[0].grad = grads[0]      # a rank whose grad was None still receives the other ranks' sum
            return
        if self.flat is None or self.flat.device != grads[0].device:
            self.flat = torch.empty(self.numel, device=grads[0].device, dtype=grads[0].dtype)
        off = 0
        for g in grads:
            self.flat[off:off + g.numel()].copy_(g.reshape(-1))
            off += g.numel()
        dist.all_reduce(self.flat)
        if average:
            self.flat.div_(dist.get_world_size())
        off = 0
        for p, g in zip(self.params, grads):
            g.copy_(self.flat[off:off + g.numel()].view_as(g))
            p.grad = g
            off += g.numel()
