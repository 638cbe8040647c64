"""Data-parallel plumbing of the hot path: one process per GPU, the batch sharded across ranks,
and -- for training -- ONE flat all-reduce per step over a contiguous gradient bucket.

The reference wraps its modules in single-process `nn.DataParallel` (interfaces/base.py:160-162), which
re-broadcasts 57 M parameters every forward and reduces gradients onto GPU 0.  Here every rank owns a replica;
inference needs no collective at all, training needs exactly one `all_reduce(sum)` (NCCL over NVLink/NVSwitch on
the GPU box, gloo in the CPU tests) followed by a 1/world scale.  BatchNorm statistics stay per replica, as under
DataParallel (no SyncBN in the reference).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a global batch for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of `params`; each `p.grad` is a view into it, so a
    single collective covers the 6 PGRMs + CMM (+ distill modules) = 57.2 M parameters / 228.7 MB."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            view = self.flat[off: off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view           # parameters that never receive a gradient contribute zeros
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None, async_op: bool = False):
        """sum over ranks, then scale by 1/world (the reference averages the loss over the global batch)."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.flat.mul_(1.0 / world)
        return None


def max_over_ranks(value: float, device) -> float:
    """Timing helper of bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
