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

    def allreduce_mean(self, group=None, async_op: bool = False, local_items: int = None, global_items: int = None):
        """sum over ranks, then scale by 1/world (the reference averages the loss over the global batch).  With uneven
        shards (`shard_range`) pass the shard sizes: each rank's gradient of its LOCAL mean is weighted by
        local_items / global_items before the sum, which gives the gradient of the global-batch mean, not a mean of means."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return None
        weighted = local_items is not None and global_items is not None
        if weighted:
            self.flat.mul_(float(local_items) / float(global_items))
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work           # an unweighted async caller scales by 1/world itself once the work has completed
        if not weighted:
            self.flat.mul_(1.0 / world)
        return None


class FlatTrainState:
    """Everything the tail of a training step touches, as FLAT fp32 buffers with one contiguous SEGMENT per module:
    parameters (every `p.data` becomes a view), gradients (`p.grad` views; each module also gets `_grad_sink`, so the C
    backward accumulates straight into the bucket), and Adam's two moments.  Segment s = [offsets[s], offsets[s+1]).

    Why segments: the reference clips the gradient norm per module (interfaces/super_resolution.py:270-275), and the
    all-reduce of a module's segment can start as soon as that module's backward has been enqueued -- the CMM (94 % of the
    bucket) finishes its backward FIRST, so its segment travels over NVLink while the six PGRMs still compute."""

    ALIGN = 64          # elements (256 bytes)

    def __init__(self, modules):
        """modules: list of nn.Module in segment order."""
        self.modules = list(modules)
        per_mod = []
        for m in self.modules:
            named = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
            per_mod.append(named)
        self.params = [p for named in per_mod for _, p in named]
        if not self.params:
            raise ValueError("FlatTrainState: no trainable parameters")
        dev = self.params[0].device
        # every PARAMETER starts on a 256-byte boundary inside the flat buffers, like a tensor of its own from torch's
        # allocator would: the kernels read weights with 16-byte vector loads and TMA.  The padding elements stay zero in
        # all four buffers (zero gradient -> zero moments -> zero update), so norms and the all-reduce are unaffected.
        A = self.ALIGN
        sizes = [sum((p.numel() + A - 1) // A * A for _, p in named) for named in per_mod]
        self.offsets = [0]
        for sz in sizes:
            self.offsets.append(self.offsets[-1] + sz)
        total = self.offsets[-1]
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = None          # allocated by the optimizer that uses them
        self.exp_avg_sq = None
        self.n_params = sum(p.numel() for p in self.params)
        for m, named, start in zip(self.modules, per_mod, self.offsets[:-1]):
            off = start
            sink = {}
            for n, p in named:
                k = p.numel()
                self.flat_params[off: off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_params[off: off + k].view(p.shape)
                g = self.flat_grads[off: off + k].view(p.shape)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
                sink[n] = g
                off += (k + A - 1) // A * A
            m._grad_sink = sink
            m._weights_epoch = getattr(m, "_weights_epoch", 0) + 1     # parameter storage moved: re-stage cached weights

    def segment(self, s: int) -> torch.Tensor:
        return self.flat_grads[self.offsets[s]: self.offsets[s + 1]]

    def zero_grads(self):
        self.flat_grads.zero_()

    def bump_epoch(self):
        for m in self.modules:
            m._weights_epoch = getattr(m, "_weights_epoch", 0) + 1


class NcclBucketComm:
    """The library's own NCCL communicator (dpmn_nccl_comm_init / dpmn_allreduce_bucket in include/dpmn_b200.h): rank 0
    makes the unique id, the existing torch.distributed group only carries those 128 bytes to the other ranks."""

    def __init__(self, device: torch.device, group=None):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        if not self.lib.dpmn_nccl_available():
            raise RuntimeError("dpmn_b200: libnccl could not be resolved (dpmn_nccl_available() == 0)")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_char * 128)()
        if self.rank == 0:
            _lib.check(self.lib.dpmn_nccl_unique_id(C.cast(buf, C.c_void_p)), "dpmn_nccl_unique_id")
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idbuf = C.create_string_buffer(box[0], 128)
        self._comm = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.dpmn_nccl_comm_init(C.cast(idbuf, C.c_void_p), self.world, self.rank, C.byref(self._comm)),
                       "dpmn_nccl_comm_init")
        self.device = device

    def allreduce_sum(self, t: torch.Tensor, stream: torch.cuda.Stream):
        """In-place sum over ranks of a contiguous fp32 / fp16 / bf16 CUDA tensor, enqueued on `stream`."""
        from . import _lib
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("NcclBucketComm.allreduce_sum: contiguous CUDA tensor required")
        dt = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[t.dtype]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dpmn_allreduce_bucket(self._comm, t.data_ptr(), t.numel(), dt, stream.cuda_stream),
                       "dpmn_allreduce_bucket")

    def close(self):
        if self._comm:
            self.lib.dpmn_nccl_comm_destroy(self._comm)
            self._comm = None


def broadcast_module_state(modules, src: int = 0, group=None):
    """Rank `src`'s parameters and buffers to every rank (what DDP does at construction): replicas must start identical,
    whatever each rank's RNG state was when it built its modules."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def max_over_ranks(value: float, device) -> float:
    """Timing helper of bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
