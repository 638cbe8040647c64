"""One data-parallel training step of the hot path as the reference runs it (interfaces/super_resolution.py:245-275):
seven image losses (six PGRM outputs + the CMM output), one backward, per-module gradient clipping at 0.25,
Adam(lr 1e-3, betas (0.5, 0.999)) (interfaces/base.py:208-221).  Forward and backward of every PGRM / CMM run in
libdpmn_b200 (dpmn_*_forward / dpmn_*_backward through the modules' autograd Functions); across ranks the ONLY
collective is one flat all-reduce over the gradient bucket.

The image loss (ImageLoss, loss/image_loss.py:15-43 -- SURVEY 8f rank 1) is one CUDA kernel per term (value + gradient);
caller-side pieces kept in torch, exactly where the reference keeps them: clip_grad_norm_ and the Adam update.
The four DistillModule terms (model/distill_module.py; super_resolution.py:245-263 -- SURVEY 8f rank 3) run in
libdpmn_b200 too (dpmn_b200.distill.DistillModule) and their parameters sit in the same gradient bucket.
Not reproduced: the recogniser / text-rendering loop that produces the priors (out of scope: priors are inputs)."""
from __future__ import annotations

from typing import List, Sequence

import torch

from .pipeline import DPMNHotPath


class _ImageLoss(torch.autograd.Function):
    """ImageLoss value and gradient in one kernel (dpmn_image_loss, csrc/loss_mask.cu)."""

    @staticmethod
    def forward(ctx, out, target, w_mse, w_gp):
        from . import _lib
        lib = _lib.load()
        if not (out.is_cuda and target.is_cuda and out.dtype == torch.float32 and target.dtype == torch.float32):
            raise RuntimeError("dpmn_b200.train.image_loss: fp32 CUDA tensors required (there is no CPU path)")
        B, C, H, W = out.shape

        def arg(t):
            if t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W:
                return t, (t.stride(0) if B > 1 else C * H * W)
            t = t.contiguous()
            return t, C * H * W
        o, o_bs = arg(out)
        t, t_bs = arg(target)
        loss = torch.zeros((), dtype=torch.float32, device=out.device)
        d_out = torch.empty((B, C, H, W), dtype=torch.float32, device=out.device) if out.requires_grad else None
        with torch.cuda.device(out.device):
            rc = lib.dpmn_image_loss(o.data_ptr(), o_bs, t.data_ptr(), t_bs, B, C, H, W, float(w_mse), float(w_gp), 1.0,
                                     loss.data_ptr(), d_out.data_ptr() if d_out is not None else None,
                                     torch.cuda.current_stream(out.device).cuda_stream)
        _lib.check(rc, "dpmn_image_loss")
        ctx.d_out = d_out
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_out
        ctx.d_out = None
        return (d * g if d is not None else None), None, None, None


def image_loss(out: torch.Tensor, target: torch.Tensor, weight=(1.0, 1.0)) -> torch.Tensor:
    """ImageLoss(gradient=True, loss_weight=[1, 1]) as interfaces/base.py:132 constructs it: MSE + L1 of the gradient
    maps (loss/image_loss.py:15-43), value and gradient from one CUDA kernel."""
    return _ImageLoss.apply(out, target, weight[0], weight[1])


def to_mask(images: torch.Tensor) -> torch.Tensor:
    """toMask (utils/util.py:27-35) for a whole batch on the device: (B, 3, H, W) in [0, 1] -> (B, 3, H, W) in {0, 1}.
    The reference loops over images through PIL on the host (interfaces/super_resolution.py:220-226)."""
    from . import _lib
    lib = _lib.load()
    if not (images.is_cuda and images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3):
        raise RuntimeError("dpmn_b200.train.to_mask: a (B, 3, H, W) fp32 CUDA tensor is required")
    B, _, H, W = images.shape
    x = images if (images.stride(3) == 1 and images.stride(2) == W and images.stride(1) == H * W) else images.contiguous()
    bs = x.stride(0) if B > 1 else 3 * H * W
    mask = torch.empty((B, 3, H, W), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        rc = lib.dpmn_to_mask(x.data_ptr(), bs, mask.data_ptr(), B, H, W, torch.cuda.current_stream(images.device).cuda_stream)
    _lib.check(rc, "dpmn_to_mask")
    return mask


def _resize_call(fn_name: str, images: torch.Tensor, out_ch: int, out_hw):
    from . import _lib
    lib = _lib.load()
    if not (images.is_cuda and images.dtype == torch.float32 and images.dim() == 4 and images.shape[1] == 3):
        raise RuntimeError(f"dpmn_b200.train.{fn_name}: a (B, 3, H, W) fp32 CUDA tensor is required (there is no CPU path)")
    B, _, H, W = images.shape
    x = images if (images.stride(3) == 1 and images.stride(2) == W and images.stride(1) == H * W) else images.contiguous()
    bs = x.stride(0) if B > 1 else 3 * H * W
    out = torch.empty((B, out_ch, out_hw[0], out_hw[1]), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        rc = getattr(lib, fn_name)(x.data_ptr(), bs, out.data_ptr(), B, H, W, out_hw[0], out_hw[1],
                                   torch.cuda.current_stream(images.device).cuda_stream)
    _lib.check(rc, fn_name)
    return out


def parse_crnn_data(imgs_input: torch.Tensor) -> torch.Tensor:
    """TextBase.parse_crnn_data (interfaces/base.py:419-425): bicubic resize to (32, 100) + luma in one kernel.
    (B, 3, H, W) -> (B, 1, 32, 100).  Accepts the channel-slice view images[:, :3] of the call sites."""
    return _resize_call("dpmn_crnn_input", imgs_input, 1, (32, 100))


def parse_visionlan_data(imgs_input: torch.Tensor) -> torch.Tensor:
    """TextBase.parse_visionlan_data (interfaces/base.py:473-478) for a whole batch on the device: (B, 3, H, W) in [0, 1]
    -> (B, 3, 64, 256), bit-exact against the reference's per-image ToPILImage -> cv2.resize -> ToTensor round trip.  A
    single (3, H, W) image (the reference's signature) gives (1, 3, 64, 256) like the reference."""
    if imgs_input.dim() == 3:
        imgs_input = imgs_input.unsqueeze(0)
    return _resize_call("dpmn_visionlan_input", imgs_input, 3, (64, 256))


class HotPathTrainer:
    """One training step of the hot path: forward (6 PGRM + CMM), 7 image losses + the DistillModule terms, backward, the
    gradient all-reduce, per-module clipping at 0.25 and Adam (interfaces/super_resolution.py:245-278, base.py:208-221).

    Data parallel: one process per GPU, replicas made identical at construction by a broadcast from rank 0,
    ONE flat fp32 gradient bucket (57.2 M parameters, 228.7 MB) reduced with the library's own NCCL communicator
    (dpmn_allreduce_bucket) on a side stream.  The bucket is laid out module by module with the CMM first: its backward
    is the first to finish, so its segment (94 % of the bytes) is reduced while the six PGRMs are still in their
    backward; the remaining segment follows when the backward ends.  Clip + Adam then run as two kernels over the flat
    buffers (dpmn_clip_adam_step).  Per-rank BatchNorm statistics, as under the reference's DataParallel (no SyncBN)."""

    def __init__(self, model: DPMNHotPath, lr: float = 1e-3, betas=(0.5, 0.999), clip: float = 0.25, group=None,
                 distill: bool = True, fused_optimizer: bool = None, overlap: bool = True, eps: float = 1e-8):
        import torch.distributed as dist
        from .dist import FlatTrainState, NcclBucketComm, broadcast_module_state
        self.model, self.clip, self.group = model, clip, group
        self.lr, self.betas, self.eps = lr, betas, eps
        dev = next(model.parameters()).device
        self.device = dev
        self.dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.rank = dist.get_rank(group) if self.dist_on else 0
        # one DistillModule per adjacent pair of cascade outputs of a branch (super_resolution.py:113-121)
        self.distill: List[torch.nn.Module] = []
        if distill:
            from .distill import DistillModule
            with torch.random.fork_rng(devices=[]):
                torch.default_generator.manual_seed(20240 + model.b1 * 16 + model.b2)   # CPU generator only
                self.distill = [DistillModule().to(dev).train() for _ in range(model.b1 + model.b2 - 2)]
        self.modules = list(model.pgrm) + [model.cmm]                  # model_list of the reference (clip order)
        seg_modules = [model.cmm] + list(model.pgrm) + self.distill    # bucket order: the CMM's segment first (see above)
        broadcast_module_state(seg_modules, src=0, group=group)        # identical replicas whatever each rank's RNG state
        for m in model.pgrm:
            m._seed_salt = self.rank                                   # different Dropout / DropPath masks per rank
        self.state = FlatTrainState(seg_modules)
        self.bucket = self.state                                       # (name kept for callers of the round-1 interface)
        cuda = dev.type == "cuda"
        self.fused = cuda if fused_optimizer is None else bool(fused_optimizer)
        self.step_count = 0
        if self.fused:
            from . import _lib
            self._lib = _lib.load()
            self.state.exp_avg = torch.zeros_like(self.state.flat_params)
            self.state.exp_avg_sq = torch.zeros_like(self.state.flat_params)
            n_seg = len(seg_modules)
            self._opt_ws = torch.zeros(int(self._lib.dpmn_clip_adam_workspace_bytes(n_seg)) + 16, dtype=torch.uint8, device=dev)
            import ctypes as C
            self._offs = (C.c_int64 * (n_seg + 1))(*self.state.offsets)
            self.opt = None
        else:
            self.opt = torch.optim.Adam(self.state.params, lr=lr, betas=betas, eps=eps)
        # the collective
        self.comm, self.comm_stream = None, None
        self.overlap = overlap
        self.timing = None               # set to {} to collect all-reduce timings (CUDA events) per step
        self._events = []
        if self.dist_on and cuda:
            self.comm = NcclBucketComm(dev, group)
            self.comm_stream = torch.cuda.Stream(dev)
            if overlap:
                model.cmm._after_backward = self._on_cmm_backward
        self._cmm_reduced = False

    # ---- collective ---------------------------------------------------------------------------------------------------
    def _reduce_segment(self, t: torch.Tensor, tag: str):
        """all-reduce(sum) of a slice of the flat gradient bucket on the communication stream, ordered after everything
        enqueued on the current stream so far."""
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        self.comm_stream.wait_event(ready)
        t0 = t1 = None
        if self.timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(self.comm_stream)
        self.comm.allreduce_sum(t, self.comm_stream)
        if self.timing is not None:
            t1.record(self.comm_stream)
            self._events.append((tag, t0, t1))

    def _on_cmm_backward(self, module):
        if self.comm is None or self._cmm_reduced:
            return
        self._reduce_segment(self.state.segment(0), "allreduce_cmm_segment")
        self._cmm_reduced = True

    def _allreduce(self):
        import torch.distributed as dist
        st = self.state
        if self.comm is None:                        # CPU tensors / gloo (the host-logic tests): torch's collective
            dist.all_reduce(st.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
            return
        main = torch.cuda.current_stream(self.device)
        rest = st.flat_grads[st.offsets[1]:] if self._cmm_reduced else st.flat_grads
        self._reduce_segment(rest, "allreduce_tail_segment" if self._cmm_reduced else "allreduce_whole_bucket")
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        w0 = w1 = None
        if self.timing is not None:
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(main)
        main.wait_event(done)
        if self.timing is not None:
            w1.record(main)
            self._events.append(("allreduce_exposed_wait", w0, w1))

    def collect_timing(self):
        """ms per tag, summed over the steps since the last call (needs self.timing = {} before the steps; synchronises)."""
        torch.cuda.synchronize(self.device)
        out = {}
        for tag, a, b in self._events:
            out[tag] = out.get(tag, 0.0) + a.elapsed_time(b)
        self._events = []
        return out

    # ---- loss -----------------------------------------------------------------------------------------------------------
    def distill_loss(self, outs: Sequence[torch.Tensor]) -> torch.Tensor:
        """super_resolution.py:245-263: per branch, from the deepest cascade image back to the first; each module
        compares its fused feature with the next shallower image and hands the feature on."""
        b1, b2 = self.model.b1, self.model.b2
        total = outs[0].new_zeros(())
        for imgs, first in ((outs[:b1], 0), (outs[b1:b1 + b2], b1 - 1)):
            feature = imgs[-1]
            for k in range(len(imgs) - 1, 0, -1):
                l, feature = self.distill[first + k - 1](feature, imgs[k - 1])      # distill_list[k-1] / [k+b1-2]
                total = total + l.sum() * 100
        return total

    def loss(self, outs: Sequence[torch.Tensor], hr: torch.Tensor) -> torch.Tensor:
        total = outs[0].new_zeros(())
        for o in outs:
            total = total + image_loss(o, hr[:, :3]) * 100        # super_resolution.py:212,239,267
        if self.distill:
            total = total + self.distill_loss(outs)                # :253,263
        return total / len(outs)                                   # :268

    # ---- optimizer --------------------------------------------------------------------------------------------------------
    def _optimizer_step(self):
        st = self.state
        self.step_count += 1
        if not self.fused:
            if self.world > 1:
                st.flat_grads.mul_(1.0 / self.world)
            for m in self.modules + self.distill:                  # super_resolution.py:270-275
                torch.nn.utils.clip_grad_norm_(m.parameters(), self.clip)
            self.opt.step()
            return
        from . import _lib
        with torch.cuda.device(self.device):
            rc = self._lib.dpmn_clip_adam_step(st.flat_params.data_ptr(), st.flat_grads.data_ptr(), st.exp_avg.data_ptr(),
                                               st.exp_avg_sq.data_ptr(), self._offs, len(st.modules), 1.0 / self.world,
                                               float(self.clip), float(self.lr), float(self.betas[0]), float(self.betas[1]),
                                               float(self.eps), self.step_count, self._opt_ws.data_ptr(), self._opt_ws.numel(),
                                               torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "dpmn_clip_adam_step")
        st.bump_epoch()          # parameters changed behind torch's version counters: cached 16-bit weights are stale

    def step(self, psn_out, priors_b1, priors_b2, hr, global_batch: int = None) -> torch.Tensor:
        """forward + backward + all-reduce + clip + Adam; returns the (local) loss as a 0-d tensor.  `global_batch`: total
        images over all ranks when the shards are uneven (dist.shard_range) -- each rank's loss is then weighted by
        local / (global / world), so that the summed gradients / world are those of the global-batch mean."""
        self.state.zero_grads()
        self._cmm_reduced = False
        outs = self.model.forward_all(psn_out, priors_b1, priors_b2)
        loss = self.loss(outs, hr)
        w = 1.0
        if self.world > 1 and global_batch:
            w = psn_out.shape[0] * self.world / float(global_batch)
        (loss * w if w != 1.0 else loss).backward()
        if self.dist_on:
            self._allreduce()
        self._optimizer_step()
        return loss.detach()
