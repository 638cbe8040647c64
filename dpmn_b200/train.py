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

from .dist import FlatGradBucket
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
    def __init__(self, model: DPMNHotPath, lr: float = 1e-3, betas=(0.5, 0.999), clip: float = 0.25, group=None,
                 distill: bool = True):
        self.model, self.clip, self.group = model, clip, group
        self.modules = list(model.pgrm) + [model.cmm]
        # one DistillModule per adjacent pair of cascade outputs of a branch (super_resolution.py:113-121)
        self.distill: List[torch.nn.Module] = []
        if distill:
            from .distill import DistillModule
            dev = next(model.parameters()).device
            with torch.random.fork_rng(devices=[]):            # same initial replicas on every rank
                torch.default_generator.manual_seed(20240 + model.b1 * 16 + model.b2)   # CPU generator only
                self.distill = [DistillModule().to(dev).train() for _ in range(model.b1 + model.b2 - 2)]
        params = list(model.parameters()) + [p for m in self.distill for p in m.parameters()]   # base.py:208-221 order
        self.bucket = FlatGradBucket(params)                   # p.grad become views of ONE flat fp32 buffer
        self.opt = torch.optim.Adam(self.bucket.params, lr=lr, betas=betas)

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

    def step(self, psn_out, priors_b1, priors_b2, hr) -> torch.Tensor:
        """forward + backward + all-reduce + clip + Adam; returns the (local) loss as a 0-d tensor."""
        self.bucket.zero()
        outs = self.model.forward_all(psn_out, priors_b1, priors_b2)
        loss = self.loss(outs, hr)
        loss.backward()
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.bucket.allreduce_mean(self.group)
        for m in self.modules + self.distill:                      # super_resolution.py:270-275
            torch.nn.utils.clip_grad_norm_(m.parameters(), self.clip)
        self.opt.step()
        return loss.detach()
