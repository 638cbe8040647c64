"""One data-parallel training step of the hot path as the reference runs it (interfaces/super_resolution.py:245-275):
seven image losses (six PGRM outputs + the CMM output), one backward, per-module gradient clipping at 0.25,
Adam(lr 1e-3, betas (0.5, 0.999)) (interfaces/base.py:208-221).  Forward and backward of every PGRM / CMM run in
libdpmn_b200 (dpmn_*_forward / dpmn_*_backward through the modules' autograd Functions); across ranks the ONLY
collective is one flat all-reduce over the gradient bucket.

Caller-side pieces kept in torch, exactly where the reference keeps them: the loss (ImageLoss, loss/image_loss.py:
15-43 -- SURVEY 8f rank 1, not yet a kernel), clip_grad_norm_ and the Adam update.
Not reproduced: the recogniser / text-rendering loop that produces the priors (out of scope: priors are inputs)
and the DistillModule terms."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F

from .dist import FlatGradBucket
from .pipeline import DPMNHotPath


def gradient_map(x: torch.Tensor) -> torch.Tensor:
    """GradientPriorLoss.gradient_map, loss/image_loss.py:33-43."""
    w = x.shape[-1]
    h = x.shape[-2]
    r = F.pad(x, (0, 1, 0, 0))[:, :, :, 1:]
    l = F.pad(x, (1, 0, 0, 0))[:, :, :, :w]
    t = F.pad(x, (0, 0, 1, 0))[:, :, :h, :]
    b = F.pad(x, (0, 0, 0, 1))[:, :, 1:, :]
    return torch.sqrt(((r - l) * 0.5) ** 2 + ((t - b) * 0.5) ** 2 + 1e-6)


def image_loss(out: torch.Tensor, target: torch.Tensor, weight=(1.0, 1.0)) -> torch.Tensor:
    """ImageLoss(gradient=True, loss_weight=[1, 1]) as main.py constructs it: MSE + L1 of the gradient maps."""
    return weight[0] * F.mse_loss(out, target) + weight[1] * F.l1_loss(gradient_map(out[:, :3]), gradient_map(target[:, :3]))


class HotPathTrainer:
    def __init__(self, model: DPMNHotPath, lr: float = 1e-3, betas=(0.5, 0.999), clip: float = 0.25, group=None):
        self.model, self.clip, self.group = model, clip, group
        self.modules = list(model.pgrm) + [model.cmm]
        self.bucket = FlatGradBucket(model.parameters())      # p.grad become views of ONE flat fp32 buffer
        self.opt = torch.optim.Adam(self.bucket.params, lr=lr, betas=betas)

    def loss(self, outs: Sequence[torch.Tensor], hr: torch.Tensor) -> torch.Tensor:
        total = outs[0].new_zeros(())
        for o in outs:
            total = total + image_loss(o, hr[:, :3]) * 100        # super_resolution.py:212,239,267
        return total / len(outs)                                   # :268

    def step(self, psn_out, priors_b1, priors_b2, hr) -> torch.Tensor:
        """forward + backward + all-reduce + clip + Adam; returns the (local) loss as a 0-d tensor."""
        self.bucket.zero()
        outs = self.model.forward_all(psn_out, priors_b1, priors_b2)
        loss = self.loss(outs, hr)
        loss.backward()
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.bucket.allreduce_mean(self.group)
        for m in self.modules:                                     # super_resolution.py:270-275
            torch.nn.utils.clip_grad_norm_(m.parameters(), self.clip)
        self.opt.step()
        return loss.detach()
