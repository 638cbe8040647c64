"""The DPMN hot path as its caller runs it: two cascades of three PGRMs over the frozen PSN output, then
the Complementation Modulation Module (interfaces/super_resolution.py:174-265 train, :418-448 eval,
:674-704 test).  Host-side glue only: every module call goes to libdpmn_b200 through the C-ABI."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from .cmm import ComplementationModulationModule
from .pgrm import PGRM

# README.md:42 of the reference: the flags every DPMN run passes (stu_iter_b1 = stu_iter_b2 = 3)
N_ITER = 6


def build_pgrm_stack(precision: str = "fp32", stu_iter_b1: int = 3, stu_iter_b2: int = 3,
                     drop: float = 0.1) -> List[PGRM]:
    """generator_init (interfaces/base.py:150-155) for k = 0..5: branch 1 gets the 2-channel rendered-text
    prior (mode=False), branch 2 the 3-channel mask prior (mode=True); hidden_size = 3.  `drop` is the value of
    --drop_rate / --attn_drop_rate / --drop_path_rate (README.md:42: 0.1); the stochastic train-mode paths are
    not implemented, so a stack that is to be put in train() must be built with drop=0."""
    n = stu_iter_b1 + stu_iter_b2
    mods = []
    for k in range(n):
        mods.append(PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                         window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[drop] * n,
                         attn_drop_rate=[drop] * n, drop_path_rate=[drop] * n, iter=k, mode=(k >= stu_iter_b1),
                         hidden_size=3, precision=precision))
    return mods


class DPMNHotPath(nn.Module):
    def __init__(self, precision: str = "fp32", stu_iter_b1: int = 3, stu_iter_b2: int = 3, drop: float = 0.1,
                 cmm_precision: str = None):
        super().__init__()
        self.b1, self.b2 = stu_iter_b1, stu_iter_b2
        self.pgrm = nn.ModuleList(build_pgrm_stack(precision, stu_iter_b1, stu_iter_b2, drop))
        self.cmm = ComplementationModulationModule(precision=cmm_precision or precision)

    def forward(self, psn_out: torch.Tensor, priors_b1: Sequence[torch.Tensor], priors_b2: Sequence[torch.Tensor]):
        """psn_out (B,4,32,128) frozen-backbone output; priors_b1[k] (B,2,32,128) rendered-text maps;
        priors_b2[k] (B,3,32,128) binary masks -> fused SR image (B,3,32,128)."""
        return self.forward_all(psn_out, priors_b1, priors_b2)[-1]

    def forward_all(self, psn_out, priors_b1, priors_b2) -> List[torch.Tensor]:
        """All seven images the training loss looks at (super_resolution.py:212,239,267): the six PGRM outputs in
        call order, then the CMM output."""
        cascade = psn_out[:, :3, :]                       # channel-slice view, super_resolution.py:196
        done: List[torch.Tensor] = []
        for k in range(self.b1):
            y = self.pgrm[k](priors_b1[k], cascade, done[:k])          # :207
            done.append(y)
            cascade = y
        sr1 = done[-1]
        outs = list(done)
        cascade = psn_out[:, :3, :]
        done = []
        for k in range(self.b1, self.b1 + self.b2):
            y = self.pgrm[k](priors_b2[k - self.b1], cascade, done[:k - self.b2])   # :234
            done.append(y)
            cascade = y
        outs += done
        outs.append(self.cmm(sr1, done[-1]))                           # :265
        return outs
