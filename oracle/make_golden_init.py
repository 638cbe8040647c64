"""Mint tests/golden/init_stats.json: per-parameter statistics of the UNMODIFIED reference's own initialisation
(PGRM._init_weights, model/pgrm.py:524-533, + torch defaults; cmm.py has no custom init), pooled over several constructions.

    python -m oracle.make_golden_init          (build container only: needs /root/reference)

For every state_dict entry: mean, std, min, max over `N_BUILDS` seeded constructions, and the element count.  The test
(tests/test_init_distributions.py) draws our modules' initial values the same number of times and compares the pooled
moments within sampling error -- initialisation parity is distributional, not bitwise.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.make_golden import OUT, load_reference

N_BUILDS = 4


def stats(tensors):
    x = np.concatenate([t.detach().double().reshape(-1).numpy() for t in tensors])
    return {"mean": float(x.mean()), "std": float(x.std()), "min": float(x.min()), "max": float(x.max()), "n": int(x.size)}


def main():
    pgrm_mod, cmm_mod = load_reference()
    out = {"n_builds": N_BUILDS, "pgrm_iter2_mode0": {}, "cmm_cnum16": {}}
    builds = []
    for s in range(N_BUILDS):
        torch.manual_seed(1000 + s)
        n = 3
        builds.append(pgrm_mod.PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                                    window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[0.1] * n,
                                    attn_drop_rate=[0.1] * n, drop_path_rate=[0.1] * n, iter=2, mode=False, hidden_size=3))
    for name, t in builds[0].state_dict().items():
        if t.dtype == torch.float32 and "attn_mask" not in name:
            out["pgrm_iter2_mode0"][name] = stats([b.state_dict()[name] for b in builds])
    builds = []
    for s in range(N_BUILDS):
        torch.manual_seed(2000 + s)
        builds.append(cmm_mod.ComplementationModulationModule(cnum=16))
    for name, t in builds[0].state_dict().items():
        if t.dtype == torch.float32:
            out["cmm_cnum16"][name] = stats([b.state_dict()[name] for b in builds])
    path = os.path.join(OUT, "init_stats.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, len(out["pgrm_iter2_mode0"]), len(out["cmm_cnum16"]))


if __name__ == "__main__":
    main()
