"""`PGRM._init_weights` (model/pgrm.py:524-533: trunc_normal(.02) Linear weights and bias tables, zero
Linear biases, LayerNorm (1, 0), xavier_uniform conv weights, torch-default conv biases, ones weight_list_*) and the CMM's
torch defaults (kaiming_uniform(a=sqrt 5), U(+-1/sqrt(fan_in)) biases, BatchNorm (1, 0, 0, 1)).  Initialisation parity is
distributional: the pooled moments of our initial values are compared with those of the UNMODIFIED reference's own
constructions (tests/golden/init_stats.json, minted by oracle/make_golden_init.py) within sampling error.  CPU only."""
import json
import math
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_stats.json")


def _pooled(builds, name):
    return np.concatenate([b.state_dict()[name].detach().double().reshape(-1).numpy() for b in builds])


def _compare(ref, x, name):
    n = ref["n"]
    assert x.size == n, (name, x.size, n)
    if ref["std"] == 0.0:                                   # constants: zeros / ones / running stats
        assert float(x.std()) == 0.0 and float(x.mean()) == ref["mean"], name
        return
    # mean within 5 standard errors (of both samples), std within 5 relative standard errors (uniform and normal alike:
    # the relative standard error of a sample std is <= 1/sqrt(n) for these light-tailed laws)
    se = math.sqrt(2.0) * ref["std"] / math.sqrt(n)
    assert abs(float(x.mean()) - ref["mean"]) < 5 * se + 1e-12, (name, float(x.mean()), ref["mean"], se)
    assert abs(float(x.std()) / ref["std"] - 1.0) < 5 * math.sqrt(2.0 / n) + 1e-3, (name, float(x.std()), ref["std"])
    # support: uniform laws share their bounds, (trunc_)normal ones their +-2 truncation far in the tail
    span = ref["max"] - ref["min"]
    if n >= 500:
        assert float(x.min()) > ref["min"] - 0.25 * span and float(x.max()) < ref["max"] + 0.25 * span, name


def test_pgrm_initialisation_has_the_reference_distributions():
    from dpmn_b200 import PGRM
    gold = json.load(open(GOLD))
    builds = []
    for s in range(gold["n_builds"]):
        torch.manual_seed(5000 + s)
        n = 3
        builds.append(PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                           window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[0.1] * n, attn_drop_rate=[0.1] * n,
                           drop_path_rate=[0.1] * n, iter=2, mode=False, hidden_size=3))
    names = {k for k, v in builds[0].state_dict().items() if v.dtype == torch.float32 and "attn_mask" not in k}
    assert names == set(gold["pgrm_iter2_mode0"])
    for name in sorted(names):
        _compare(gold["pgrm_iter2_mode0"][name], _pooled(builds, name), name)
    # different seeds give different draws, the same seed the same draw (torch's generator, like the reference)
    torch.manual_seed(5000)
    again = PGRM(patch_size=[2] * 3, embed_dim=[96] * 3, depths=[1] * 3, num_heads=[[6]] * 3, window_size=[[2, 4, 8]] * 3,
                 mlp_ratio=[4.] * 3, drop_rate=[0.1] * 3, attn_drop_rate=[0.1] * 3, drop_path_rate=[0.1] * 3, iter=2,
                 mode=False, hidden_size=3)
    k = "layers.0.blocks.0.mlp.fc1.weight"
    assert torch.equal(again.state_dict()[k], builds[0].state_dict()[k])
    assert not torch.equal(builds[1].state_dict()[k], builds[0].state_dict()[k])


def test_cmm_initialisation_has_the_reference_distributions():
    from dpmn_b200 import ComplementationModulationModule
    gold = json.load(open(GOLD))
    builds = []
    for s in range(gold["n_builds"]):
        torch.manual_seed(6000 + s)
        builds.append(ComplementationModulationModule(cnum=16))
    names = {k for k, v in builds[0].state_dict().items() if v.dtype == torch.float32}
    assert names == set(gold["cmm_cnum16"])
    for name in sorted(names):
        _compare(gold["cmm_cnum16"][name], _pooled(builds, name), name)
