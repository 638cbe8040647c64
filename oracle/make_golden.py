"""Mint the golden fixtures under tests/golden/ from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python -m oracle.make_golden

`model/pgrm.py` imports three symbols from timm (pgrm.py:10), which is not installed here; they are
shimmed below with their timm 0.6.5 semantics (DropPath is identity in eval(), which is all the
fixtures exercise).  `model/cmm.py` imports as is.  Weights are the deterministic synthetic values of
oracle/params.py, inputs the generators of oracle/inputs.py, so each fixture stores only seeds,
configuration and reference OUTPUTS (fp32).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from dpmn_b200.schema import PGRMConfig, cmm_schema, pgrm_schema  # noqa: E402
from oracle import inputs as gen  # noqa: E402
from oracle.params import synth_params  # noqa: E402


def _install_timm_shim():
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load_reference():
    _install_timm_shim()
    mods = {}
    for name in ("pgrm", "cmm"):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, "model", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["pgrm"], mods["cmm"]


def _load_synth(module: torch.nn.Module, schema, seed):
    """Check the reference's state_dict against our schema, then load synthetic values."""
    sd = module.state_dict()
    want = {n: tuple(s) for n, s, _ in schema}
    got = {k: tuple(v.shape) for k, v in sd.items()}
    assert want == got, f"schema mismatch: {set(want) ^ set(got)} / " \
                        f"{[(k, want[k], got[k]) for k in want if k in got and want[k] != got[k]]}"
    vals = synth_params(schema, seed)
    new = {k: torch.from_numpy(vals[k]) if k in vals else v for k, v in sd.items()}
    module.load_state_dict(new, strict=True)
    return vals


PGRM_CASES = [
    # name, ctor kwargs, batch, #residuals, seed
    dict(name="pgrm_i0_m0", iter=0, mode=False, B=2, nres=0, seed=11),
    dict(name="pgrm_i2_m0", iter=2, mode=False, B=2, nres=2, seed=12),
    dict(name="pgrm_i3_m1", iter=3, mode=True, B=2, nres=0, seed=13),
    dict(name="pgrm_i5_m1", iter=5, mode=True, B=1, nres=3, seed=14),
    dict(name="pgrm_w2", iter=0, mode=True, B=2, nres=0, seed=15, window=(2,)),
    dict(name="pgrm_w4", iter=0, mode=True, B=1, nres=0, seed=16, window=(4,)),
    dict(name="pgrm_w8", iter=0, mode=True, B=1, nres=0, seed=17, window=(8,)),
    dict(name="pgrm_w16_c192", iter=0, mode=True, B=1, nres=0, seed=18, window=(16,), embed=192),
    dict(name="pgrm_w48_c96_h4", iter=0, mode=True, B=1, nres=0, seed=19, window=(4, 8), heads=4),
]


def make_pgrm(pgrm_mod):
    for case in PGRM_CASES:
        window = tuple(case.get("window", (2, 4, 8)))
        embed = case.get("embed", 96)
        heads = case.get("heads", 6)
        it = case["iter"]
        n = it + 1
        torch.manual_seed(0)
        m = pgrm_mod.PGRM(patch_size=[2] * n, embed_dim=[embed] * n, depths=[1] * n, num_heads=[[heads]] * n,
                          window_size=[list(window)] * n, mlp_ratio=[4.] * n, drop_rate=[0.1] * n,
                          attn_drop_rate=[0.1] * n, drop_path_rate=[0.1] * n, iter=it, mode=case["mode"],
                          hidden_size=3)
        cfg = PGRMConfig(embed_dim=embed, num_heads=heads, window_size=window, iter=it, mode=case["mode"])
        _load_synth(m, pgrm_schema(cfg), case["seed"])
        m.eval()
        B = case["B"]
        x_q = gen.prior_branch2(case["seed"], B) if case["mode"] else gen.prior_branch1(case["seed"], B)
        x_kv = gen.image_stream(case["seed"], B)
        res = gen.residuals(case["seed"], B, case["nres"])
        captured = {}
        hooks = []
        for b, blk in enumerate(m.layers[0].blocks):
            hooks.append(blk.attn.sknet.register_forward_pre_hook(
                lambda mod, args, b=b: captured.__setitem__(f"attn_core_b{b}", args[0].detach().numpy().copy())))
            hooks.append(blk.register_forward_hook(
                lambda mod, args, out, b=b: captured.__setitem__(f"block{b}_out", out[1].detach().numpy().copy())))
        with torch.no_grad():
            y = m(torch.from_numpy(x_q), torch.from_numpy(x_kv), [torch.from_numpy(r) for r in res])
        for h in hooks:
            h.remove()
        save = {"out": y.numpy()}
        if case["name"] in ("pgrm_i0_m0", "pgrm_w16_c192"):
            # per-stage probes: attention core (pre-SK, window-major) and block outputs, image 0 only
            for k, v in captured.items():
                save[k] = v.reshape(B, -1, embed)[:1].astype(np.float32)
        meta = dict(case)
        meta["window"] = list(window)
        meta["embed"] = embed
        meta["heads"] = heads
        save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **save)
        print("wrote", case["name"], y.shape, float(y.abs().max()))

    # reference-built buffers for the production window list (index closed form + shift masks)
    torch.manual_seed(0)
    m = pgrm_mod.PGRM(patch_size=[2], embed_dim=[96], depths=[1], num_heads=[[6]], window_size=[[2, 4, 8]],
                      mlp_ratio=[4.], drop_rate=[0.], attn_drop_rate=[0.], drop_path_rate=[0.], iter=0, mode=True,
                      hidden_size=3)
    sd = m.state_dict()
    bufs = {}
    for g in range(3):
        bufs[f"index_{g}"] = sd[f"layers.0.blocks.0.attn.relative_position_index_{g}"].numpy().astype(np.int16)
        bufs[f"mask_{g}"] = (sd[f"layers.0.blocks.1.attn.attn_mask_{g}"].numpy() != 0).astype(np.uint8)
        assert set(np.unique(sd[f"layers.0.blocks.1.attn.attn_mask_{g}"].numpy())) <= {0.0, -100.0}
    np.savez_compressed(os.path.join(OUT, "pgrm_buffers_248.npz"), **bufs)


def make_schema_dump(pgrm_mod, cmm_mod):
    dump = {}
    for it, mode in ((0, False), (2, False), (5, True)):
        n = it + 1
        m = pgrm_mod.PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                          window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[0.] * n,
                          attn_drop_rate=[0.] * n, drop_path_rate=[0.] * n, iter=it, mode=mode, hidden_size=3)
        dump[f"pgrm_iter{it}_mode{int(mode)}"] = {k: list(v.shape) for k, v in m.state_dict().items()}
        dump[f"pgrm_iter{it}_mode{int(mode)}_nparams"] = sum(p.numel() for p in m.parameters())
    c = cmm_mod.ComplementationModulationModule()
    dump["cmm_cnum64"] = {k: list(v.shape) for k, v in c.state_dict().items()}
    dump["cmm_cnum64_nparams"] = sum(p.numel() for p in c.parameters())
    with open(os.path.join(OUT, "reference_state_dict_schema.json"), "w") as f:
        json.dump(dump, f, indent=0, sort_keys=True)


CMM_CASES = [
    dict(name="cmm_c8_eval", cnum=8, B=2, train=False, seed=21),
    dict(name="cmm_c8_train", cnum=8, B=3, train=True, seed=22),
    dict(name="cmm_c64_eval", cnum=64, B=1, train=False, seed=23),
]


def make_cmm(cmm_mod):
    for case in CMM_CASES:
        torch.manual_seed(0)
        m = cmm_mod.ComplementationModulationModule(cnum=case["cnum"])
        _load_synth(m, cmm_schema(3, case["cnum"]), case["seed"])
        m.train(case["train"])
        x1 = gen.image_stream(case["seed"], case["B"], tag=31)
        x2 = gen.image_stream(case["seed"], case["B"], tag=32)
        with torch.no_grad():
            y = m(torch.from_numpy(x1), torch.from_numpy(x2))
        save = {"out": y.numpy(), "meta": np.frombuffer(json.dumps(case).encode(), dtype=np.uint8)}
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **save)
        print("wrote", case["name"], y.shape, float(y.abs().max()))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    pgrm_mod, cmm_mod = load_reference()
    make_schema_dump(pgrm_mod, cmm_mod)
    make_pgrm(pgrm_mod)
    make_cmm(cmm_mod)


if __name__ == "__main__":
    main()
