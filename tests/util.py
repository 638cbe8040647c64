"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from dpmn_b200.schema import PGRMConfig, cmm_schema, pgrm_schema
from oracle import inputs as gen
from oracle.params import synth_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, ref):
    """BASELINE.md tolerance metric: max|a - ref| / max|ref|."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def pgrm_case(meta):
    """(cfg, params, x_q, x_kv, residuals) rebuilt from a golden fixture's seeds."""
    cfg = PGRMConfig(embed_dim=meta["embed"], num_heads=meta["heads"], window_size=tuple(meta["window"]),
                     iter=meta["iter"], mode=meta["mode"])
    P = synth_params(pgrm_schema(cfg), meta["seed"])
    B = meta["B"]
    x_q = gen.prior_branch2(meta["seed"], B) if meta["mode"] else gen.prior_branch1(meta["seed"], B)
    x_kv = gen.image_stream(meta["seed"], B)
    res = gen.residuals(meta["seed"], B, meta["nres"])
    return cfg, P, x_q, x_kv, res


def cmm_case(meta):
    P = synth_params(cmm_schema(3, meta["cnum"]), meta["seed"])
    x1 = gen.image_stream(meta["seed"], meta["B"], tag=31)
    x2 = gen.image_stream(meta["seed"], meta["B"], tag=32)
    return P, x1, x2


PGRM_GOLDEN = ["pgrm_i0_m0", "pgrm_i2_m0", "pgrm_i3_m1", "pgrm_i5_m1", "pgrm_w2", "pgrm_w4", "pgrm_w8",
               "pgrm_w16_c192", "pgrm_w48_c96_h4"]
CMM_GOLDEN = ["cmm_c8_eval", "cmm_c8_train", "cmm_c64_eval"]


# ---- builders of OUR modules from a golden fixture's meta (GPU tests, bench, smoke) --------------------
def build_pgrm(meta, device="cuda", precision="fp32"):
    import torch
    from dpmn_b200 import PGRM
    it = meta["iter"]
    n = it + 1
    m = PGRM(patch_size=[2] * n, embed_dim=[meta["embed"]] * n, depths=[1] * n, num_heads=[[meta["heads"]]] * n,
             window_size=[list(meta["window"])] * n, mlp_ratio=[4.] * n, drop_rate=[0.1] * n,
             attn_drop_rate=[0.1] * n, drop_path_rate=[0.1] * n, iter=it, mode=meta["mode"], hidden_size=3,
             precision=precision)
    cfg = PGRMConfig(embed_dim=meta["embed"], num_heads=meta["heads"], window_size=tuple(meta["window"]),
                     iter=it, mode=meta["mode"])
    P = synth_params(pgrm_schema(cfg), meta["seed"])
    sd = m.state_dict()
    m.load_state_dict({k: (torch.from_numpy(P[k]) if k in P else v) for k, v in sd.items()}, strict=True)
    return m.to(device).eval(), P


def build_cmm(meta, device="cuda", precision="fp32"):
    import torch
    from dpmn_b200 import ComplementationModulationModule
    m = ComplementationModulationModule(cnum=meta["cnum"], precision=precision)
    P = synth_params(cmm_schema(3, meta["cnum"]), meta["seed"])
    sd = m.state_dict()
    assert set(sd) == set(P)
    m.load_state_dict({k: torch.from_numpy(np.asarray(P[k])) for k in sd}, strict=True)
    m = m.to(device)
    m.train(bool(meta.get("train", False)))
    return m, P


# ---- gradient fixtures (oracle/make_golden_grads.py) -------------------------------------------------------
PGRM_GRAD_GOLDEN = ["pgrm_i2_m0_grad", "pgrm_i3_m1_grad", "pgrm_i5_m1_grad"]
CMM_GRAD_GOLDEN = ["cmm_c8_train_grad", "cmm_c8_eval_grad", "cmm_c16_train_grad"]
GRAD_SAMPLE = 512


def grad_seed_out(seed, B):
    """d loss / d out of the gradient fixtures: loss = sum(out * G), G ~ N(0,1) seeded."""
    return np.random.default_rng([seed, 77]).standard_normal((B, 3, 32, 128)).astype(np.float32)


def golden_grad_view(g, full):
    """The part of a gradient tensor a fixture stores (everything, or a strided sample of <= 512 entries)."""
    g = np.asarray(g, dtype=np.float32)
    if full:
        return g
    step = max(1, g.size // GRAD_SAMPLE)
    return g.reshape(-1)[np.arange(0, g.size, step)[:GRAD_SAMPLE]]


def torch_ref_pgrm_grads(meta):
    """Gradients of the torch-CPU oracle (autograd through oracle/torch_ref.py) for a gradient fixture."""
    import torch
    from oracle import torch_ref
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    Pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in P.items()}
    xkv = torch.from_numpy(x_kv).requires_grad_(True)
    rs = [torch.from_numpy(r).requires_grad_(True) for r in res]
    y = torch_ref.pgrm_forward(Pt, torch.from_numpy(x_q), xkv, rs, windows=cfg.window_size, num_heads=cfg.num_heads)
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"]))).sum().backward()
    out = {k: (v.grad.numpy() if v.grad is not None else None) for k, v in Pt.items()}
    out["x_kv"] = xkv.grad.numpy()
    for i, r in enumerate(rs):
        out[f"res{i}"] = r.grad.numpy() if r.grad is not None else None
    return y.detach().numpy(), out


def torch_ref_cmm_grads(meta):
    import torch
    from oracle import torch_ref
    P, x1, x2 = cmm_case(meta)
    Pt = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
    for k, v in Pt.items():
        if v.dtype == torch.float32 and "running" not in k:
            v.requires_grad_(True)
    a, b = torch.from_numpy(x1).requires_grad_(True), torch.from_numpy(x2).requires_grad_(True)
    y = torch_ref.cmm_forward(Pt, a, b, training=bool(meta["train"]))
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"]))).sum().backward()
    out = {k: v.grad.numpy() for k, v in Pt.items() if v.requires_grad and v.grad is not None}
    out["x1"], out["x2"] = a.grad.numpy(), b.grad.numpy()
    return y.detach().numpy(), out
