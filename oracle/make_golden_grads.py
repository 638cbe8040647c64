"""Mint GRADIENT golden fixtures (tests/golden/*_grad.npz) from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_grads

Scalar loss = sum(out * G) with G ~ N(0,1) seeded (so d loss / d out = G); the fixture stores the
reference's autograd gradients w.r.t. every parameter, x_kv and the residual inputs (PGRM) or x1, x2 (CMM).
`full` cases store every gradient tensor; `sampled` cases store a strided sample of <= 512 entries per
parameter plus the full input gradients (keeps tests/golden small).  SURVEY.md 8c item (v).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from dpmn_b200.schema import PGRMConfig, cmm_schema, pgrm_schema  # noqa: E402
from oracle import inputs as gen  # noqa: E402
from oracle.make_golden import _load_synth, load_reference  # noqa: E402

SAMPLE = 512


def d_out(seed, B):
    return np.random.default_rng([seed, 77]).standard_normal((B, 3, 32, 128)).astype(np.float32)


def sample_idx(n):
    step = max(1, n // SAMPLE)
    return np.arange(0, n, step)[:SAMPLE]


def _pack(save, name, g, full):
    g = np.zeros(1, np.float32) if g is None else g.detach().numpy().astype(np.float32)
    save["g:" + name] = g if full else g.reshape(-1)[sample_idx(g.size)]


PGRM_GRAD_CASES = [
    dict(name="pgrm_i2_m0_grad", iter=2, mode=False, B=2, nres=2, seed=41, full=True),
    dict(name="pgrm_i3_m1_grad", iter=3, mode=True, B=1, nres=0, seed=42, full=False),
    dict(name="pgrm_i5_m1_grad", iter=5, mode=True, B=1, nres=3, seed=43, full=False),
]
CMM_GRAD_CASES = [
    dict(name="cmm_c8_train_grad", cnum=8, B=2, train=True, seed=51, full=True),
    dict(name="cmm_c8_eval_grad", cnum=8, B=2, train=False, seed=52, full=False),
    dict(name="cmm_c16_train_grad", cnum=16, B=1, train=True, seed=53, full=False),
]
MARGIN = 3e-6   # min |pre-activation| over every ReLU / LeakyReLU input of a CMM gradient fixture


def make_pgrm(pgrm_mod):
    for case in PGRM_GRAD_CASES:
        it = case["iter"]
        n = it + 1
        torch.manual_seed(0)
        m = pgrm_mod.PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                          window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[0.] * n,
                          attn_drop_rate=[0.] * n, drop_path_rate=[0.] * n, iter=it, mode=case["mode"], hidden_size=3)
        cfg = PGRMConfig(embed_dim=96, num_heads=6, window_size=(2, 4, 8), iter=it, mode=case["mode"])
        _load_synth(m, pgrm_schema(cfg), case["seed"])
        m.eval()   # drop rates are 0: eval == train for this module
        B = case["B"]
        x_q = torch.from_numpy(gen.prior_branch2(case["seed"], B) if case["mode"] else gen.prior_branch1(case["seed"], B))
        x_kv = torch.from_numpy(gen.image_stream(case["seed"], B)).requires_grad_(True)
        res = [torch.from_numpy(r).requires_grad_(True) for r in gen.residuals(case["seed"], B, case["nres"])]
        y = m(x_q, x_kv, res)
        (y * torch.from_numpy(d_out(case["seed"], B))).sum().backward()
        save = {"out": y.detach().numpy()}
        for k, p in m.named_parameters():
            _pack(save, k, p.grad, case["full"])
        _pack(save, "x_kv", x_kv.grad, True)
        for i, r in enumerate(res):
            _pack(save, f"res{i}", r.grad, True)
        meta = dict(case, window=[2, 4, 8], embed=96, heads=6)
        save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **save)
        print("wrote", case["name"])


def _cmm_run(cmm_mod, case, dtype):
    torch.manual_seed(0)
    m = cmm_mod.ComplementationModulationModule(cnum=case["cnum"])
    _load_synth(m, cmm_schema(3, case["cnum"]), case["seed"])
    m = m.to(dtype)
    m.train(case["train"])
    x1 = torch.from_numpy(gen.image_stream(case["seed"], case["B"], tag=31)).to(dtype).requires_grad_(True)
    x2 = torch.from_numpy(gen.image_stream(case["seed"], case["B"], tag=32)).to(dtype).requires_grad_(True)
    y = m(x1, x2)
    (y * torch.from_numpy(d_out(case["seed"], case["B"])).to(dtype)).sum().backward()
    return m, x1, x2, y


def _cmm_is_stable(cmm_mod, case):
    """ReLU / LeakyReLU make the gradient a discontinuous function of the pre-activations: one element within
    rounding distance of 0 flips its mask and moves every upstream gradient by ~1e-2 (torch's own fp32 and fp64
    backward then disagree by that much).  A fixture is only usable as a parity target where the reference is
    unambiguous, so a seed is skipped when any ReLU / LeakyReLU input of the reference forward lies within MARGIN
    of 0 (a different but equally valid fp32 rounding could flip it) or when the reference's own fp32 and fp64
    gradients differ by more than 1e-5."""
    margin = [float("inf")]

    def pre(mod, args):
        margin[0] = min(margin[0], args[0].detach().abs().min().item())

    def post(mod, args, out):
        margin[0] = min(margin[0], out.detach().abs().min().item())
    torch.manual_seed(0)
    probe = cmm_mod.ComplementationModulationModule(cnum=case["cnum"])
    _load_synth(probe, cmm_schema(3, case["cnum"]), case["seed"])
    probe.train(case["train"])
    for mod in probe.modules():
        if isinstance(mod, (torch.nn.ReLU, torch.nn.LeakyReLU)):
            mod.register_forward_pre_hook(pre)
    probe.fc_1.register_forward_hook(post)     # the SE gate's inline ReLU (cmm.py:141)
    with torch.no_grad():
        probe(torch.from_numpy(gen.image_stream(case["seed"], case["B"], tag=31)),
              torch.from_numpy(gen.image_stream(case["seed"], case["B"], tag=32)))
    if margin[0] < MARGIN:
        return False, margin[0]
    m32, a32, b32, _ = _cmm_run(cmm_mod, case, torch.float32)
    m64, a64, b64, _ = _cmm_run(cmm_mod, case, torch.float64)
    worst = 0.0
    for (k, p32), (_, p64) in zip(m32.named_parameters(), m64.named_parameters()):
        ref = p64.grad.abs().max().item()
        if ref < 1e-2:
            continue
        worst = max(worst, (p32.grad.double() - p64.grad).abs().max().item() / ref)
    return worst < 1e-5, worst


def make_cmm(cmm_mod):
    for case in CMM_GRAD_CASES:
        case = dict(case)
        while True:
            ok, worst = _cmm_is_stable(cmm_mod, case)
            print(case["name"], "seed", case["seed"], "margin / fp32-vs-fp64 gap", f"{worst:.2e}", "ok" if ok else "SKIP", flush=True)
            if ok:
                break
            case["seed"] += 1
        m, x1, x2, y = _cmm_run(cmm_mod, case, torch.float32)
        save = {"out": y.detach().numpy()}
        for k, p in m.named_parameters():
            _pack(save, k, p.grad, case["full"])
        _pack(save, "x1", x1.grad, True)
        _pack(save, "x2", x2.grad, True)
        save["meta"] = np.frombuffer(json.dumps(case).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **save)
        print("wrote", case["name"])


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    pgrm_mod, cmm_mod = load_reference()
    make_pgrm(pgrm_mod)
    make_cmm(cmm_mod)


if __name__ == "__main__":
    main()
