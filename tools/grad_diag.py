"""Per-tensor gradient error table for a gradient fixture (debugging aid): python tools/grad_diag.py cmm_c8_train_grad"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import (build_cmm, build_pgrm, cmm_case, golden_grad_view, grad_seed_out, load_golden, pgrm_case,  # noqa: E402
                        rel_err)

name = sys.argv[1]
z, meta = load_golden(name)
dev = torch.device("cuda")
G = torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)
if name.startswith("cmm"):
    P, x1, x2 = cmm_case(meta)
    m, _ = build_cmm(meta, "cuda", precision="fp32")
    a = torch.from_numpy(x1).to(dev).requires_grad_(True)
    b = torch.from_numpy(x2).to(dev).requires_grad_(True)
    (m(a, b) * G).sum().backward()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    grads["x1"], grads["x2"] = a.grad.cpu().numpy(), b.grad.cpu().numpy()
else:
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, "cuda", precision="fp32")
    xkv = torch.from_numpy(x_kv).to(dev).requires_grad_(True)
    rs = [torch.from_numpy(r).to(dev).requires_grad_(True) for r in res]
    (m(torch.from_numpy(x_q).to(dev), xkv, rs) * G).sum().backward()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    grads["x_kv"] = xkv.grad.cpu().numpy()
for key in z.files:
    if not key.startswith("g:"):
        continue
    n = key[2:]
    want = z[key]
    got = grads.get(n)
    if got is None:
        print(f"{n:40s} (none)")
        continue
    full = meta["full"] or n in ("x_kv", "x1", "x2") or n.startswith("res")
    got = golden_grad_view(got, full)
    print(f"{n:40s} ref_max {np.abs(want).max():10.3e} got_max {np.abs(got).max():10.3e} rel {rel_err(got, want):9.2e}")
