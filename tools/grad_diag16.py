"""Per-tensor relative-L2 / cosine gradient errors of a 16-bit-mode CMM vs a gradient fixture (debugging aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import build_cmm, cmm_case, golden_grad_view, grad_seed_out, load_golden  # noqa: E402

name, prec = sys.argv[1], sys.argv[2]
z, meta = load_golden(name)
P, x1, x2 = cmm_case(meta)
m, _ = build_cmm(meta, "cuda", precision=prec)
dev = torch.device("cuda")
a = torch.from_numpy(x1).to(dev).requires_grad_(True)
b = torch.from_numpy(x2).to(dev).requires_grad_(True)
y = m(a, b)
(y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)).sum().backward()
grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
grads["x1"], grads["x2"] = a.grad.cpu().numpy(), b.grad.cpu().numpy()
for key in z.files:
    if not key.startswith("g:") or key[2:] not in grads:
        continue
    n = key[2:]
    full = meta["full"] or n in ("x1", "x2")
    want = z[key].astype(np.float64).ravel()
    got = golden_grad_view(grads[n], full).astype(np.float64).ravel()
    if np.abs(want).max() < 1e-2:
        continue
    l2 = np.linalg.norm(got - want) / np.linalg.norm(want)
    cos = np.dot(got, want) / (np.linalg.norm(got) * np.linalg.norm(want))
    print(f"{n:28s} l2 {l2:8.2e} cos {cos:.5f}")
