"""CMM gradient check at arbitrary sizes: our CUDA backward vs autograd through oracle/torch_ref.py (CPU).
python tools/cmm_grad_sweep.py B H W cnum train"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpmn_b200 import ComplementationModulationModule  # noqa: E402
from dpmn_b200.schema import cmm_schema  # noqa: E402
from oracle import torch_ref  # noqa: E402
from oracle.params import synth_params  # noqa: E402
from tests.util import rel_err  # noqa: E402

B, H, W, cnum, train = [int(v) for v in sys.argv[1:6]]
P = synth_params(cmm_schema(3, cnum), 7)
r = np.random.default_rng(5)
x1 = r.uniform(0, 1, (B, 3, H, W)).astype(np.float32)
x2 = r.uniform(0, 1, (B, 3, H, W)).astype(np.float32)
G = r.standard_normal((B, 3, H, W)).astype(np.float32)
Pt = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
for k, v in Pt.items():
    if v.dtype == torch.float32 and "running" not in k:
        v.requires_grad_(True)
a, b = torch.from_numpy(x1).requires_grad_(True), torch.from_numpy(x2).requires_grad_(True)
y = torch_ref.cmm_forward(Pt, a, b, training=bool(train))
(y * torch.from_numpy(G)).sum().backward()
m = ComplementationModulationModule(cnum=cnum, precision="fp32")
m.load_state_dict({k: torch.from_numpy(np.asarray(P[k])) for k in m.state_dict()}, strict=True)
m = m.cuda()
m.train(bool(train))
ca, cb = torch.from_numpy(x1).cuda().requires_grad_(True), torch.from_numpy(x2).cuda().requires_grad_(True)
yc = m(ca, cb)
(yc * torch.from_numpy(G).cuda()).sum().backward()
print("fwd", rel_err(yc.detach().cpu().numpy(), y.detach().numpy()))
bad = 0
for k, p in m.named_parameters():
    ref = Pt[k].grad.numpy()
    e = rel_err(p.grad.cpu().numpy(), ref)
    flag = "" if (e < 5e-4 or np.abs(ref).max() < 1e-2) else "  <-- BAD"
    bad += bool(flag)
    if flag or "-v" in sys.argv:
        print(f"{k:36s} ref_max {np.abs(ref).max():9.2e} rel {e:9.2e}{flag}")
print("x1", rel_err(ca.grad.cpu().numpy(), a.grad.numpy()), "x2", rel_err(cb.grad.cpu().numpy(), b.grad.numpy()), "bad", bad)
