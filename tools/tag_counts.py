"""Kernel-class launch counts and times of one PGRM forward+backward (profile hook), for debugging."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dpmn_b200 import _lib  # noqa: E402
from dpmn_b200.pipeline import DPMNHotPath  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
model = DPMNHotPath(precision="fp16", drop=0.1, cmm_precision="fp32")
pg, cm = bench.synth_weights(2)
bench.load_weights(model, pg, cm)
model = model.to(dev).train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
psn, p1, p2 = bench.synth_inputs(1, B)
m = model.pgrm[0]
xq = torch.from_numpy(p1[0]).to(dev)
xkv = torch.from_numpy(psn[:, :3].copy()).to(dev).requires_grad_(True)
for it in range(2):
    if it == 1:
        lib.dpmn_profile_enable(1)
    y = m(xq, xkv, [])
    y.sum().backward()
    torch.cuda.synchronize()
cap = 20000
tags, nk, ms = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_float * cap)()
n = lib.dpmn_profile_collect(tags, nk, ms, cap)
agg = {}
for j in range(n):
    name = lib.dpmn_profile_tag_name(tags[j]).decode()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += nk[j]
    a[1] += ms[j]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:18s} kernels {v[0]:4d}  {v[1]:8.3f} ms")
