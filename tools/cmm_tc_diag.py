"""GPU diagnostic for the tensor-core CMM: per-layer error against the torch CPU port (run under gpurun)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpmn_b200 import _lib  # noqa: E402
from dpmn_b200.pgrm import workspace  # noqa: E402
from oracle import torch_ref  # noqa: E402
from tests.util import build_cmm, cmm_case, load_golden, rel_err  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")


def fetch(m, d, which, dtype, shape):
    n = lib.dpmn_cmm_debug_bytes(C.byref(d), which)
    buf = torch.empty(n, dtype=torch.uint8, device=dev)
    ws = workspace(dev, 0)
    rc = lib.dpmn_cmm_debug_copy(C.byref(d), ws.data_ptr(), which, buf.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return buf.view(dtype).view(*shape).float().cpu()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def run(name, prec):
    z, meta = load_golden(name)
    P, x1, x2 = cmm_case(meta)
    m, _ = build_cmm(meta, dev, precision=prec)
    tt = torch.float16 if prec == "fp16" else torch.bfloat16
    with torch.no_grad():
        y = m(torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev))
        torch.cuda.synchronize()
        Pt = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
        ref, parts = torch_ref.cmm_forward(Pt, torch.from_numpy(x1), torch.from_numpy(x2), training=False, return_parts=True)
    print(f"== {name} [{prec}] out vs golden: {rel_err(y.cpu().numpy(), z['out']):.3e}")
    B, _, H, W = x1.shape
    d = m._descriptor(B, H, W)
    c = meta["cnum"]
    Cl = {1: c, 2: 2 * c, 3: 4 * c, 4: 8 * c, 5: 8 * c}
    Co = {5: 8 * c, 4: 4 * c, 3: 2 * c, 2: c}
    for l in range(1, 6):
        h, w_ = H >> (l - 1), W >> (l - 1)
        got = fetch(m, d, l, tt, (2, B, h, w_, Cl[l]))
        want = torch.stack([nhwc(F.leaky_relu(parts[f"o{l}_{br}"], 0.2)) for br in (1, 2)])
        print(f"  e[{l}]    {rel_err(got.numpy(), want.numpy()):.3e}")
        if l >= 2:
            got = fetch(m, d, 10 + l, tt, (2, B, h, w_, Cl[l - 1]))
            want = torch.stack([nhwc(F.leaky_relu(parts[f"mid{l}_{br}"], 0.2)) for br in (1, 2)])
            print(f"  mid[{l}]  {rel_err(got.numpy(), want.numpy()):.3e}")
    got = fetch(m, d, 40, torch.float32, (2, B, (H >> 5) * (W >> 5), 8 * c))
    want = torch.stack([nhwc(parts[f"z6_{br}"]).reshape(B, -1, 8 * c) for br in (1, 2)])
    print(f"  z6      {rel_err(got.numpy(), want.numpy()):.3e}")
    got = fetch(m, d, 41, tt, (B, (H >> 5) * (W >> 5), 16 * c))
    print(f"  zg      {rel_err(got.numpy(), nhwc(F.relu(parts['zgate'])).reshape(B, -1, 16 * c).numpy()):.3e}")
    dprev = parts["d6"]
    for l in (5, 4, 3, 2, 1):
        h, w_ = H >> (l - 1), W >> (l - 1)
        cat = torch.cat([dprev, parts[f"o{l}_1"], parts[f"o{l}_2"]], dim=1)
        got = fetch(m, d, 20 + l, tt, (B, h, w_, cat.shape[1]))
        want = nhwc(F.relu(cat))
        nd = dprev.shape[1]
        print(f"  cat[{l}]  all {rel_err(got.numpy(), want.numpy()):.3e}  dec-part {rel_err(got[..., :nd].numpy(), want[..., :nd].numpy()):.3e}")
        if l >= 2:
            got = fetch(m, d, 30 + l, tt, (B, h, w_, Co[l]))
            print(f"  dmid[{l}] {rel_err(got.numpy(), nhwc(F.relu(parts[f'dmid{l}'])).numpy()):.3e}")
            dprev = parts[f"d{l}"]


if __name__ == "__main__":
    for name in ("cmm_c64_eval", "cmm_c8_eval"):
        for prec in ("fp16", "bf16"):
            run(name, prec)
