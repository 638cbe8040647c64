"""EXPERIMENTAL kernels (next-round drafts): skipped unless DPMN_EXPERIMENTAL=1.  Nothing here is on a default path.

dpmnx_mlp_fc1_dwconv (csrc/mlp_fused_a.cu): fc1 + GELU + depthwise 3x3 + GELU of Mlp.forward (pgrm.py:30-36) in one kernel,
checked against the numpy oracle on 16-bit-rounded operands (bar 2e-3 of the output maximum: fp16 storage of the hidden
tensor and of the result, sigmoid-form GELU)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DPMN_EXPERIMENTAL") != "1", reason="experimental draft: set DPMN_EXPERIMENTAL=1")]


@pytest.mark.parametrize("B", [1, 3])
def test_fused_fc1_dwconv_matches_oracle(B):
    from dpmn_b200 import _lib
    from oracle import pgrm_oracle as po
    lib = C.CDLL(_lib.LIB_PATH)
    fn = lib.dpmnx_mlp_fc1_dwconv
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    r = np.random.default_rng(B)
    L, Cc, hid, side = 1024, 96, 384, 32
    x = r.normal(0, 1, (B * L, Cc)).astype(np.float32)
    w1 = r.normal(0, 0.1, (hid, Cc)).astype(np.float32)
    b1 = r.normal(0, 0.1, hid).astype(np.float32)
    dw = r.normal(0, 0.3, (hid, 1, 3, 3)).astype(np.float32)
    db = r.normal(0, 0.1, hid).astype(np.float32)
    dev = torch.device("cuda")
    tx, tw, tb1, tdw, tdb = (torch.from_numpy(a).to(dev) for a in (x, w1, b1, dw, db))
    out = torch.empty(B, L, hid, device=dev)
    ws = torch.empty(2 * (x.size + w1.size + out.numel()) + 4096, dtype=torch.uint8, device=dev)
    rc = fn(tx.data_ptr(), tw.data_ptr(), tb1.data_ptr(), tdw.data_ptr(), tdb.data_ptr(), out.data_ptr(), B, 1,
            ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    # oracle on fp16-rounded operands
    x16 = x.astype(np.float16).astype(np.float32).reshape(B, L, Cc)
    w16 = w1.astype(np.float16).astype(np.float32)
    h = po.gelu(x16 @ w16.T + b1).astype(np.float16).astype(np.float32)              # (B, L, hid), stored 16-bit
    planes = h.reshape(B, hid, side, side)                                           # raw view (quirk 2)
    y = po.gelu(po.conv2d(planes, dw, db, pad=1, groups=hid))                        # (B, hid, 32, 32)
    want = y.reshape(B, hid, L).transpose(0, 2, 1)                                   # pixel-major (B, L, hid)
    got = out.cpu().numpy()
    assert np.abs(got - want).max() < 2e-3 * np.abs(want).max(), float(np.abs(got - want).max())
