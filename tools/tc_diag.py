"""GPU diagnostic for the tcgen05 GEMM (run under gpurun): error per shape, and where a wrong tile is wrong."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpmn_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
rng = np.random.default_rng(0)


def run(M, N, K, prec, pattern="rand"):
    if pattern == "rand":
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = rng.standard_normal((N, K)).astype(np.float32)
    else:   # structured: exact in 16-bit, exposes row/column/k permutations
        A = (np.arange(M)[:, None] % 7 + 1.0) * ((np.arange(K)[None, :] % 5) == 0) + (np.arange(K)[None, :] % 3) * 0.25
        B = (np.arange(N)[:, None] % 11 + 1.0) * ((np.arange(K)[None, :] % 4) == 1) + (np.arange(K)[None, :] % 2) * 0.5
        A, B = A.astype(np.float32), B.astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    tt = torch.float16 if prec == 1 else torch.bfloat16
    a, b, bi = (torch.from_numpy(x).to(dev) for x in (A, B, bias))
    c = torch.full((M, N), float("nan"), device=dev)
    nbytes = lib.dpmn_gemm_nt_workspace_bytes(M, N, K, prec)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = lib.dpmn_gemm_nt(a.data_ptr(), b.data_ptr(), bi.data_ptr(), c.data_ptr(), M, N, K, prec, ws.data_ptr(), nbytes,
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = a.to(tt).double() @ b.to(tt).double().T + bi.double()
    err = (c.double() - ref).abs()
    rel = float(err.max() / ref.abs().max())
    print(f"M={M} N={N} K={K} prec={prec} {pattern}: rc={rc} rel={rel:.3e} nan={int(torch.isnan(c).sum())}", flush=True)
    if not (rel < 1e-4):
        e = err.cpu().numpy()
        bad = np.argwhere(~(e < 1e-3 * float(ref.abs().max())))
        print("   first bad (m, n):", bad[:8].tolist(), " bad rows%8:", np.bincount(bad[:, 0] % 8, minlength=8).tolist(),
              " bad cols%8:", np.bincount(bad[:, 1] % 8, minlength=8).tolist())
        print("   got ", c[:2, :6].cpu().numpy().round(3).tolist())
        print("   want", ref[:2, :6].cpu().numpy().round(3).tolist())
    return rel


if __name__ == "__main__":
    ok = True
    for prec in (1, 2):
        for (M, N, K) in ((128, 64, 64), (128, 96, 96), (256, 192, 96), (1024, 384, 96), (384, 1024, 384), (200, 96, 384),
                          (77, 200, 48), (49152, 96, 96)):
            for pat in ("struct", "rand"):
                ok &= run(M, N, K, prec, pat) < 1e-4
    print("TC_DIAG", "PASS" if ok else "FAIL")
