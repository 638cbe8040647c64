"""The index facts DESIGN.md section 9 rests on (fused Mlp kernels of the next round), checked against the oracle's
Mlp.forward restatement (pgrm.py:29-41, raw views of quirk 2) on random data -- CPU only:

  A. a 128-token tile of fc1's output is exactly 48 complete 32x32 planes of the raw (384, 32, 32) view, so
     fc1 + GELU + depthwise 3x3 + GELU can be computed tile by tile with no halo exchange;
  B. the fc2 input rows of tokens [128 j, 128 j + 128) are exactly the pointwise-conv outputs of channels
     [48 j, 48 j + 48) over all 1024 pixels, flat index 1024 * c_local + pixel -> (token, k) = divmod(flat, 384).
"""
import numpy as np

from oracle import pgrm_oracle as po


def _params(r, C=96, hid=384):
    pre = "mlp."
    return pre, {
        pre + "fc1.weight": r.normal(0, 0.1, (hid, C)).astype(np.float32), pre + "fc1.bias": r.normal(0, 0.1, hid).astype(np.float32),
        pre + "depthwise_conv.weight": r.normal(0, 0.3, (hid, 1, 3, 3)).astype(np.float32),
        pre + "depthwise_conv.bias": r.normal(0, 0.1, hid).astype(np.float32),
        pre + "pointwise_conv.weight": r.normal(0, 0.05, (hid, hid, 1, 1)).astype(np.float32),
        pre + "pointwise_conv.bias": r.normal(0, 0.1, hid).astype(np.float32),
        pre + "fc2.weight": r.normal(0, 0.05, (C, hid)).astype(np.float32), pre + "fc2.bias": r.normal(0, 0.1, C).astype(np.float32),
    }


def test_tilewise_mlp_equals_the_oracle():
    r = np.random.default_rng(0)
    L, C, hid, side, TILE = 1024, 96, 384, 32, 128
    pre, P = _params(r)
    x = r.normal(0, 1, (1, L, C)).astype(np.float32)
    want = po.mlp(x, P, pre)[0]                                                  # (L, C)

    planes_per_tile = TILE * hid // L                                            # 48
    assert planes_per_tile * L == TILE * hid and L // TILE == 8
    dw_w, dw_b = P[pre + "depthwise_conv.weight"][:, 0], P[pre + "depthwise_conv.bias"]
    # ---- kernel A, tile by tile: dt[pixel, c'] for c' in [48 t, 48 t + 48) from tokens [128 t, 128 t + 128) only
    dt = np.zeros((L, hid), np.float32)
    for t in range(L // TILE):
        rows = x[0, t * TILE:(t + 1) * TILE]                                     # (128, 96)
        h = po.gelu(rows @ P[pre + "fc1.weight"].T + P[pre + "fc1.bias"])        # (128, 384): the TMEM tile
        planes = h.reshape(planes_per_tile, side, side)                          # flat 384 m + j IS plane-major
        pad = np.pad(planes, ((0, 0), (1, 1), (1, 1)))
        for p in range(planes_per_tile):
            c = t * planes_per_tile + p
            acc = np.full((side, side), dw_b[c], np.float32)
            for ky in range(3):
                for kx in range(3):
                    acc += pad[p, ky:ky + side, kx:kx + side] * dw_w[c, ky, kx]
            dt[:, c] = po.gelu(acc).reshape(-1)                                  # pixel-major operand of the pointwise GEMM
    # ---- kernel B, CTA j of the image: P^T[pixel, c_out in 48 j .. 48 j + 48) -> fc2 rows of tokens [128 j, 128 j + 128)
    Wpw, bpw = P[pre + "pointwise_conv.weight"][:, :, 0, 0], P[pre + "pointwise_conv.bias"]
    out = np.zeros((L, C), np.float32)
    for j in range(L // TILE):
        cs = slice(planes_per_tile * j, planes_per_tile * (j + 1))
        pt = dt @ Wpw[cs].T + bpw[cs]                                            # (1024 pixels, 48): the TMEM tile, lane = pixel
        a_tile = np.zeros((TILE, hid), np.float32)
        for c_local in range(planes_per_tile):
            flat = L * c_local + np.arange(L)
            tok, k = np.divmod(flat, hid)
            a_tile[tok, k] = pt[:, c_local]
        assert not np.any(a_tile == 0.0)                                         # every (token, k) of the tile was written once
        out[j * TILE:(j + 1) * TILE] = a_tile @ P[pre + "fc2.weight"].T + P[pre + "fc2.bias"]
    assert np.abs(out - want).max() < 2e-5 * np.abs(want).max()


def test_tile_plane_boundaries_in_token_terms():
    """Plane c' = tokens [8 c' / 3, 8 (c' + 1) / 3): a row of the fc1 tile (one token, 384 values) spans at most two planes,
    and a 32-column tcgen05.ld chunk of a row is 64 contiguous bytes of one plane row pair."""
    hid, L = 384, 1024
    for m in range(128):
        first, last = (hid * m) // L, (hid * m + hid - 1) // L
        assert last - first <= 1
    for m in range(128):
        for j0 in range(0, hid, 32):
            flat = hid * m + j0
            assert flat // L == (flat + 31) // L                                 # a 32-value chunk never straddles two planes
            assert (flat % L) // 32 == ((flat + 31) % L) // 32                   # ... nor two rows of a plane
