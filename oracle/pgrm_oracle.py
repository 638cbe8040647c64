"""numpy restatement of the reference PGRM forward (eval mode).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/model/pgrm.py; each function cites the lines it restates.  Parameters are a
flat dict {state_dict key: ndarray} exactly as `PGRM.state_dict()` names them.  Parity: pinned against
tests/golden/*.npz (minted from the unmodified reference by oracle/make_golden.py).

The five behaviours a "clean Swin" would get wrong are kept on purpose:
  1. attention output stays in window-major token order (pgrm.py:249,263; window_reverse result unused)
  2. Mlp reinterprets (B, L, hidden) memory as (B, hidden, sqrt(L), sqrt(L)) without transposing (pgrm.py:33-38)
  3. residual_list[0] is skipped by the final affine mix (pgrm.py:563)
  4. shift mask value is -100.0, not -inf (pgrm.py:173)
  5. no output projection; an SK gate with a per-image global average pool follows attention (pgrm.py:79-96)
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

try:  # scipy is in the image; keep a slow exact fallback so the oracle never silently approximates
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


def gelu(x: np.ndarray) -> np.ndarray:
    """nn.GELU() default = exact erf form (pgrm.py:17,63)."""
    return (0.5 * x * (1.0 + _erf(x * (1.0 / math.sqrt(2.0))))).astype(x.dtype)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm over the last axis, biased variance (pgrm.py:303-304,311,415)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return ((x - mu) / np.sqrt(var + eps) * w + b).astype(x.dtype)


def linear(x: np.ndarray, w: np.ndarray, b: np.ndarray | None) -> np.ndarray:
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(x.dtype)


def conv2d(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, stride: int = 1, pad: int = 0,
           dil: int = 1, groups: int = 1) -> np.ndarray:
    """Direct NCHW cross-correlation (torch Conv2d semantics), tap by tap."""
    B, C, H, W = x.shape
    O, Cg, kh, kw = w.shape
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    y = np.zeros((B, O, Ho, Wo), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, :, ky * dil: ky * dil + stride * (Ho - 1) + 1: stride,
                       kx * dil: kx * dil + stride * (Wo - 1) + 1: stride]
            if groups == 1:
                y += np.einsum("bchw,oc->bohw", patch, w[:, :, ky, kx], optimize=True)
            elif groups == C and Cg == 1:
                y += patch * w[:, 0, ky, kx][None, :, None, None]
            else:
                raise NotImplementedError
    if b is not None:
        y += b[None, :, None, None]
    return y


def relative_position_index(ws: int) -> np.ndarray:
    """pgrm.py:133-143: idx(n, m) = (i_n - i_m + ws-1) * (2ws-1) + (j_n - j_m + ws-1), n = i*ws + j."""
    ii, jj = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    i = ii.reshape(-1)
    j = jj.reshape(-1)
    return (i[:, None] - i[None, :] + ws - 1) * (2 * ws - 1) + (j[:, None] - j[None, :] + ws - 1)


def shift_mask(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """pgrm.py:153-173: (nW, N, N) of {0, -100} from the 3x3 slice labelling of the rolled image."""
    def region(x, L):
        return (x >= L - ws).astype(np.int64) + (x >= L - shift).astype(np.int64)
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    label = 3 * region(hh, H) + region(ww, W)                         # (H, W) in rolled coordinates
    lw = label.reshape(H // ws, ws, W // ws, ws).transpose(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = lw[:, None, :] - lw[:, :, None]
    return np.where(diff != 0, -100.0, 0.0).astype(np.float32)


def window_token_order(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """Row p = w*N + n of the window-major list -> original token index t (pgrm.py:209-221, 43-52).

    Rolled coords (h', w') read original ((h'+shift) % H, (w'+shift) % W)  [torch.roll by -shift]."""
    nWw = W // ws
    p = np.arange(H * W)
    w_idx, n = p // (ws * ws), p % (ws * ws)
    hp = (w_idx // nWw) * ws + n // ws
    wp = (w_idx % nWw) * ws + n % ws
    return ((hp + shift) % H) * W + (wp + shift) % W


def effective_windows(cfg_windows: Sequence[int], block: int, H: int, W: int):
    """pgrm.py:147-151 (clamp) with pgrm.py:362 (block 0: shift 0, block 1: ws//2)."""
    ws_eff, sh_eff = [], []
    for ws in cfg_windows:
        sh = 0 if block % 2 == 0 else ws // 2
        if min(H, W) <= ws:
            ws, sh = min(H, W), 0
        ws_eff.append(ws)
        sh_eff.append(sh)
    return ws_eff, sh_eff


def window_attention_core(q: np.ndarray, kv: np.ndarray, tables: List[np.ndarray], windows: Sequence[int],
                          shifts: Sequence[int], H: int, W: int, heads_per_group: int) -> np.ndarray:
    """pgrm.py:197-268: per-group windowed MHA; returns the concatenated (B, L, C) tensor that enters SKConv.

    q: (B, L, C) projected queries, kv: (B, L, 2C) projected keys|values (channel halves, pgrm.py:194).
    `tables[g]` is relative_position_bias_table_g ((2ws-1)^2, heads_per_group) for the CONFIGURED window."""
    B, L, C = q.shape
    G = len(windows)
    cg = C // G
    d = cg // heads_per_group
    k_all, v_all = kv[..., :C], kv[..., C:]
    outs = []
    for g in range(G):
        ws, sh = windows[g], shifts[g]
        N = ws * ws
        nW = L // N
        order = window_token_order(H, W, ws, sh)
        sl = slice(g * cg, (g + 1) * cg)

        def part(x):  # (B, L, cg) -> (B, nW, heads, N, d)
            return x[:, order, sl].reshape(B, nW, N, heads_per_group, d).transpose(0, 1, 3, 2, 4)
        qg = part(q) * (d ** -0.5)                                    # pgrm.py:230-231
        kg, vg = part(k_all), part(v_all)
        s = qg @ kg.transpose(0, 1, 2, 4, 3)                          # (B, nW, h, N, N)  pgrm.py:232
        tbl = tables[g]
        side = int(round(math.sqrt(tbl.shape[0])))
        ws_tbl = (side + 1) // 2
        idx = relative_position_index(ws_tbl)
        bias = tbl[idx.reshape(-1)].reshape(ws_tbl * ws_tbl, ws_tbl * ws_tbl, -1).transpose(2, 0, 1)
        if bias.shape[1] != N:  # reference: .view(N, N, -1) fails for a clamped window; mirror that
            raise ValueError("relative position table does not match the effective window (pgrm.py:234-236)")
        s = s + bias[None, None].astype(s.dtype)                      # pgrm.py:234-238
        if sh > 0:
            s = s + shift_mask(H, W, ws, sh)[None, :, None].astype(s.dtype)   # pgrm.py:240-243
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s)
        p = e / e.sum(axis=-1, keepdims=True)                         # pgrm.py:244-246
        o = p @ vg                                                    # (B, nW, h, N, d)  pgrm.py:249
        o = o.transpose(0, 1, 3, 2, 4).reshape(B, nW * N, cg)         # window-major rows kept (quirk 1)
        outs.append(o.astype(q.dtype))
    return np.concatenate(outs, axis=-1)


def sk_gate(x: np.ndarray, P: Dict[str, np.ndarray], pre: str, G: int) -> np.ndarray:
    """SKConv.forward, pgrm.py:79-96, on token-major x (B, L, C); returns token-major (B, L, C)."""
    B, L, C = x.shape
    cg = C // G
    f = linear(x, P[pre + "proj.weight"], P[pre + "proj.bias"])                       # :82
    s = gelu(f).mean(axis=1)                                                          # :84-86 (GAP over L)
    z = gelu(linear(s, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))                   # :87-88
    a = linear(z, P[pre + "fc2.weight"], P[pre + "fc2.bias"]).reshape(B, G, cg)       # :89-90
    a = a - a.max(axis=1, keepdims=True)
    a = np.exp(a)
    a = a / a.sum(axis=1, keepdims=True)                                              # :91 softmax over groups
    v = (x.reshape(B, L, G, cg) * a[:, None]).sum(axis=2)                             # :92
    v = linear(v.astype(x.dtype), P[pre + "proj_head.weight"], P[pre + "proj_head.bias"])   # :93
    return (f + v).astype(x.dtype)                                                    # :95


def mlp(x: np.ndarray, P: Dict[str, np.ndarray], pre: str) -> np.ndarray:
    """Mlp.forward, pgrm.py:29-41, including the raw (no-transpose) views (quirk 2)."""
    B, L, _ = x.shape
    side = int(math.sqrt(L))
    h = gelu(linear(x, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))                   # :30-31
    hid = h.shape[-1]
    h = np.ascontiguousarray(h).reshape(B, hid, side, side)                           # :34 raw view
    h = gelu(conv2d(h, P[pre + "depthwise_conv.weight"], P[pre + "depthwise_conv.bias"], pad=1, groups=hid))
    h = conv2d(h, P[pre + "pointwise_conv.weight"], P[pre + "pointwise_conv.bias"])   # :37
    h = np.ascontiguousarray(h).reshape(B, L, hid)                                    # :38 raw view back
    return linear(h, P[pre + "fc2.weight"], P[pre + "fc2.bias"])                      # :39


def swin_block(x_q: np.ndarray, x_kv: np.ndarray, P: Dict[str, np.ndarray], pre: str, block: int,
               windows: Sequence[int], H: int, W: int, heads_per_group: int, return_parts: bool = False):
    """SwinTransformerBlock.forward, pgrm.py:315-331 (eval: DropPath is identity)."""
    G = len(windows)
    ws_eff, sh_eff = effective_windows(windows, block, H, W)
    qn = layer_norm(x_q, P[pre + "norm1_q.weight"], P[pre + "norm1_q.bias"])
    kvn = layer_norm(x_kv, P[pre + "norm1_kv.weight"], P[pre + "norm1_kv.bias"])
    q = linear(qn, P[pre + "attn.q.weight"], P[pre + "attn.q.bias"])                  # :188
    kv = linear(kvn, P[pre + "attn.kv.weight"], P[pre + "attn.kv.bias"])              # :194
    tables = [P[pre + f"attn.relative_position_bias_table_{g}"] for g in range(G)]
    a = window_attention_core(q, kv, tables, ws_eff, sh_eff, H, W, heads_per_group)
    y = x_kv + sk_gate(a, P, pre + "attn.sknet.", G)                                  # :329
    out = y + mlp(layer_norm(y, P[pre + "norm2.weight"], P[pre + "norm2.bias"]), P, pre + "mlp.")   # :330
    if return_parts:
        return out, {"q": q, "kv": kv, "attn_core": a, "y": y}
    return out


def patch_embed(x: np.ndarray, P: Dict[str, np.ndarray], patch: int) -> np.ndarray:
    """PatchEmbed.forward, pgrm.py:419-426: conv(k=s=patch) -> flatten(2).transpose(1,2) -> LayerNorm."""
    y = conv2d(x, P["patch_embed.proj.weight"], P["patch_embed.proj.bias"], stride=patch)
    B, C, H, W = y.shape
    t = y.reshape(B, C, H * W).transpose(0, 2, 1)
    return layer_norm(np.ascontiguousarray(t), P["patch_embed.norm.weight"], P["patch_embed.norm.bias"])


def pixel_shuffle(x: np.ndarray, r: int) -> np.ndarray:
    B, C, H, W = x.shape
    c = C // (r * r)
    return x.reshape(B, c, r, r, H, W).transpose(0, 1, 4, 2, 5, 3).reshape(B, c, H * r, W * r)


def pgrm_forward(P: Dict[str, np.ndarray], x_q: np.ndarray, x_kv: np.ndarray,
                 residual_list: Sequence[np.ndarray], *, windows: Sequence[int] = (2, 4, 8),
                 num_heads: int = 6, patch: int = 2, dtype=np.float32) -> np.ndarray:
    """PGRM.forward, pgrm.py:546-565 (eval mode, ape=False)."""
    P = {k: (v.astype(dtype) if np.issubdtype(v.dtype, np.floating) else v) for k, v in P.items()}
    x_q = x_q.astype(dtype)
    x_kv = x_kv.astype(dtype)
    if x_q.shape[1] == 2:                                                             # :547-548
        x_q = conv2d(x_q, P["prior_fusion.weight"], P["prior_fusion.bias"], pad=1)
    H, W = x_kv.shape[2] // patch, x_kv.shape[3] // patch
    tq = patch_embed(x_q, P, patch)                                                   # :549
    tkv = patch_embed(x_kv, P, patch)                                                 # :550
    G = len(windows)
    for blk in range(2):                                                              # :557-558, depth 2
        tkv = swin_block(tq, tkv, P, f"layers.0.blocks.{blk}.", blk, windows, H, W, num_heads // G)
    B, L, C = tkv.shape
    x = np.ascontiguousarray(tkv.transpose(0, 2, 1)).reshape(B, C, H, W)              # :559 (proper transpose)
    x = conv2d(x, P["conv_before_upsample.0.weight"], P["conv_before_upsample.0.bias"], pad=1)
    x = conv2d(x, P["conv_before_upsample.1.weight"], P["conv_before_upsample.1.bias"], pad=1)
    x = np.where(x >= 0, x, 0.01 * x).astype(dtype)                                   # LeakyReLU default slope
    x = pixel_shuffle(x, patch)                                                       # :561
    x = x * P["weight_list_0"]                                                        # :562
    for i in range(1, len(residual_list)):                                            # :563-564 (index 0 skipped)
        x = x + residual_list[i].astype(dtype) * P[f"weight_list_{i}"]
    return x.astype(dtype)
