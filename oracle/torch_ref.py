"""torch-CPU functional restatement of the reference PGRM / CMM forward (eval or train-BN).
TEST INFRASTRUCTURE ONLY -- and the CPU baseline `bench.py` times (`cpu_baseline.kind = "port"`).

Why a second oracle: the reference's own CPU path is multi-threaded PyTorch (oneDNN/MKL); the numpy
oracle (pgrm_oracle.py / cmm_oracle.py) is exact but single-threaded einsum.  This file restates the
same algorithm with torch ops so that the CPU baseline is timed on the libraries and thread count the
reference would use on the same host.  /root/reference does not exist on the GPU box, so the
reference itself cannot be timed there.  Pinned by tests/test_oracle_golden.py against the golden
fixtures (outputs of the unmodified reference).  Citations: /root/reference/model/pgrm.py, cmm.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


# ---- train-mode Dropout / DropPath masks: numpy restatement of dpmn_mask_hash (include/dpmn_b200.h) -----------------
# The reference draws its masks from torch's RNG stream (not reproducible outside torch); the CUDA path derives them
# from (seed, site, index).  This oracle applies the SAME masks at the reference's Dropout / DropPath sites
# (pgrm.py:32,40,248,329-330,554-555), so train-mode forward and backward can be checked exactly.
SITE_POS_Q, SITE_POS_KV, SITE_BLOCK = 1, 2, 16
SITE_ATTN, SITE_MLP1, SITE_MLP2, SITE_PATH1, SITE_PATH2 = 0, 1, 2, 3, 4


def mask_hash(seed: int, site: int, idx):
    import numpy as np
    M = np.uint64
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = M(seed) + M(0x9E3779B97F4A7C15) * (idx + M(1)) + M(0xD1B54A32D192ED03) * M(site + 1)
        z = (z ^ (z >> M(30))) * M(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> M(27))) * M(0x94D049BB133111EB)
        z = z ^ (z >> M(31))
    return (z >> M(32)).astype(np.uint32)


def drop_scale(p: float, seed: int, site: int, idx) -> torch.Tensor:
    """multiplier per element: 1/(1-p) if kept, 0 if dropped (all ones for p == 0), fp32 like the kernels."""
    import numpy as np
    idx = np.asarray(idx, dtype=np.uint64)
    if p <= 0:
        return torch.ones(idx.shape, dtype=torch.float32)
    u = (mask_hash(seed, site, idx) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    keep = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(u >= np.float32(p), keep, np.float32(0.0)).astype(np.float32))


def _flat_idx(shape):
    import numpy as np
    return np.arange(int(np.prod(shape)), dtype=np.uint64).reshape(shape)


# ---- emulation of the 16-bit tensor-core modes' operand rounding (tests only) -----------------------------------------
# In the fp16 / bf16 modes the CUDA path rounds both OPERANDS of every tensor-core contraction to 16 bits and accumulates in
# fp32; everything else (LayerNorm, softmax, BatchNorm statistics, activations, the residual stream) stays fp32.  With
# OPERAND_ROUND set, the contractions below round their operands the same way (straight-through: the rounding is invisible
# to autograd, exactly like the CUDA backward, which differentiates the fp32 formulas), so this oracle reproduces the
# 16-bit forward -- and therefore its ReLU / LeakyReLU / Dropout masks -- to fp32 accumulation order.  Gradient tests of
# the 16-bit modes compare against THIS evaluation: what is left is the rounding of the backward's own operands.
OPERAND_ROUND = None          # None | torch.float16 | torch.bfloat16


class operand_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_ROUND
        self.prev, OPERAND_ROUND = OPERAND_ROUND, self.dtype

    def __exit__(self, *exc):
        global OPERAND_ROUND
        OPERAND_ROUND = self.prev


def _r(x, on=True):
    if OPERAND_ROUND is None or not on:
        return x
    return x + (x.detach().to(OPERAND_ROUND).to(x.dtype) - x.detach())


def _linear_tc(x, w, b):
    return F.linear(_r(x), _r(w), b)


def _order(H, W, ws, shift):
    """window-major row -> original token (pgrm.py:209-221,43-52)."""
    p = torch.arange(H * W)
    w_idx, n = p // (ws * ws), p % (ws * ws)
    nWw = W // ws
    hp = (w_idx // nWw) * ws + n // ws
    wp = (w_idx % nWw) * ws + n % ws
    return ((hp + shift) % H) * W + (wp + shift) % W


def _rel_index(ws):
    i, j = torch.arange(ws * ws) // ws, torch.arange(ws * ws) % ws
    return (i[:, None] - i[None, :] + ws - 1) * (2 * ws - 1) + (j[:, None] - j[None, :] + ws - 1)


def _shift_mask(H, W, ws, shift):
    def region(x, n):
        return (x >= n - ws).long() + (x >= n - shift).long()
    lab = 3 * region(torch.arange(H), H)[:, None] + region(torch.arange(W), W)[None, :]
    lab = lab.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    return torch.where(lab[:, None, :] != lab[:, :, None], -100.0, 0.0)


def window_attention_core(q, kv, tables, windows, shifts, H, W, hpg, drop=None, site=0):
    """pgrm.py:197-268 -> (B, L, C) in window-major row order per group (quirk 1).  drop = (p, seed): attn_drop masks."""
    B, L, C = q.shape
    G = len(windows)
    cg = C // G
    d = cg // hpg
    k_all, v_all = kv[..., :C], kv[..., C:]
    outs = []
    for g, (ws, sh) in enumerate(zip(windows, shifts)):
        N = ws * ws
        nW = L // N
        order = _order(H, W, ws, sh)
        sl = slice(g * cg, (g + 1) * cg)

        def part(x):
            return x[:, order][..., sl].reshape(B, nW, N, hpg, d).permute(0, 1, 3, 2, 4)
        # OPERAND_ROUND (16-bit training emulation): the tcgen05 kernel reads q, k, v as 16-bit values, rounds the
        # UNNORMALISED probabilities exp(s - max) to 16 bits for the P V product, divides by the fp32 row sum afterwards and
        # stores a 16-bit result (csrc/attn2_tc.cu)
        qg, kg, vg = _r(part(q)), _r(part(k_all)), _r(part(v_all))
        s = (qg * d ** -0.5) @ kg.transpose(-1, -2)
        bias = tables[g][_rel_index(ws).reshape(-1)].view(N, N, hpg).permute(2, 0, 1)
        s = s + bias[None, None]
        if sh > 0:
            s = s + _shift_mask(H, W, ws, sh)[None, :, None]
        keep = None
        if drop is not None and drop[0] > 0:                      # attn_drop, pgrm.py:248
            import numpy as np
            b_i, w_i, h_i, n_i, m_i = np.meshgrid(np.arange(B), np.arange(nW), np.arange(hpg), np.arange(N), np.arange(N),
                                                  indexing="ij")
            idx = ((((b_i * G + g) * hpg + h_i) * L + w_i * N + n_i) * N + m_i).astype(np.uint64)
            keep = drop_scale(drop[0], drop[1], site, idx)
        if OPERAND_ROUND is None:
            prob = torch.softmax(s, dim=-1)
            if keep is not None:
                prob = prob * keep
            o = prob @ vg
        else:
            e = torch.exp(s - s.amax(dim=-1, keepdim=True).detach())
            den = e.sum(dim=-1, keepdim=True)
            e = _r(e)
            if keep is not None:
                e = e * keep
            o = _r((e @ vg) / den)
        outs.append(o.permute(0, 1, 3, 2, 4).reshape(B, L, cg))
    return torch.cat(outs, dim=-1)


def _sk(x, P, pre, G):
    """SKConv.forward, pgrm.py:79-96, token-major."""
    B, L, C = x.shape
    cg = C // G
    f = _linear_tc(x, P[pre + "proj.weight"], P[pre + "proj.bias"])
    s = F.gelu(f).mean(dim=1)
    z = F.gelu(F.linear(s, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))
    a = torch.softmax(F.linear(z, P[pre + "fc2.weight"], P[pre + "fc2.bias"]).view(B, G, cg), dim=1)
    v = (x.view(B, L, G, cg) * a[:, None]).sum(dim=2)
    return f + _linear_tc(v, P[pre + "proj_head.weight"], P[pre + "proj_head.bias"])


def _mlp(x, P, pre, drop=None, site=0):
    """Mlp.forward, pgrm.py:29-41, raw views kept (quirk 2).  drop = (p, seed): the two Dropout sites."""
    B, L, _ = x.shape
    side = int(math.sqrt(L))
    h = F.gelu(_linear_tc(x, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))
    if drop is not None and drop[0] > 0:
        h = h * drop_scale(drop[0], drop[1], site + SITE_MLP1, _flat_idx(h.shape))
    hid = h.shape[-1]
    h = h.reshape(B, hid, side, side)
    h = F.gelu(F.conv2d(h, P[pre + "depthwise_conv.weight"], P[pre + "depthwise_conv.bias"], padding=1, groups=hid))
    h = F.conv2d(_r(h), _r(P[pre + "pointwise_conv.weight"]), P[pre + "pointwise_conv.bias"])
    y = _linear_tc(h.reshape(B, L, hid), P[pre + "fc2.weight"], P[pre + "fc2.bias"])
    if drop is not None and drop[0] > 0:
        y = y * drop_scale(drop[0], drop[1], site + SITE_MLP2, _flat_idx(y.shape))
    return y


def _block(tq, tkv, P, pre, blk, windows, H, W, hpg, drop=None):
    """SwinTransformerBlock.forward, pgrm.py:315-331.  drop = dict(seed, drop_rate, attn_drop_rate, drop_path=[..])
    applies the train-mode masks; None = eval."""
    C = tq.shape[-1]
    G = len(windows)
    mn = min(H, W)
    shifts = [0 if (blk % 2 == 0 or mn <= ws) else ws // 2 for ws in windows]
    wins = [min(ws, mn) for ws in windows]
    qn = F.layer_norm(tq, (C,), P[pre + "norm1_q.weight"], P[pre + "norm1_q.bias"])
    kvn = F.layer_norm(tkv, (C,), P[pre + "norm1_kv.weight"], P[pre + "norm1_kv.bias"])
    q = _linear_tc(qn, P[pre + "attn.q.weight"], P[pre + "attn.q.bias"])
    kv = _linear_tc(kvn, P[pre + "attn.kv.weight"], P[pre + "attn.kv.bias"])
    tables = [P[pre + f"attn.relative_position_bias_table_{g}"] for g in range(G)]
    if drop is None:
        a = window_attention_core(q, kv, tables, wins, shifts, H, W, hpg)
        y = tkv + _sk(a, P, pre + "attn.sknet.", G)
        return y + _mlp(F.layer_norm(y, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"]), P, pre + "mlp.")
    import numpy as np
    site, seed, B = SITE_BLOCK * (blk + 1), drop["seed"], tq.shape[0]
    a = window_attention_core(q, kv, tables, wins, shifts, H, W, hpg, drop=(drop["attn_drop_rate"], seed), site=site + SITE_ATTN)
    path = drop["drop_path"][blk]
    dp1 = drop_scale(path, seed, site + SITE_PATH1, np.arange(B)).view(B, 1, 1)
    dp2 = drop_scale(path, seed, site + SITE_PATH2, np.arange(B)).view(B, 1, 1)
    y = tkv + dp1 * _sk(a, P, pre + "attn.sknet.", G)                                     # pgrm.py:329
    m = _mlp(F.layer_norm(y, (C,), P[pre + "norm2.weight"], P[pre + "norm2.bias"]), P, pre + "mlp.",
             drop=(drop["drop_rate"], seed), site=site)
    return y + dp2 * m                                                                   # pgrm.py:330


def pgrm_forward(P: Dict[str, torch.Tensor], x_q, x_kv, residual_list: Sequence[torch.Tensor], *,
                 windows=(2, 4, 8), num_heads=6, patch=2, drop=None, parts=None):
    """PGRM.forward, pgrm.py:546-565.  drop=None: eval mode; drop=dict(seed, drop_rate, attn_drop_rate, drop_path):
    train mode with the CUDA path's masks."""
    if x_q.shape[1] == 2:
        x_q = F.conv2d(x_q, P["prior_fusion.weight"], P["prior_fusion.bias"], padding=1)
    C = P["patch_embed.proj.weight"].shape[0]

    def embed(x):
        y = F.conv2d(x, P["patch_embed.proj.weight"], P["patch_embed.proj.bias"], stride=patch)
        return F.layer_norm(y.flatten(2).transpose(1, 2), (C,), P["patch_embed.norm.weight"], P["patch_embed.norm.bias"])
    H, W = x_kv.shape[2] // patch, x_kv.shape[3] // patch
    tq, tkv = embed(x_q), embed(x_kv)
    if drop is not None and drop["drop_rate"] > 0:                                        # pos_drop, pgrm.py:554-555
        tq = tq * drop_scale(drop["drop_rate"], drop["seed"], SITE_POS_Q, _flat_idx(tq.shape))
        tkv = tkv * drop_scale(drop["drop_rate"], drop["seed"], SITE_POS_KV, _flat_idx(tkv.shape))
    G = len(windows)
    for blk in range(2):
        tkv = _block(tq, tkv, P, f"layers.0.blocks.{blk}.", blk, windows, H, W, num_heads // G, drop=drop)
    B = tkv.shape[0]
    x = tkv.transpose(1, 2).reshape(B, C, H, W)
    x = F.conv2d(x, P["conv_before_upsample.0.weight"], P["conv_before_upsample.0.bias"], padding=1)
    x = F.conv2d(x, P["conv_before_upsample.1.weight"], P["conv_before_upsample.1.bias"], padding=1)
    if parts is not None:
        parts["head_pre"] = x                 # LeakyReLU input (B, hs * patch^2, H, W): tests mask out elements next to 0
    x = F.pixel_shuffle(F.leaky_relu(x, 0.01), patch)
    x = x * P["weight_list_0"]
    for i in range(1, len(residual_list)):   # residual_list[0] skipped (pgrm.py:563)
        x = x + residual_list[i] * P[f"weight_list_{i}"]
    return x


def _bn(x, P, pre, training):
    return F.batch_norm(x, None if training else P[pre + ".running_mean"], None if training else P[pre + ".running_var"],
                        P[pre + ".weight"], P[pre + ".bias"], training=training, eps=1e-5)


def _tc_conv_ok(ho, wo):
    """conv_tc_im2col_ok (csrc/cmm_im2col.cu): which convs of the fp32-structured CMM path run on the tensor cores."""
    return (ho * wo) % 8 == 0 and ho * wo >= 16


def _conv(x, w, b, **kw):
    y = F.conv2d(x, w, b, **kw)
    if OPERAND_ROUND is None or not _tc_conv_ok(y.shape[2], y.shape[3]):
        return y
    return F.conv2d(_r(x), _r(w), b, **kw)


def _convT(x, w, b, **kw):
    y = F.conv_transpose2d(x, w, b, **kw)
    if OPERAND_ROUND is None or not _tc_conv_ok(y.shape[2], y.shape[3]):
        return y
    return F.conv_transpose2d(_r(x), _r(w), b, **kw)


def cmm_forward(P: Dict[str, torch.Tensor], x1, x2, training: bool = False, return_parts: bool = False):
    """ComplementationModulationModule.forward, cmm.py:120-161.  With return_parts also the per-layer tensors
    (post-BatchNorm, pre-activation): o{l}_{br}, mid{l}_{br}, z6_{br}, zgate, d6, dmid{l}, d{l}."""
    parts = {}
    skips, bott = [], []
    for br, x in ((1, x1), (2, x2)):
        o = [_conv(x, P[f"en_1_{br}.weight"], P[f"en_1_{br}.bias"], padding=1)]
        parts[f"o1_{br}"] = o[0]
        for lvl in (2, 3, 4, 5):
            pre = f"en_{lvl}_{br}.encode."
            t = _conv(F.leaky_relu(o[-1], 0.2), P[pre + "1.weight"], P[pre + "1.bias"], stride=2, padding=3, dilation=2)
            t = _bn(t, P, pre + "2", training)
            parts[f"mid{lvl}_{br}"] = t
            t = _conv(F.leaky_relu(t, 0.2), P[pre + "4.weight"], P[pre + "4.bias"], padding=1)
            o.append(_bn(t, P, pre + "5", training))
            parts[f"o{lvl}_{br}"] = o[-1]
        bott.append(_conv(F.leaky_relu(o[-1], 0.2), P[f"en_6_{br}.1.weight"], P[f"en_6_{br}.1.bias"], stride=2, padding=1))
        parts[f"z6_{br}"] = bott[-1]
        skips.append(o)
    z = torch.cat(bott, dim=1)
    g = z.mean(dim=(2, 3))
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, P["fc_1.weight"], P["fc_1.bias"])), P["fc_2.weight"], P["fc_2.bias"]))
    z = z * g[:, :, None, None] + z
    parts["zgate"] = z
    d = _bn(_convT(F.relu(z), P["de_6.1.weight"], P["de_6.1.bias"], stride=2, padding=1), P, "de_6.2", training)
    parts["d6"] = d
    for lvl in (5, 4, 3, 2):
        pre = f"de_{lvl}.decode."
        cat = torch.cat([d, skips[0][lvl - 1], skips[1][lvl - 1]], dim=1)
        t = _bn(_convT(F.relu(cat), P[pre + "1.weight"], P[pre + "1.bias"], stride=1, padding=1), P, pre + "2", training)
        parts[f"dmid{lvl}"] = t
        d = _bn(_convT(F.relu(t), P[pre + "4.weight"], P[pre + "4.bias"], stride=2, padding=1), P, pre + "5", training)
        parts[f"d{lvl}"] = d
    cat = torch.cat([d, skips[0][0], skips[1][0]], dim=1)
    y = _convT(F.relu(cat), P["de_1.1.weight"], P["de_1.1.bias"], stride=1, padding=1)
    return (y, parts) if return_parts else y


def hot_path_forward(pgrm_params: List[Dict[str, torch.Tensor]], cmm_params, psn_out, priors_b1, priors_b2,
                     windows=(2, 4, 8), num_heads=6, alpha=None):
    """The call-site data flow of interfaces/super_resolution.py:174-265 (inference): two cascades of
    three PGRMs on the PSN output, then the CMM.  priors_b1[k] (B,2,H,W), priors_b2[k] (B,3,H,W).
    alpha: the eval / test blend with the PSN image, super_resolution.py:449,705 (None = plain CMM output)."""
    outs = []
    for branch, priors in ((0, priors_b1), (1, priors_b2)):
        cascade = psn_out[:, :3]
        done = []
        for k in range(3):
            y = pgrm_forward(pgrm_params[3 * branch + k], priors[k], cascade, done[:k], windows=windows, num_heads=num_heads)
            done.append(y)
            cascade = y
        outs.append(done[-1])
    y = cmm_forward(cmm_params, outs[0], outs[1], training=False)
    return y if alpha is None else alpha * y + (1 - alpha) * psn_out[:, :3]


# ---- the callers' steps on either side of the hot path ------------------------------------------------
def gradient_map(x: torch.Tensor) -> torch.Tensor:
    """GradientPriorLoss.gradient_map, loss/image_loss.py:33-43."""
    w = x.shape[-1]
    h = x.shape[-2]
    r = F.pad(x, (0, 1, 0, 0))[:, :, :, 1:]
    l = F.pad(x, (1, 0, 0, 0))[:, :, :, :w]
    t = F.pad(x, (0, 0, 1, 0))[:, :, :h, :]
    b = F.pad(x, (0, 0, 0, 1))[:, :, 1:, :]
    return torch.sqrt(((r - l) * 0.5) ** 2 + ((t - b) * 0.5) ** 2 + 1e-6)


def image_loss(out: torch.Tensor, target: torch.Tensor, weight=(1.0, 1.0)) -> torch.Tensor:
    """ImageLoss(gradient=True, loss_weight=weight).forward, loss/image_loss.py:15-24."""
    return weight[0] * F.mse_loss(out, target) + weight[1] * F.l1_loss(gradient_map(out[:, :3]), gradient_map(target[:, :3]))


def to_mask(img):
    """toMask, utils/util.py:27-35, for one (3, H, W) image in [0, 1]: numpy restatement (uint8 truncation, Pillow's
    fixed-point ITU-R 601 luma, threshold at the mean luma, inverted, / 255, repeated on 3 channels)."""
    import numpy as np
    u8 = ((np.asarray(img, dtype=np.float32) * np.float32(255)).astype(np.int64) & 255)    # torch: float -> int64 -> uint8 (wraps)
    luma = (u8[0] * 19595 + u8[1] * 38470 + u8[2] * 7471 + 0x8000) >> 16
    m = np.where(luma * luma.size > luma.sum(), 0.0, 1.0).astype(np.float32)
    return np.repeat(m[None], 3, axis=0)


def distill_forward(P: Dict[str, torch.Tensor], x_deep: torch.Tensor, x_shallow: torch.Tensor, training: bool = True,
                    eps: float = 1e-5):
    """DistillModule.forward, model/distill_module.py:18-31 -> (loss, feature_cat).  P: the module's state_dict
    (conv_cat_feature.*, bn_1.*, conv_feature.*, bn_2.*).  training=True: batch statistics (biased variance), the
    running buffers are not touched here (see distill_running_update)."""
    def bn(x, stem):
        if training:
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        else:
            mean = P[stem + ".running_mean"].view(1, -1, 1, 1)
            var = P[stem + ".running_var"].view(1, -1, 1, 1)
        return (x - mean) / torch.sqrt(var + eps) * P[stem + ".weight"].view(1, -1, 1, 1) + P[stem + ".bias"].view(1, -1, 1, 1)
    u = F.conv2d(torch.cat([x_deep, x_shallow], dim=1), P["conv_cat_feature.weight"], P["conv_cat_feature.bias"], padding=1)
    a = torch.relu(bn(u, "bn_1"))                                                                    # :19-22
    v = F.conv2d(x_shallow, P["conv_feature.weight"], P["conv_feature.bias"], padding=1)
    s = torch.relu(bn(v, "bn_2"))                                                                    # :24-26
    return (a - s).abs().mean(), a                                                                   # :28,31


def distill_running_update(P: Dict[str, torch.Tensor], x_deep, x_shallow, momentum: float = 0.1):
    """What one train-mode forward does to the BatchNorm buffers (nn.BatchNorm2d: unbiased variance, momentum 0.1)."""
    out = {}
    u = F.conv2d(torch.cat([x_deep, x_shallow], dim=1), P["conv_cat_feature.weight"], P["conv_cat_feature.bias"], padding=1)
    v = F.conv2d(x_shallow, P["conv_feature.weight"], P["conv_feature.bias"], padding=1)
    for stem, x in (("bn_1", u), ("bn_2", v)):
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False) * (n / (n - 1))
        out[stem + ".running_mean"] = (1 - momentum) * P[stem + ".running_mean"] + momentum * mean
        out[stem + ".running_var"] = (1 - momentum) * P[stem + ".running_var"] + momentum * var
    return out


# ---- recogniser input resizes (interfaces/base.py:419-425,473-478) --------------------------------
def _cubic_weights(t, A=-0.75):
    """Keys cubic convolution coefficients as ATen's upsample_bicubic2d uses them (A = -0.75)."""
    import numpy as np
    def inner(x):
        return ((A + 2) * x - (A + 3)) * x * x + 1
    def outer(x):
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return np.stack([outer(t + 1), inner(t), inner(1 - t), outer(2 - t)], axis=-1)


def parse_crnn_data(imgs, out_hw=(32, 100)):
    """TextBase.parse_crnn_data, interfaces/base.py:419-425: F.interpolate(imgs, (32, 100), mode='bicubic')
    (align_corners=False: src = (dst + 0.5) * in/out - 0.5, taps clamped to the image) then 0.299 R + 0.587 G + 0.114 B.
    numpy restatement; imgs (B, 3, H, W) -> (B, 1, 32, 100)."""
    import numpy as np
    x = np.asarray(imgs, dtype=np.float32)
    B, Ch, H, W = x.shape
    oh, ow = out_hw
    def axis(n_in, n_out):
        src = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) * np.float32(n_in / n_out) - np.float32(0.5)
        i0 = np.floor(src)
        w = _cubic_weights((src - i0).astype(np.float32)).astype(np.float32)            # (n_out, 4)
        idx = np.clip(i0.astype(np.int64)[:, None] + np.arange(-1, 3)[None, :], 0, n_in - 1)
        return idx, w
    iy, wy = axis(H, oh)
    ix, wx = axis(W, ow)
    rows = (x[:, :, :, ix] * wx[None, None, None]).sum(-1, dtype=np.float32)            # (B, C, H, ow)
    out = (rows[:, :, iy, :] * wy[None, None, :, :, None]).sum(3, dtype=np.float32)     # (B, C, oh, ow)
    return (np.float32(0.299) * out[:, 0:1] + np.float32(0.587) * out[:, 1:2] + np.float32(0.114) * out[:, 2:3]).astype(np.float32)


def parse_visionlan_data(img, out_hw=(64, 256)):
    """TextBase.parse_visionlan_data, interfaces/base.py:473-478, for one (3, H, W) image: ToPILImage (x * 255 truncated
    to uint8) -> cv2.resize(..., (256, 64)) (INTER_LINEAR on uint8: OpenCV's 11-bit fixed-point bilinear) -> ToTensor
    (/ 255) -> (1, 3, 64, 256).  Integer restatement, bit-exact against cv2 (opencv-python 4.13 in the build container)."""
    import numpy as np
    u8 = ((np.asarray(img, dtype=np.float32) * np.float32(255)).astype(np.int64) & 255)          # (3, H, W)
    _, sh, sw = u8.shape
    dh, dw = out_hw

    def coefs(dn, sn, clamp):
        scale = 1.0 / (dn / float(sn))
        f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp:                              # columns: the weight collapses onto the border pixel
            lo, hi = s < 0, s >= sn - 1
            f[lo | hi] = 0
            s[lo] = 0
            s[hi] = sn - 1
        return s, np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64), np.rint(f * np.float32(2048)).astype(np.int64)
    sx, a0, a1 = coefs(dw, sw, True)
    sy, b0, b1 = coefs(dh, sh, False)           # rows: indices are clipped instead, the weights stay
    rows = u8[:, :, sx] * a0 + u8[:, :, np.minimum(sx + 1, sw - 1)] * a1                         # (3, sh, dw), scale 2^11
    r0, r1 = rows[:, np.clip(sy, 0, sh - 1)], rows[:, np.clip(sy + 1, 0, sh - 1)]
    out = (((b0[None, :, None] * (r0 >> 4)) >> 16) + ((b1[None, :, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return (out.astype(np.float32) / np.float32(255))[None]
