"""numpy restatement of ComplementationModulationModule.forward.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/model/cmm.py:38-161.  Parameters are a flat dict keyed like the reference
state_dict.  Parity: pinned against tests/golden/cmm_*.npz (minted by oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .pgrm_oracle import conv2d


def conv_transpose2d(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, stride: int, pad: int) -> np.ndarray:
    """torch ConvTranspose2d semantics: w is (Cin, Cout, kh, kw); out = (H-1)*s - 2p + k."""
    B, Ci, H, W = x.shape
    _, Co, kh, kw = w.shape
    Hf, Wf = (H - 1) * stride + kh, (W - 1) * stride + kw
    y = np.zeros((B, Co, Hf, Wf), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            y[:, :, ky: ky + stride * (H - 1) + 1: stride, kx: kx + stride * (W - 1) + 1: stride] += \
                np.einsum("bchw,co->bohw", x, w[:, :, ky, kx], optimize=True)
    y = y[:, :, pad: Hf - pad, pad: Wf - pad]
    if b is not None:
        y = y + b[None, :, None, None]
    return np.ascontiguousarray(y)


def batch_norm(x: np.ndarray, P: Dict[str, np.ndarray], pre: str, training: bool, eps: float = 1e-5) -> np.ndarray:
    """nn.BatchNorm2d (cmm.py:12): running stats in eval, biased batch stats in train."""
    if training:
        mu = x.mean(axis=(0, 2, 3))
        var = x.var(axis=(0, 2, 3))
    else:
        mu, var = P[pre + ".running_mean"], P[pre + ".running_var"]
    s = P[pre + ".weight"] / np.sqrt(var + eps)
    return ((x - mu[None, :, None, None]) * s[None, :, None, None] + P[pre + ".bias"][None, :, None, None]).astype(x.dtype)


def _lrelu(x, slope=0.2):   # cmm.py:26
    return np.where(x >= 0, x, slope * x).astype(x.dtype)


def _relu(x):
    return np.maximum(x, 0).astype(x.dtype)


def _encode_block(x, P, pre, training):
    """EncodeBlock, cmm.py:38-55: act, conv4x4 s2 d2 p3, BN, act, conv3x3 p1, BN."""
    x = conv2d(_lrelu(x), P[pre + "1.weight"], P[pre + "1.bias"], stride=2, pad=3, dil=2)
    x = batch_norm(x, P, pre + "2", training)
    x = conv2d(_lrelu(x), P[pre + "4.weight"], P[pre + "4.bias"], stride=1, pad=1)
    return batch_norm(x, P, pre + "5", training)


def _decode_block(x, P, pre, training):
    """DecodeBlock, cmm.py:58-77: act, convT3x3 p1, BN, act, convT4x4 s2 p1, BN."""
    x = conv_transpose2d(_relu(x), P[pre + "1.weight"], P[pre + "1.bias"], stride=1, pad=1)
    x = batch_norm(x, P, pre + "2", training)
    x = conv_transpose2d(_relu(x), P[pre + "4.weight"], P[pre + "4.bias"], stride=2, pad=1)
    return batch_norm(x, P, pre + "5", training)


def cmm_forward(P: Dict[str, np.ndarray], x1: np.ndarray, x2: np.ndarray, training: bool = False,
                dtype=np.float32) -> np.ndarray:
    """ComplementationModulationModule.forward, cmm.py:120-161."""
    P = {k: (v.astype(dtype) if np.issubdtype(v.dtype, np.floating) else v) for k, v in P.items()}
    skips = []
    bott = []
    for br, x in ((1, x1.astype(dtype)), (2, x2.astype(dtype))):
        o1 = conv2d(x, P[f"en_1_{br}.weight"], P[f"en_1_{br}.bias"], pad=1)                  # :121/:128
        o2 = _encode_block(o1, P, f"en_2_{br}.encode.", training)
        o3 = _encode_block(o2, P, f"en_3_{br}.encode.", training)
        o4 = _encode_block(o3, P, f"en_4_{br}.encode.", training)
        o5 = _encode_block(o4, P, f"en_5_{br}.encode.", training)
        o6 = conv2d(_lrelu(o5), P[f"en_6_{br}.1.weight"], P[f"en_6_{br}.1.bias"], stride=2, pad=1)  # :91-93
        skips.append((o1, o2, o3, o4, o5))
        bott.append(o6)
    z = np.concatenate(bott, axis=1)                                                         # :135
    g = z.mean(axis=(2, 3))                                                                  # :137-139
    g = np.maximum(g @ P["fc_1.weight"].T + P["fc_1.bias"], 0)                               # :140-141
    g = 1.0 / (1.0 + np.exp(-(g @ P["fc_2.weight"].T + P["fc_2.bias"])))                     # :142-143
    z = (z * g[:, :, None, None] + z).astype(dtype)                                          # :146-147
    d = conv_transpose2d(_relu(z), P["de_6.1.weight"], P["de_6.1.bias"], stride=2, pad=1)    # :108-111
    d = batch_norm(d, P, "de_6.2", training)
    for lvl in (5, 4, 3, 2):                                                                 # :150-157
        cat = np.concatenate([d, skips[0][lvl - 1], skips[1][lvl - 1]], axis=1)
        d = _decode_block(cat, P, f"de_{lvl}.decode.", training)
    cat = np.concatenate([d, skips[0][0], skips[1][0]], axis=1)                              # :158
    return conv_transpose2d(_relu(cat), P["de_1.1.weight"], P["de_1.1.bias"], stride=1, pad=1).astype(dtype)
