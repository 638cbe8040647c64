"""Seeded synthetic inputs and parameters for measurement and tests (numpy only; no algorithm of the path lives here).

Inputs have the shapes / value ranges of the reference call sites:
  branch-1 prior  x_q = round(U[0,255])  (B,2,32,128)  un-normalised uint8 glyph maps, super_resolution.py:188-193
  branch-2 prior  x_q = Bernoulli(.5) in {0,1} repeated on 3 channels  (toMask, utils/util.py:27-35)
  image stream    x_kv = U[0,1)  (B,3,32,128)   (PSN output slice cascade[:, :3], super_resolution.py:196)

Parameters: the reference's default init leaves every Linear/LayerNorm bias at 0, LayerNorm weights at 1 and
`weight_list_*` at 1 (pgrm.py:496-497,524-533), which hides whole classes of bugs (a dropped bias, a swapped LayerNorm) and
makes a benchmark forward unrepresentative.  `synth_params` gives non-trivial values that are a pure function of
(seed, name, shape), so the golden fixtures only need to store seeds and outputs, and bench.py / the tests / the fixture
generators (oracle/make_golden*.py) all see the same tensors.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np


def prior_branch1(seed: int, B: int, H: int = 32, W: int = 128) -> np.ndarray:
    r = np.random.default_rng([seed, 1])
    return np.round(r.uniform(0.0, 255.0, size=(B, 2, H, W))).astype(np.float32)


def prior_branch2(seed: int, B: int, H: int = 32, W: int = 128) -> np.ndarray:
    r = np.random.default_rng([seed, 2])
    m = (r.uniform(size=(B, 1, H, W)) < 0.5).astype(np.float32)
    return np.ascontiguousarray(np.repeat(m, 3, axis=1))


def image_stream(seed: int, B: int, H: int = 32, W: int = 128, tag: int = 3) -> np.ndarray:
    r = np.random.default_rng([seed, tag])
    return r.uniform(0.0, 1.0, size=(B, 3, H, W)).astype(np.float32)


def residuals(seed: int, B: int, n: int, H: int = 32, W: int = 128):
    return [image_stream(seed, B, H, W, tag=10 + i) for i in range(n)]


# ---- parameters ---------------------------------------------------------------------------------------------------



def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synth_value(seed: int, name: str, shape: Tuple[int, ...]) -> np.ndarray:
    r = _rng(seed, name)
    leaf = name.split(".")[-1]
    if leaf == "num_batches_tracked":
        return np.zeros((), dtype=np.int64)
    n = r.standard_normal(shape).astype(np.float32)
    if leaf == "running_mean":
        return 0.1 * n
    if leaf == "running_var":
        return (0.5 + np.abs(n)).astype(np.float32)
    if "relative_position_bias_table" in name:
        return 0.5 * n
    if name.startswith("weight_list_"):
        return (1.0 + 0.2 * n).astype(np.float32)
    if leaf == "bias" or leaf.startswith("bias_") or leaf.endswith("_bias"):     # incl. GRU bias_ih_l0, MHA in_proj_bias
        return 0.05 * n
    if shape == (1,):                # PReLU slope (torch default 0.25)
        return (0.25 + 0.05 * n).astype(np.float32)
    if len(shape) == 1:              # LayerNorm / BatchNorm scale
        return (1.0 + 0.1 * n).astype(np.float32)
    if "depthwise_conv" in name:
        return (n / 3.0).astype(np.float32)
    # conv-transpose weights are (Cin, Cout, k, k): fan-in is Cin*k*k/stride^2-ish; a plain
    # 1/sqrt(prod(shape[1:])) keeps activations O(1) for every layer kind, which is all that matters.
    fan = int(np.prod(shape[1:])) if len(shape) > 1 else 1
    if ".decode." in name or name.startswith("de_"):
        fan = int(shape[0] * np.prod(shape[2:]))
    return (n / np.sqrt(max(fan, 1))).astype(np.float32)


def synth_params(schema: Iterable[Tuple[str, Tuple[int, ...], str]], seed: int,
                 skip_computed_buffers: bool = True) -> Dict[str, np.ndarray]:
    """Values for every learnable / running-stat entry of `schema`; index & mask buffers are derived
    from the configuration (closed form), never synthesized."""
    out = {}
    for name, shape, kind in schema:
        if skip_computed_buffers and ("relative_position_index" in name or "attn_mask" in name):
            continue
        out[name] = synth_value(seed, name, tuple(shape))
    return out
