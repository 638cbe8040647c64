"""Deterministic synthetic parameters keyed by state_dict name.  TEST INFRASTRUCTURE ONLY.

The reference's default init leaves every Linear/LayerNorm bias at 0, LayerNorm weights at 1 and
`weight_list_*` at 1 (pgrm.py:496-497,524-533), which hides whole classes of bugs (a dropped bias, a
swapped LayerNorm).  Parity fixtures therefore use these non-trivial values instead; they are a pure
function of (seed, name, shape), so fixtures only need to store inputs and outputs, not weights.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np


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
    if leaf == "bias":
        return 0.05 * n
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
