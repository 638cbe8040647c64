"""Seeded synthetic inputs with the shapes / value ranges of the reference call sites.

TEST INFRASTRUCTURE (shared by oracle/make_golden.py, tests/ and bench.py so that the golden fixtures
only need to store seeds and outputs).

  branch-1 prior  x_q = round(U[0,255])  (B,2,32,128)  un-normalised uint8 glyph maps, super_resolution.py:188-193
  branch-2 prior  x_q = Bernoulli(.5) in {0,1} repeated on 3 channels  (toMask, utils/util.py:27-35)
  image stream    x_kv = U[0,1)  (B,3,32,128)   (PSN output slice cascade[:, :3], super_resolution.py:196)
"""
from __future__ import annotations

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
