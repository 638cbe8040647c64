"""Golden fixture for the frozen TATT PSN from the UNMODIFIED reference class
`model/tatt.py::TSRN_TL_TRANS` (with its `transformer_v2.InfoTransformer`), eval mode, the constructor arguments of
`interfaces/base.py:145-148`.  The 7.6 M parameters are not stored: every state_dict entry is
`dpmn_b200.synth.synth_value(seed, name, shape)`, loaded into the reference here and into `dpmn_b200.psn.TATT` by the
test; the fixture holds the reference's key -> shape dump, the seeded inputs' seed and the reference outputs.
Run in the build container only:   python -m oracle.make_golden_tatt"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "tatt.npz")
SEED = 77


def tatt_inputs(seed: int, B: int):
    """LR image (B,4,16,64): U[0,1) with a {0,1} mask channel; text prior softmax(N(0,1)) (B,37,1,26)  """
    r = np.random.default_rng([seed, 41])
    x = r.uniform(0.0, 1.0, size=(B, 4, 16, 64)).astype(np.float32)
    x[:, 3] = (x[:, 3] > 0.5).astype(np.float32)
    t = r.standard_normal((B, 37, 1, 26)).astype(np.float32)
    t = np.exp(t - t.max(axis=1, keepdims=True))
    t = (t / t.sum(axis=1, keepdims=True)).astype(np.float32)
    return x, t


def main():
    shim = types.ModuleType("IPython")
    shim.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", shim)
    sys.path.insert(0, REF)
    from model import tatt as ref_tatt          # the reference package, unmodified
    from dpmn_b200.synth import synth_value
    torch.manual_seed(0)
    m = ref_tatt.TSRN_TL_TRANS(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
    sd = m.state_dict()
    shapes = {k: list(v.shape) for k, v in sd.items()}
    new = {}
    for k, v in sd.items():
        if k.startswith("stn_head.") or k.startswith("tps.") or k.endswith("pe.pe"):
            new[k] = v                           # training-only STN / TPS; the positional table is computed, not learned
        else:
            new[k] = torch.from_numpy(np.asarray(synth_value(SEED, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape)
    m.load_state_dict(new, strict=True)
    m.eval()
    save = {"shapes": np.frombuffer(json.dumps(shapes).encode(), dtype=np.uint8), "seed": np.int64(SEED)}
    for B in (1, 3):
        x, t = tatt_inputs(SEED + B, B)
        with torch.no_grad():
            y, w = m(torch.from_numpy(x), torch.from_numpy(t))
        save[f"y_b{B}"] = y.numpy()
        save[f"w_b{B}"] = w.numpy()
        print(B, y.shape, w.shape, float(y.abs().max()), float(y.std()))
    np.savez_compressed(OUT, **save)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
