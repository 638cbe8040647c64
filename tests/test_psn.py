"""Frozen TATT PSN (dpmn_b200/psn.py) against outputs of the UNMODIFIED reference
`TSRN_TL_TRANS` (fixture minted by oracle/make_golden_tatt.py).  The CPU test pins the restatement itself; the GPU test
pins the CUDA-graph replay the bench's "PSN included" leg uses."""
import json
import os

import numpy as np
import pytest
import torch

from dpmn_b200.psn import TATT
from dpmn_b200.synth import synth_value
from oracle.make_golden_tatt import tatt_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tatt.npz")


def _build():
    z = np.load(GOLD)
    shapes = json.loads(bytes(z["shapes"]).decode())
    seed = int(z["seed"])
    m = TATT()
    own = m.state_dict()
    ref_keys = {k for k in shapes if not (k.startswith("stn_head.") or k.startswith("tps."))}
    assert set(own) == ref_keys                       # same names as the reference checkpoint (minus the training-only STN)
    sd = {}
    for k, v in own.items():
        assert list(v.shape) == shapes[k], k
        sd[k] = v if k.endswith("pe.pe") else torch.from_numpy(np.asarray(synth_value(seed, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape)
    full = dict(sd)
    full["stn_head.stn_fc2.weight"] = torch.zeros(1)  # reference-only entries are skipped by the loader
    full["tps.inverse_kernel"] = torch.zeros(1)
    m.load_reference_state_dict(full)
    return m, z, seed


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_tatt_matches_the_reference_on_cpu():
    m, z, seed = _build()
    for B in (1, 3):
        x, t = tatt_inputs(seed + B, B)
        with torch.no_grad():
            y, w = m(torch.from_numpy(x), torch.from_numpy(t))
        assert _rel(y.numpy(), z[f"y_b{B}"]) < 1e-5
        assert _rel(w.numpy(), z[f"w_b{B}"]) < 1e-5
    with pytest.raises(RuntimeError):
        m.train()


@pytest.mark.gpu
def test_tatt_graph_replay_matches_the_reference_on_gpu():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m, z, seed = _build()
    m = m.cuda()
    for B, slot in ((3, 0), (1, 0), (3, 0), (3, 1), (3, 1)):     # re-capture on a batch change, replay on a repeat, a second slot
        x, t = tatt_inputs(seed + B, B)
        y, w = m.graphed(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), slot=slot)
        assert _rel(y.cpu().numpy(), z[f"y_b{B}"]) < 2e-5
        assert _rel(w.cpu().numpy(), z[f"w_b{B}"]) < 2e-5
