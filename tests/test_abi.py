"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dpmn_b200.h declares, and the ctypes structs have the compiled layout.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dpmn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpmn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = _declared_symbols()
    for must in ("dpmn_pgrm_forward", "dpmn_cmm_forward", "dpmn_window_attn_forward", "dpmn_pgrm_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from dpmn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/dpmn_b200.h but not exported"
    assert set(_lib.SYMBOLS) == set(_declared_symbols())


def test_struct_layouts_match_the_compiled_library():
    from dpmn_b200 import _lib
    lib = _lib.load()   # raises on any sizeof mismatch
    assert lib.dpmn_version().startswith(b"dpmn_b200")
    assert lib.dpmn_abi_sizeof(1) == ctypes.sizeof(_lib.PgrmDesc)
    assert lib.dpmn_abi_sizeof(4) == ctypes.sizeof(_lib.CmmDesc)


def test_workspace_query_and_argument_errors_need_no_gpu():
    from dpmn_b200 import _lib
    lib = _lib.load()
    d = _lib.PgrmDesc()
    d.batch, d.img_h, d.img_w, d.patch, d.q_chans = 48, 32, 128, 2, 3
    d.embed_dim, d.num_heads, d.n_groups, d.mlp_hidden, d.hidden_size, d.n_mix = 96, 6, 3, 384, 3, 1
    for g, ws in enumerate((2, 4, 8)):
        d.window[g] = ws
    assert lib.dpmn_pgrm_workspace_bytes(ctypes.byref(d)) > 48 * 1024 * 96 * 4
    d.window[2] = 32          # larger than min(H, W) = 16: the reference itself cannot run this (pgrm.py:234-236)
    assert lib.dpmn_pgrm_workspace_bytes(ctypes.byref(d)) == 0
    d.window[2] = 3           # needs the reference's broken pad path
    assert lib.dpmn_pgrm_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.dpmn_pgrm_forward(ctypes.byref(d), None, None, None, None, 0, None) == -2


def test_modules_refuse_cpu_tensors():
    import torch
    from dpmn_b200 import PGRM, ComplementationModulationModule
    m = PGRM(hidden_size=3).eval()
    x = torch.zeros(1, 3, 32, 128)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, x, [])
    c = ComplementationModulationModule(cnum=8).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        c(x, x)


def test_state_dict_keys_match_the_reference_dump():
    import json
    from dpmn_b200 import PGRM, ComplementationModulationModule
    with open(os.path.join(ROOT, "tests", "golden", "reference_state_dict_schema.json")) as f:
        dump = json.load(f)
    for it, mode in ((0, False), (2, False), (5, True)):
        n = it + 1
        m = PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                 window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[0.] * n, attn_drop_rate=[0.] * n,
                 drop_path_rate=[0.] * n, iter=it, mode=mode, hidden_size=3)
        ours = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert ours == dump[f"pgrm_iter{it}_mode{int(mode)}"]
        assert sum(p.numel() for p in m.parameters()) == dump[f"pgrm_iter{it}_mode{int(mode)}_nparams"]
    c = ComplementationModulationModule()
    assert {k: list(v.shape) for k, v in c.state_dict().items()} == dump["cmm_cnum64"]
    assert sum(p.numel() for p in c.parameters()) == dump["cmm_cnum64_nparams"]


def test_closed_form_buffers_equal_the_reference_buffers():
    import numpy as np
    from dpmn_b200.pgrm import relative_position_index, shift_mask
    z = np.load(os.path.join(ROOT, "tests", "golden", "pgrm_buffers_248.npz"))
    for g, ws in enumerate((2, 4, 8)):
        assert np.array_equal(relative_position_index(ws), z[f"index_{g}"].astype(np.int64))
        assert np.array_equal((shift_mask(16, 64, ws, ws // 2) != 0).astype(np.uint8), z[f"mask_{g}"])
