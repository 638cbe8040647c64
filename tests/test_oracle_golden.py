"""The numpy oracle is pinned against outputs of the unmodified reference (tests/golden, minted by
oracle/make_golden.py).  Tolerance 2e-5 on max|d|/max|ref| (fp32 reassociation between torch/oneDNN
and numpy/BLAS)."""
import numpy as np
import pytest

from oracle import cmm_oracle, pgrm_oracle
from tests.util import CMM_GOLDEN, PGRM_GOLDEN, cmm_case, load_golden, pgrm_case, rel_err

TOL = 2e-5


@pytest.mark.parametrize("name", PGRM_GOLDEN)
def test_pgrm_oracle_matches_reference(name):
    z, meta = load_golden(name)
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    y = pgrm_oracle.pgrm_forward(P, x_q, x_kv, res, windows=cfg.window_size, num_heads=cfg.num_heads)
    assert y.shape == z["out"].shape
    assert rel_err(y, z["out"]) < TOL


@pytest.mark.parametrize("name", ["pgrm_i0_m0", "pgrm_w16_c192"])
def test_pgrm_oracle_stage_probes(name):
    """attention core (pre-SK, window-major order: quirk 1) and both block outputs, image 0."""
    z, meta = load_golden(name)
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    x_q, x_kv = x_q[:1], x_kv[:1]
    if x_q.shape[1] == 2:
        x_q = pgrm_oracle.conv2d(x_q, P["prior_fusion.weight"], P["prior_fusion.bias"], pad=1)
    tq = pgrm_oracle.patch_embed(x_q, P, 2)
    tkv = pgrm_oracle.patch_embed(x_kv, P, 2)
    H, W = cfg.grid
    for b in range(2):
        tkv, parts = pgrm_oracle.swin_block(tq, tkv, P, f"layers.0.blocks.{b}.", b, cfg.window_size, H, W,
                                            cfg.heads_per_group, return_parts=True)
        assert rel_err(parts["attn_core"], z[f"attn_core_b{b}"]) < TOL
        assert rel_err(tkv, z[f"block{b}_out"]) < TOL


def test_closed_form_buffers_match_reference():
    """relative_position_index (pgrm.py:133-145) and the {0,-100} shift masks (pgrm.py:153-173)."""
    z = np.load(__import__("os").path.join(__import__("tests.util", fromlist=["GOLDEN"]).GOLDEN, "pgrm_buffers_248.npz"))
    for g, ws in enumerate((2, 4, 8)):
        assert np.array_equal(pgrm_oracle.relative_position_index(ws), z[f"index_{g}"].astype(np.int64))
        m = pgrm_oracle.shift_mask(16, 64, ws, ws // 2)
        assert np.array_equal((m != 0).astype(np.uint8), z[f"mask_{g}"])
        assert set(np.unique(m)) <= {0.0, -100.0}


@pytest.mark.parametrize("name", CMM_GOLDEN)
def test_cmm_oracle_matches_reference(name):
    z, meta = load_golden(name)
    P, x1, x2 = cmm_case(meta)
    y = cmm_oracle.cmm_forward(P, x1, x2, training=meta["train"])
    assert rel_err(y, z["out"]) < TOL


def test_residual_zero_is_skipped():
    """quirk 3 (pgrm.py:563): residual_list[0] never contributes."""
    z, meta = load_golden("pgrm_i2_m0")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    res2 = [res[0] * 0 + 123.0, res[1]]
    a = pgrm_oracle.pgrm_forward(P, x_q[:1], x_kv[:1], [r[:1] for r in res], windows=cfg.window_size)
    b = pgrm_oracle.pgrm_forward(P, x_q[:1], x_kv[:1], [r[:1] for r in res2], windows=cfg.window_size)
    assert np.array_equal(a, b)


# ---- the torch-CPU restatement (the timed CPU baseline) is pinned against the same fixtures -------------
def _torch_params(P):
    import torch
    return {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}


@pytest.mark.parametrize("name", PGRM_GOLDEN)
def test_torch_port_pgrm_matches_reference(name):
    import torch
    from oracle import torch_ref
    z, meta = load_golden(name)
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    with torch.no_grad():
        y = torch_ref.pgrm_forward(_torch_params(P), torch.from_numpy(x_q), torch.from_numpy(x_kv),
                                   [torch.from_numpy(r) for r in res], windows=cfg.window_size, num_heads=cfg.num_heads)
    assert rel_err(y.numpy(), z["out"]) < TOL


@pytest.mark.parametrize("name", CMM_GOLDEN)
def test_torch_port_cmm_matches_reference(name):
    import torch
    from oracle import torch_ref
    z, meta = load_golden(name)
    P, x1, x2 = cmm_case(meta)
    with torch.no_grad():
        y = torch_ref.cmm_forward(_torch_params(P), torch.from_numpy(x1), torch.from_numpy(x2), training=meta["train"])
    assert rel_err(y.numpy(), z["out"]) < TOL


# ---- backward: autograd through the torch oracle vs the reference's own gradients ------------------------
import pytest  # noqa: E402

from tests.util import (CMM_GRAD_GOLDEN, PGRM_GRAD_GOLDEN, golden_grad_view, load_golden, rel_err,  # noqa: E402
                        torch_ref_cmm_grads, torch_ref_pgrm_grads)


def _check_grads(z, meta, grads, tol):
    n = 0
    for key in z.files:
        if not key.startswith("g:"):
            continue
        name = key[2:]
        want = z[key]
        got = grads.get(name)
        if got is None:   # parameter without gradient in the reference too (unused weight_list_i)
            assert want.size == 1 and float(np.abs(want).max()) == 0.0, name
            continue
        full = meta["full"] or name in ("x_kv", "x1", "x2") or name.startswith("res")
        got = golden_grad_view(got, full)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if float(np.abs(want).max()) == 0.0:
            assert float(np.abs(got).max()) < 1e-6, name
        else:
            assert rel_err(got, want) < tol, (name, rel_err(got, want))
        n += 1
    assert n > 10


@pytest.mark.parametrize("name", PGRM_GRAD_GOLDEN)
def test_torch_oracle_pgrm_grads_match_reference(name):
    z, meta = load_golden(name)
    y, grads = torch_ref_pgrm_grads(meta)
    assert rel_err(y, z["out"]) < 2e-5
    _check_grads(z, meta, grads, 2e-4)


@pytest.mark.parametrize("name", CMM_GRAD_GOLDEN)
def test_torch_oracle_cmm_grads_match_reference(name):
    z, meta = load_golden(name)
    y, grads = torch_ref_cmm_grads(meta)
    assert rel_err(y, z["out"]) < 2e-5
    _check_grads(z, meta, grads, 2e-4)
