"""Edge cases of the drop-in boundary on the GPU (the reference ships no tests; these are the cases its own code
singles out): batch 1 (SKConv's `feats_S.squeeze()` drops the batch dimension there, pgrm.py:87 -- results must not
change), batch sizes that are not multiples of anything, the error behaviour of the reference's asserts, and the C-ABI's
return codes with real device pointers."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import cmm_oracle, pgrm_oracle
from tests.util import build_cmm, build_pgrm, cmm_case, load_golden, pgrm_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("fp16", 1e-3)])
def test_batch_one_matches_oracle(prec, tol):
    z, meta = load_golden("pgrm_i2_m0")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, DEV, precision=prec)
    m.eval()
    with torch.no_grad():
        y = m(_t(x_q[:1]), _t(x_kv[:1]), [_t(r[:1]) for r in res])
    ref = pgrm_oracle.pgrm_forward(P, x_q[:1], x_kv[:1], [r[:1] for r in res], windows=cfg.window_size,
                                   num_heads=cfg.num_heads)
    assert y.shape == (1, 3, 32, 128)
    assert rel_err(y.cpu().numpy(), ref) < tol
    zc, metac = load_golden("cmm_c8_eval")
    Pc, x1, x2 = cmm_case(metac)
    c, _ = build_cmm(metac, DEV, precision="fp32")
    c.eval()
    with torch.no_grad():
        yc = c(_t(x1[:1]), _t(x2[:1]))
    assert rel_err(yc.cpu().numpy(), cmm_oracle.cmm_forward(Pc, x1[:1], x2[:1], training=False)) < 1e-5


@pytest.mark.parametrize("B", [3, 5, 7])
def test_odd_batch_sizes_equal_per_image_runs(B):
    """fp16 tensor-core path: B * 1024 rows are tiled by 128, images never share a reduction."""
    z, meta = load_golden("pgrm_i3_m1")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    reps = (B + x_q.shape[0] - 1) // x_q.shape[0]
    r = np.random.default_rng(B)
    xq = np.concatenate([x_q] * reps)[:B]
    xkv = (np.concatenate([x_kv] * reps)[:B] + r.uniform(0, 0.1, (B, 3, 32, 128))).astype(np.float32)
    rs = [np.concatenate([t] * reps)[:B] for t in res]
    m, _ = build_pgrm(meta, DEV, precision="fp16")
    m.eval()
    with torch.no_grad():
        y = m(_t(xq), _t(xkv), [_t(t) for t in rs])
        for i in (0, B - 1):
            yi = m(_t(xq[i:i + 1]), _t(xkv[i:i + 1]), [_t(t[i:i + 1]) for t in rs])
            assert torch.equal(y[i:i + 1], yi), i


def test_reference_error_behaviour():
    z, meta = load_golden("pgrm_i2_m0")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, DEV)
    m.eval()
    with torch.no_grad():
        with pytest.raises(AssertionError):                      # pgrm.py:421 "Input image size ... doesn't match model"
            m(_t(x_q[:1, :, :16]), _t(x_kv[:1, :, :16]), [])
        with pytest.raises(AttributeError):                      # pgrm.py:564: no weight_list_3 in a PGRM of iter 2
            m(_t(x_q[:1]), _t(x_kv[:1]), [_t(x_kv[:1])] * 4)
        with pytest.raises((RuntimeError, ValueError)):          # fp64 input: the reference's convs would raise too
            m(_t(x_q[:1]).double(), _t(x_kv[:1]).double(), [])
    zc, metac = load_golden("cmm_c8_eval")
    c, _ = build_cmm(metac, DEV)
    with pytest.raises(ValueError):
        c(torch.zeros(1, 3, 32, 128, device=DEV), torch.zeros(1, 3, 16, 64, device=DEV))


def test_cmm_train_mode_single_32x32_image_runs_like_the_reference():
    """B = 1 at 32x32 in train(): the reference does NOT raise -- the 1x1 bottleneck (en_6, cmm.py:91-93) has no BatchNorm and
    the smallest normalised maps (en_5, de_6) are 2x2 = 4 values per channel.  Checked against the oracle (batch statistics)."""
    from oracle import cmm_oracle
    zc, metac = load_golden("cmm_c8_train")
    Pc, _, _ = cmm_case(metac)
    c, _ = build_cmm(metac, DEV)
    c.train()
    r = np.random.default_rng(5)
    x1 = r.uniform(0, 1, (1, 3, 32, 32)).astype(np.float32)
    x2 = r.uniform(0, 1, (1, 3, 32, 32)).astype(np.float32)
    with torch.no_grad():
        y = c(_t(x1), _t(x2))
    ref = cmm_oracle.cmm_forward(Pc, x1, x2, training=True)
    assert rel_err(y.cpu().numpy(), ref) < 1e-4


def test_c_abi_return_codes_with_device_pointers():
    from dpmn_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    a = torch.rand(2, 3, 32, 128, device=DEV)
    out = torch.empty(2, 3, 32, 128, device=DEV)
    loss = torch.zeros((), device=DEV)
    assert lib.dpmn_image_loss(a.data_ptr(), 0, None, 0, 2, 3, 32, 128, 1.0, 1.0, 1.0, loss.data_ptr(), None, st) == -1
    assert lib.dpmn_to_mask(a.data_ptr(), 0, out.data_ptr(), 0, 32, 128, st) == -1
    assert lib.dpmn_crnn_input(a.data_ptr(), 0, None, 2, 32, 128, 32, 100, st) == -1
    assert lib.dpmn_visionlan_input(a.data_ptr(), 0, out.data_ptr(), 2, 32, 128, 0, 256, st) == -1
    z, meta = load_golden("pgrm_i3_m1")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, DEV)
    d = m._descriptor(2, 3, 1)
    need = lib.dpmn_pgrm_workspace_bytes(C.byref(d))
    assert need > 0
    ws = torch.empty(need // 2, dtype=torch.uint8, device=DEV)
    rc = lib.dpmn_pgrm_forward(C.byref(d), a.data_ptr(), a.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert rc == -3                                              # DPMN_E_WORKSPACE
    assert lib.dpmn_pgrm_forward(C.byref(d), None, a.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st) == -1
    torch.cuda.synchronize()
    assert lib.dpmn_check_device() == 0
