"""The tail of the training step on the GPU: dpmn_clip_adam_step against torch's clip_grad_norm_ + Adam (the reference's
own calls, interfaces/super_resolution.py:270-278, base.py:208-221), dpmn_allreduce_bucket through the library's NCCL
communicator (world size 1 here; N > 1 runs in bench.py --gpus N), the prepared-weights cache across train / eval forwards
 and the eval-time alpha blend (super_resolution.py:449,705)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.util import build_cmm, build_pgrm, cmm_case, load_golden, pgrm_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_fused_clip_adam_matches_torch_clip_grad_norm_and_adam():
    from dpmn_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    sizes = [1000, 37, 4096 + 3, 5]                         # segments with ragged lengths (padded to multiples of 4)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + (s + 3) // 4 * 4)
    total = offs[-1]
    p = torch.zeros(total, device=DEV)
    g = torch.zeros(total, device=DEV)
    refs = []
    for i, s in enumerate(sizes):
        w = torch.randn(s, device=DEV)
        p[offs[i]: offs[i] + s] = w
        refs.append(torch.nn.Parameter(w.clone()))
    opt = torch.optim.Adam(refs, lr=1e-3, betas=(0.5, 0.999))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ws = torch.zeros(lib.dpmn_clip_adam_workspace_bytes(len(sizes)) + 16, dtype=torch.uint8, device=DEV)
    c_offs = (C.c_int64 * (len(sizes) + 1))(*offs)
    world = 4
    for step in range(1, 6):
        scale = [0.01, 3.0, 0.2, 50.0][step % 4]            # some segments above, some below the 0.25 threshold
        g.zero_()
        for i, s in enumerate(sizes):
            gi = torch.randn(s, device=DEV) * scale * (i + 1)
            g[offs[i]: offs[i] + s] = gi * world            # the bucket holds the SUM over ranks; grad_scale = 1 / world
            refs[i].grad = gi.clone()
        for r in refs:                                      # one module per segment, as the reference clips
            torch.nn.utils.clip_grad_norm_([r], 0.25)
        opt.step()
        rc = lib.dpmn_clip_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), c_offs, len(sizes), 1.0 / world,
                                     0.25, 1e-3, 0.5, 0.999, 1e-8, step, ws.data_ptr(), ws.numel(),
                                     torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
        for i, s in enumerate(sizes):
            got = p[offs[i]: offs[i] + s]
            assert float((got - refs[i].data).abs().max()) < 2e-6, (step, i)
    # padding between segments is never touched
    for i, s in enumerate(sizes):
        assert float(p[offs[i] + s: offs[i + 1]].abs().max() if offs[i + 1] > offs[i] + s else 0.0) == 0.0
    # argument errors
    assert lib.dpmn_clip_adam_step(None, g.data_ptr(), m.data_ptr(), v.data_ptr(), c_offs, len(sizes), 1.0, 0.25, 1e-3, 0.5,
                                   0.999, 1e-8, 1, ws.data_ptr(), ws.numel(), None) == -1
    assert lib.dpmn_clip_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), c_offs, len(sizes), 1.0, 0.25, 1e-3,
                                   0.5, 0.999, 1e-8, 1, ws.data_ptr(), 8, None) == -3


def test_allreduce_bucket_through_the_library_nccl_comm_world1():
    from dpmn_b200 import _lib
    lib = _lib.load()
    assert lib.dpmn_nccl_available() == 1 and lib.dpmn_nccl_version() >= 22000
    idb = (C.c_char * 128)()
    assert lib.dpmn_nccl_unique_id(C.cast(idb, C.c_void_p)) == 0
    comm = C.c_void_p()
    assert lib.dpmn_nccl_comm_init(C.cast(idb, C.c_void_p), 1, 0, C.byref(comm)) == 0
    x = torch.randn(1 << 20, device=DEV)
    want = x.clone()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    assert lib.dpmn_allreduce_bucket(comm, x.data_ptr(), x.numel(), 0, st.cuda_stream) == 0
    st.synchronize()
    assert torch.equal(x, want)                             # sum over one rank
    assert lib.dpmn_allreduce_bucket(comm, None, 4, 0, None) == -1
    assert lib.dpmn_allreduce_bucket(comm, x.data_ptr(), 4, 7, None) == -1
    assert lib.dpmn_nccl_comm_destroy(comm) == 0


def test_trainer_fused_tail_equals_the_torch_tail_over_three_steps():
    """HotPathTrainer with the fused clip + Adam kernels and direct-to-bucket gradients against the same trainer driving
    torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (fp32, no dropout: both runs are deterministic)."""
    import bench
    from dpmn_b200.pipeline import DPMNHotPath
    from dpmn_b200.train import HotPathTrainer
    pg, cm = bench.synth_weights(2)
    B = 2
    psn, p1, p2 = bench.synth_inputs(3, B)
    hr = _t(np.random.default_rng(9).uniform(0, 1, (B, 4, 32, 128)).astype(np.float32))
    args = (_t(psn), [_t(a) for a in p1], [_t(a) for a in p2], hr)
    finals = []
    for fused in (True, False):
        model = DPMNHotPath(precision="fp32", drop=0.0)
        bench.load_weights(model, pg, cm)
        model = model.to(DEV).train()
        tr = HotPathTrainer(model, fused_optimizer=fused)
        losses = [float(tr.step(*args)) for _ in range(3)]
        finals.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
    (l0, w0), (l1, w1) = finals
    assert np.allclose(l0, l1, rtol=2e-4), (l0, l1)
    assert l0[2] < l0[0]                                     # and the step actually trains
    # Adam's first steps move every weight by ~lr = 1e-3 in the direction sign(g): where a gradient element is within fp32
    # summation noise of 0 (atomics order differs from run to run) the two runs may legitimately step in opposite directions,
    # so the bar is statistical per tensor (the exact check of the kernel is the stand-alone test above, 2e-6)
    for n in w0:
        d = (w0[n] - w1[n]).abs()
        assert float(d.max()) < 3 * 2e-3 and float(d.mean()) < 2e-4, (n, float(d.max()), float(d.mean()))


def test_train_mode_forward_does_not_validate_the_prepared_weight_cache():
    """ADVICE r1 (pgrm.py:396): a train-mode forward with non-zero drop rates runs the fp32 training sequence, which never
    stages the 16-bit weights; an eval() forward right after it must stage them itself and match the oracle."""
    from oracle import pgrm_oracle
    z, meta = load_golden("pgrm_i2_m0")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, DEV, precision="fp16")
    m.train()
    with torch.no_grad():
        m(_t(x_q), _t(x_kv), [_t(r) for r in res])           # stochastic path (rates 0.1): must not mark the cache valid
    assert m._prepared.key is None
    m.eval()
    with torch.no_grad():
        y = m(_t(x_q), _t(x_kv), [_t(r) for r in res])
    ref = pgrm_oracle.pgrm_forward(P, x_q, x_kv, res, windows=cfg.window_size, num_heads=cfg.num_heads)
    assert rel_err(y.cpu().numpy(), ref) < 1e-3
    assert m._prepared.key is not None


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("fp16", 1e-3)])
def test_cmm_alpha_blend_is_the_reference_expression(prec, tol):
    """image_sr = alpha * CMM(b1, b2) + (1 - alpha) * images_lr_psn[:, :3]  (super_resolution.py:449,705), blended inside
    the kernel that writes the output; blend_with is the channel-slice view of a 4-channel tensor as at the call site."""
    from oracle import cmm_oracle
    zc, metac = load_golden("cmm_c64_eval" if prec == "fp16" else "cmm_c8_eval")
    Pc, x1, x2 = cmm_case(metac)
    c, _ = build_cmm(metac, DEV, precision=prec)
    c.eval()
    psn = np.random.default_rng(11).uniform(0, 1, (x1.shape[0], 4, 32, 128)).astype(np.float32)
    with torch.no_grad():
        y = c(_t(x1), _t(x2), blend_with=_t(psn)[:, :3], alpha=0.3)
    ref = 0.3 * cmm_oracle.cmm_forward(Pc, x1, x2, training=False) + 0.7 * psn[:, :3]
    assert rel_err(y.cpu().numpy(), ref) < tol


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_two_stream_training_step_gives_the_one_stream_gradients(prec):
    """DPMNHotPath.concurrent_train: under autograd the two PGRM cascades run on two CUDA streams (every backward node on
    the stream of its forward).  Same seeds -> same Dropout / DropPath masks, so the seven output images, the loss and the
    flat gradient bucket must equal the one-stream step's up to the run-to-run noise of the kernels' fp32 atomics (pooled
    sums, BatchNorm statistics, weight gradients) -- which the 16-bit modes amplify to 16-bit rounding flips.  The control
    is the one-stream step run twice: the two-stream step may differ from it by no more than that noise (x3) + 1e-6."""
    import bench
    from dpmn_b200.pipeline import DPMNHotPath
    from dpmn_b200.train import HotPathTrainer
    pg, cm = bench.synth_weights(2)
    B = 4
    psn, p1, p2 = bench.synth_inputs(3, B)
    hr = _t(np.random.default_rng(9).uniform(0, 1, (B, 4, 32, 128)).astype(np.float32))
    args = (_t(psn), [_t(a) for a in p1], [_t(a) for a in p2], hr)
    got = []
    for two_streams in (True, False, False):
        torch.manual_seed(1234)                                  # the mask seeds come from the CPU generator
        model = DPMNHotPath(precision=prec, drop=0.1)
        bench.load_weights(model, pg, cm)
        model = model.to(DEV).train()
        model.concurrent_train = two_streams
        tr = HotPathTrainer(model)
        tr.state.zero_grads()
        outs = model.forward_all(*args[:3])
        loss = tr.loss(outs, hr)
        loss.backward()
        torch.cuda.synchronize()
        got.append((float(loss.detach()), tr.state.flat_grads.detach().clone(), [o.detach().clone() for o in outs]))
    (l2, g2, o2), (l1, g1, o1), (lc, gc, oc) = got

    def rel(a, b):
        return float((a - b).abs().max()) / float(b.abs().max())
    noise_o = max(rel(a, b) for a, b in zip(oc, o1))
    noise_g = rel(gc, g1)
    diff_o = max(rel(a, b) for a, b in zip(o2, o1))
    diff_g = rel(g2, g1)
    print(f"{prec}: outputs two-vs-one {diff_o:.3e} (control {noise_o:.3e}); gradients {diff_g:.3e} (control {noise_g:.3e}); "
          f"loss {l2:.7f} / {l1:.7f} / {lc:.7f}")
    assert float(g1.abs().max()) > 0
    # floor: which kernels run next to each other changes the order of fp32 atomics (pooled sums, BatchNorm statistics) even when
    # two one-stream runs happen to be bit-identical (under compute-sanitizer they are); a 16-bit mode turns that last-bit noise
    # into 16-bit rounding flips (2^-11 relative per flipped value)
    floor_o = 1e-6 if prec == "fp32" else 5e-4
    assert diff_o <= 3 * noise_o + floor_o
    assert diff_g <= 3 * noise_g + 2e-4
    assert abs(l2 - l1) <= 3 * abs(lc - l1) + 1e-6 * abs(l1)
