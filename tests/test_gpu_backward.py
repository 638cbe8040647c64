"""GPU parity of the backward kernels (dpmn_pgrm_backward / dpmn_cmm_backward, through the autograd Functions of the
drop-in modules) against the reference's own autograd gradients (tests/golden/*_grad.npz, minted by
oracle/make_golden_grads.py).  Metric: max|d| / max|ref| per gradient tensor; bar 2e-4 (fp32 arithmetic, split-K /
atomic accumulation order differs from torch's)."""
import numpy as np
import pytest
import torch

from tests.util import (CMM_GRAD_GOLDEN, PGRM_GRAD_GOLDEN, build_cmm, build_pgrm, cmm_case, golden_grad_view,
                        grad_seed_out, load_golden, pgrm_case, rel_err)

pytestmark = pytest.mark.gpu
TOL = 2e-4


NOISE_FLOOR = 1e-2   # see test_cmm_backward_matches_reference


def _compare(z, meta, grads, tol=TOL, noise_floor=0.0):
    worst, n = [], 0
    for key in z.files:
        if not key.startswith("g:"):
            continue
        name = key[2:]
        want = z[key]
        got = grads.get(name)
        if got is None:
            assert want.size == 1 and float(np.abs(want).max()) == 0.0, f"{name}: no gradient produced"
            continue
        full = meta["full"] or name in ("x_kv", "x1", "x2") or name.startswith("res")
        got = golden_grad_view(got, full)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        scale = float(np.abs(want).max())
        if scale < noise_floor:       # a gradient that is exactly 0 in exact arithmetic: both sides are rounding noise
            assert float(np.abs(got).max()) < noise_floor, (name, float(np.abs(got).max()))
            continue
        err = float(np.abs(got - want).max()) if scale == 0.0 else rel_err(got, want)
        worst.append((err, name))
        n += 1
    worst.sort(reverse=True)
    bad = [(e, k) for e, k in worst if not e < tol]
    assert not bad, f"{len(bad)}/{n} gradient tensors out of tolerance; worst: {worst[:12]}"
    assert n > 10


@pytest.mark.parametrize("name", PGRM_GRAD_GOLDEN)
def test_pgrm_backward_matches_reference(name):
    z, meta = load_golden(name)
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, "cuda", precision="fp32")
    dev = torch.device("cuda")
    xq = torch.from_numpy(x_q).to(dev)
    xkv = torch.from_numpy(x_kv).to(dev).requires_grad_(True)
    rs = [torch.from_numpy(r).to(dev).requires_grad_(True) for r in res]
    y = m(xq, xkv, rs)
    assert y.requires_grad
    assert rel_err(y.detach().cpu().numpy(), z["out"]) < 2e-5
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)).sum().backward()
    grads = {k: (p.grad.cpu().numpy() if p.grad is not None else None) for k, p in m.named_parameters()}
    grads["x_kv"] = xkv.grad.cpu().numpy()
    for i, r in enumerate(rs):
        grads[f"res{i}"] = r.grad.cpu().numpy() if r.grad is not None else None
    _compare(z, meta, grads)


def test_pgrm_backward_sliced_x_kv_and_accumulation():
    """x_kv as a channel-slice view of a 4-channel tensor (super_resolution.py:196) and two backward passes
    accumulating into .grad like torch does."""
    z, meta = load_golden("pgrm_i3_m1_grad")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, "cuda", precision="fp32")
    dev = torch.device("cuda")
    xq = torch.from_numpy(x_q).to(dev)
    four = torch.cat([torch.from_numpy(x_kv), torch.zeros(meta["B"], 1, 32, 128)], dim=1).to(dev).requires_grad_(True)
    G = torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)
    for _ in range(2):
        (m(xq, four[:, :3], []) * G).sum().backward()
    g = four.grad.cpu().numpy()
    assert float(np.abs(g[:, 3]).max()) == 0.0
    assert rel_err(g[:, :3], 2.0 * z["g:x_kv"]) < TOL
    w = m.get_parameter("layers.0.blocks.1.mlp.fc2.weight").grad.cpu().numpy()
    assert rel_err(golden_grad_view(w, False), 2.0 * z["g:layers.0.blocks.1.mlp.fc2.weight"]) < TOL


@pytest.mark.parametrize("name", CMM_GRAD_GOLDEN)
def test_cmm_backward_matches_reference(name):
    z, meta = load_golden(name)
    P, x1, x2 = cmm_case(meta)
    m, _ = build_cmm(meta, "cuda", precision="fp32")
    dev = torch.device("cuda")
    a = torch.from_numpy(x1).to(dev).requires_grad_(True)
    b = torch.from_numpy(x2).to(dev).requires_grad_(True)
    y = m(a, b)
    assert rel_err(y.detach().cpu().numpy(), z["out"]) < 2e-5
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)).sum().backward()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    grads["x1"], grads["x2"] = a.grad.cpu().numpy(), b.grad.cpu().numpy()
    # a conv bias followed by a train-mode BatchNorm has an exactly-zero true gradient; the reference's own value is
    # fp32 cancellation noise (<= 4e-4 where real gradients are >= 1), so those are only required to be noise too
    _compare(z, meta, grads, tol=5e-4 if meta["train"] else TOL, noise_floor=NOISE_FLOOR if meta["train"] else 0.0)


def test_hot_path_training_step_gradients_vs_oracle():
    """Whole path (6 PGRM cascade with gradients flowing through x_kv and the residual lists, CMM in train mode,
    the reference's 7-term image loss): gradients vs autograd through the torch oracle on CPU.  At cnum 64 a CMM
    forward has ~6M ReLU / LeakyReLU inputs, so a few sit within rounding distance of 0 and their masks may differ
    between two valid fp32 evaluations (see oracle/make_golden_grads.py); the bar here is therefore a relative L2
    error of 3e-2 and cosine > 0.999 per tensor -- exact parity is covered by the per-module fixtures above."""
    import bench
    from dpmn_b200.pipeline import DPMNHotPath
    from dpmn_b200.train import HotPathTrainer
    from oracle import torch_ref
    B = 2
    dev = torch.device("cuda")
    model = DPMNHotPath(precision="fp32", drop=0.0)
    pg, cm = bench.synth_weights(2)
    bench.load_weights(model, pg, cm)
    model = model.to(dev).train()
    tr = HotPathTrainer(model)
    psn, p1, p2 = bench.synth_inputs(3, B)
    hr = np.random.default_rng(9).uniform(0, 1, (B, 4, 32, 128)).astype(np.float32)
    outs = model.forward_all(torch.from_numpy(psn).to(dev), [torch.from_numpy(a).to(dev) for a in p1],
                             [torch.from_numpy(a).to(dev) for a in p2])
    loss = tr.loss(outs, torch.from_numpy(hr).to(dev))
    loss.backward()
    # oracle
    pgt = [{k: torch.from_numpy(np.asarray(v)).requires_grad_(True) for k, v in p.items()} for p in pg]
    cmt = {k: torch.from_numpy(np.asarray(v)) for k, v in cm.items()}
    for k, v in cmt.items():
        if v.dtype == torch.float32 and "running" not in k:
            v.requires_grad_(True)
    psn_t = torch.from_numpy(psn)
    o_outs = []
    for branch, priors in ((0, p1), (1, p2)):
        cascade, done = psn_t[:, :3], []
        for k in range(3):
            y = torch_ref.pgrm_forward(pgt[3 * branch + k], torch.from_numpy(priors[k]), cascade, done[:k])
            done.append(y)
            cascade = y
        o_outs += done
    o_outs.append(torch_ref.cmm_forward(cmt, o_outs[2], o_outs[5], training=True))
    # the four DistillModule terms (super_resolution.py:245-263), with the trainer's own distill parameters
    dps = [{k: (v.detach().cpu().clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k
                else v.detach().cpu().clone()) for k, v in m.state_dict().items()} for m in tr.distill]
    o_distill = 0
    for imgs, first in ((o_outs[:3], 0), (o_outs[3:6], 2)):
        feat = imgs[-1]
        for k in range(2, 0, -1):
            l, feat = torch_ref.distill_forward(dps[first + k - 1], feat, imgs[k - 1], training=True)
            o_distill = o_distill + l * 100
    o_loss = (sum(torch_ref.image_loss(o, torch.from_numpy(hr)[:, :3]) * 100 for o in o_outs) + o_distill) / len(o_outs)
    o_loss.backward()
    for j, m in enumerate(tr.distill):
        for n, p in m.named_parameters():
            a, b = p.grad.cpu().double().flatten(), dps[j][n].grad.double().flatten()
            if n in ("conv_cat_feature.bias", "conv_feature.bias"):      # identically zero through a train-mode BatchNorm
                assert float(a.abs().max()) < 1e-4 and float(b.abs().max()) < 1e-4, (j, n)
                continue
            assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 3e-2, ("distill", j, n)
    assert abs(float(loss.detach()) - float(o_loss.detach())) < 1e-4 * abs(float(o_loss.detach()))
    checked = 0
    for k, m in enumerate(model.pgrm):
        for n, p in m.named_parameters():
            ref = pgt[k][n].grad
            if ref is None:
                assert float(p.grad.abs().max()) == 0.0, (k, n)
                continue
            a, b = p.grad.cpu().double().flatten(), ref.double().flatten()
            assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 3e-2, (k, n)
            assert float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)) > 0.999, (k, n)
            checked += 1
    for n, p in model.cmm.named_parameters():
        a, b = p.grad.cpu().double().flatten(), cmt[n].grad.double().flatten()
        if float(b.abs().max()) < 1e-2:
            continue
        assert float((a - b).norm() / b.norm()) < 3e-2, n
        checked += 1
    assert checked > 300


@pytest.mark.parametrize("name,rates", [("pgrm_i2_m0_grad", (0.1, 0.1, 0.1)), ("pgrm_i3_m1_grad", (0.2, 0.0, 0.3)),
                                        ("pgrm_i5_m1_grad", (0.0, 0.15, 0.0))])
def test_pgrm_train_mode_dropout_droppath_forward_and_backward(name, rates):
    """module.train() with non-zero drop_rate / attn_drop_rate / drop_path_rate (pgrm.py:32,40,248,
    329-330,554-555).  The reference's RNG stream cannot be reproduced, so the oracle (oracle/torch_ref.py) applies
    the masks the CUDA path derives from (seed, site, index) -- restated in numpy -- at the reference's own Dropout /
    DropPath sites; forward and every gradient must then agree like the eval-mode fixtures do."""
    from oracle import torch_ref
    z, meta = load_golden(name)
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    it = meta["iter"]
    n = it + 1
    from dpmn_b200 import PGRM
    drop, attn, path = rates
    m = PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n, window_size=[[2, 4, 8]] * n,
             mlp_ratio=[4.] * n, drop_rate=[drop] * n, attn_drop_rate=[attn] * n, drop_path_rate=[path] * n, iter=it,
             mode=meta["mode"], hidden_size=3, precision="fp32")
    sd = m.state_dict()
    m.load_state_dict({k: (torch.from_numpy(P[k]) if k in P else v) for k, v in sd.items()}, strict=True)
    dev = torch.device("cuda")
    m = m.to(dev).train()
    seed = 123456789 + it
    m._new_seed = lambda: seed
    xq = torch.from_numpy(x_q).to(dev)
    xkv = torch.from_numpy(x_kv).to(dev).requires_grad_(True)
    rs = [torch.from_numpy(r).to(dev).requires_grad_(True) for r in res]
    G = torch.from_numpy(grad_seed_out(meta["seed"], meta["B"]))
    y = m(xq, xkv, rs)
    (y * G.to(dev)).sum().backward()
    # oracle with the same masks
    cfg_drop = dict(seed=seed, drop_rate=drop, attn_drop_rate=attn, drop_path=list(m.drop_path))
    Pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in P.items()}
    oxkv = torch.from_numpy(x_kv).requires_grad_(True)
    ors = [torch.from_numpy(r).requires_grad_(True) for r in res]
    oy = torch_ref.pgrm_forward(Pt, torch.from_numpy(x_q), oxkv, ors, windows=cfg.window_size, num_heads=cfg.num_heads,
                                drop=cfg_drop)
    (oy * G).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), oy.detach().numpy()) < 2e-5
    # the masks really are active: the eval-mode output differs
    assert rel_err(y.detach().cpu().numpy(), z["out"]) > 1e-3
    assert rel_err(xkv.grad.cpu().numpy(), oxkv.grad.numpy()) < TOL
    bad = []
    for k, p in m.named_parameters():
        ref = Pt[k].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        e = rel_err(p.grad.cpu().numpy(), ref.numpy())
        if not e < TOL:
            bad.append((e, k))
    assert not bad, sorted(bad, reverse=True)[:10]
    for a, b in zip(rs[1:], ors[1:]):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) < TOL
    # a second forward draws different masks (fresh seed); the same seed reproduces the output bit for bit
    m._new_seed = lambda: seed + 1
    with torch.no_grad():
        y2 = m(xq, xkv.detach(), [r.detach() for r in rs])
        m._new_seed = lambda: seed
        y3 = m(xq, xkv.detach(), [r.detach() for r in rs])
    assert not torch.equal(y2, y.detach())
    assert torch.equal(y3, y.detach())


@pytest.mark.parametrize("precision,l2_tol,cos_min", [("fp16", 5e-2, 0.999), ("bf16", 1e-1, 0.99)])
def test_pgrm_backward_tensor_core_gemms(precision, l2_tol, cos_min):
    """16-bit modes: the GEMMs of the training forward sequence and every Linear / pointwise-conv data and weight
    gradient run on the tcgen05 GEMM with 16-bit staged operands and fp32 accumulation (the rest of the backward stays
    fp32).  The 16-bit forward moves every activation by ~5e-4, which flips the sign of a few of the head's 24 K
    LeakyReLU(0.01) inputs that sit next to 0 -- each flip changes that element's derivative 100-fold (measured: the
    same gradients are within 3e-3 max-norm when only the backward GEMMs are 16-bit, and individual tensors move by up
    to 8e-2 max-norm once a forward GEMM is, whichever GEMM it is).  The bar is therefore the flip-tolerant one of the
    whole-path test: relative L2 error and cosine per gradient tensor against the reference's fp32 gradients.
    (With the attention core on the tcgen05 kernel as well -- 16-bit q / k / v / P -- the 12-element bias gradient of the
    head's second conv, a plain sum over those LeakyReLU derivatives, sits at 3.3e-2; every other tensor is below 3e-2.  The
    TIGHT check of the 16-bit backward is test_pgrm_16bit_gradients_against_the_same_arithmetic_oracle below, where the
    flips are excluded by construction.)"""
    z, meta = load_golden("pgrm_i2_m0_grad")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    m, _ = build_pgrm(meta, "cuda", precision=precision)
    dev = torch.device("cuda")
    # batch 2 -> rows = 2048 (a multiple of the 512-row split the tensor-core weight gradient needs)
    xq = torch.from_numpy(x_q).to(dev)
    xkv = torch.from_numpy(x_kv).to(dev).requires_grad_(True)
    rs = [torch.from_numpy(r).to(dev).requires_grad_(True) for r in res]
    y = m(xq, xkv, rs)
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)).sum().backward()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    grads["x_kv"] = xkv.grad.cpu().numpy()
    n, bad = 0, []
    for key in z.files:
        if not key.startswith("g:") or key[2:] not in grads:
            continue
        want = z[key].astype(np.float64).ravel()
        got = grads[key[2:]].astype(np.float64).ravel()
        if np.abs(want).max() == 0.0:
            continue
        l2 = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        cos = float(np.dot(got, want) / (np.linalg.norm(got) * np.linalg.norm(want)))
        n += 1
        if not (l2 < l2_tol and cos > cos_min):
            bad.append((l2, cos, key[2:]))
    assert n > 60 and not bad, sorted(bad, reverse=True)[:10]


@pytest.mark.parametrize("name", ["cmm_c8_train_grad", "cmm_c16_train_grad"])
def test_cmm_backward_tensor_core_convs(name):
    """fp16 mode: the convs of the CMM's fp32-structured forward (train-mode BatchNorm), their data gradients and their
    weight gradients run on the tcgen05 GEMM through a 16-bit im2col (cmm_im2col.cu).  A 16-bit forward moves every
    pre-activation by ~1e-3, so the ReLU / LeakyReLU mask of the ~0.1 % of elements that close to 0 differs from the fp32
    reference's; each such layer adds ~sqrt(fraction flipped) ~ 2-3 % of gradient noise (measured: 1e-3 at the first
    conv of the backward chain, where no mask is involved yet, 2 % after the first BatchNorm+ReLU, saturating at
    6-9 % fifteen layers deep, up to 16 % with batch-1 BatchNorm statistics -- tools/grad_diag16.py).  The bar is therefore relative L2 < 0.2 and cosine > 0.98 per
    tensor, plus 3e-3 on the two tensors whose gradient involves no activation mask."""
    z, meta = load_golden(name)
    P, x1, x2 = cmm_case(meta)
    m, _ = build_cmm(meta, "cuda", precision="fp16")
    dev = torch.device("cuda")
    a = torch.from_numpy(x1).to(dev).requires_grad_(True)
    b = torch.from_numpy(x2).to(dev).requires_grad_(True)
    y = m(a, b)
    assert rel_err(y.detach().cpu().numpy(), z["out"]) < 2e-3
    (y * torch.from_numpy(grad_seed_out(meta["seed"], meta["B"])).to(dev)).sum().backward()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    grads["x1"], grads["x2"] = a.grad.cpu().numpy(), b.grad.cpu().numpy()
    n, bad = 0, []
    for key in z.files:
        if not key.startswith("g:") or key[2:] not in grads:
            continue
        name_ = key[2:]
        full = meta["full"] or name_ in ("x1", "x2")
        want = z[key].astype(np.float64).ravel()
        got = golden_grad_view(grads[name_], full).astype(np.float64).ravel()
        if np.abs(want).max() < NOISE_FLOOR:
            continue
        l2 = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        cos = float(np.dot(got, want) / (np.linalg.norm(got) * np.linalg.norm(want)))
        n += 1
        if not (l2 < 0.2 and cos > 0.98):
            bad.append((l2, cos, name_))
        if name_ in ("de_1.1.weight", "de_1.1.bias") and not l2 < 3e-3:
            bad.append((l2, cos, name_))
    assert n > 40 and not bad, sorted(bad, reverse=True)[:10]


def _l2_cos(got, want):
    got, want = np.asarray(got, np.float64).ravel(), np.asarray(want, np.float64).ravel()
    nw = np.linalg.norm(want)
    return float(np.linalg.norm(got - want) / max(nw, 1e-30)), float(np.dot(got, want) / max(np.linalg.norm(got) * nw, 1e-30))


@pytest.mark.parametrize("precision,fwd_tol,l2_tol", [("fp16", 5e-4, 2e-3), ("bf16", 4e-3, 2e-2)])
def test_pgrm_16bit_gradients_against_the_same_arithmetic_oracle(precision, fwd_tol, l2_tol):
    """the loose 16-bit gradient bars compared a 16-bit forward with the reference's fp32 one, so every
    LeakyReLU / Dropout mask that flipped showed up as gradient error.  Here the oracle evaluates the SAME forward
    arithmetic -- operands of the tensor-core contractions rounded to 16 bits, fp32 everywhere else, the CUDA path's
    Dropout / DropPath masks (oracle/torch_ref.operand_rounding) -- so forward and masks coincide and what is measured is
    the backward itself: its 16-bit staged operands (relative rounding 2^-11 fp16 / 2^-8 bf16 per element, averaged over
    K >= 96 terms).  Bars: relative L2 < 2e-3 (fp16) per gradient tensor, cosine > 0.9999.  Since the attention core of the
    16-bit training forward runs on the tcgen05 kernel (16-bit q / k / v / P, attn_drop inside the kernel), the oracle rounds
    those operands too and the backward recomputes P from the rounded values: measured fp16 1.4e-3, bf16 1.6e-2."""
    from oracle import torch_ref
    z, meta = load_golden("pgrm_i2_m0_grad")
    cfg, P, x_q, x_kv, res = pgrm_case(meta)
    it = meta["iter"]
    n = it + 1
    from dpmn_b200 import PGRM
    m = PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n, window_size=[[2, 4, 8]] * n,
             mlp_ratio=[4.] * n, drop_rate=[0.1] * n, attn_drop_rate=[0.1] * n, drop_path_rate=[0.1] * n, iter=it,
             mode=meta["mode"], hidden_size=3, precision=precision)
    sd = m.state_dict()
    m.load_state_dict({k: (torch.from_numpy(P[k]) if k in P else v) for k, v in sd.items()}, strict=True)
    dev = torch.device("cuda")
    m = m.to(dev).train()
    seed = 424242
    m._new_seed = lambda: seed
    xq = torch.from_numpy(x_q).to(dev)
    xkv = torch.from_numpy(x_kv).to(dev).requires_grad_(True)
    rs = [torch.from_numpy(r).to(dev).requires_grad_(True) for r in res]
    G = torch.from_numpy(grad_seed_out(meta["seed"], meta["B"]))
    Pt = {k: torch.from_numpy(v).requires_grad_(True) for k, v in P.items()}
    oxkv = torch.from_numpy(x_kv).requires_grad_(True)
    ors = [torch.from_numpy(r).requires_grad_(True) for r in res]
    parts = {}
    with torch_ref.operand_rounding(torch.float16 if precision == "fp16" else torch.bfloat16):
        oy = torch_ref.pgrm_forward(Pt, torch.from_numpy(x_q), oxkv, ors, windows=cfg.window_size, num_heads=cfg.num_heads,
                                    drop=dict(seed=seed, drop_rate=0.1, attn_drop_rate=0.1, drop_path=list(m.drop_path)),
                                    parts=parts)
    # The only non-smooth function of the PGRM is the head's LeakyReLU(0.01) (pgrm.py:520): an input within forward
    # rounding noise of 0 may take either slope in two valid evaluations, a 100-fold change of that element's derivative.
    # Each output pixel is exactly one LeakyReLU element (PixelShuffle is a permutation), so the loss simply gives those
    # pixels no weight: both evaluations then differentiate the same function.
    pre = parts["head_pre"].detach()
    near0 = torch.nn.functional.pixel_shuffle((pre.abs() < 4 * fwd_tol * pre.abs().max()).float(), 2) > 0
    assert 0 < int(near0.sum()) < 0.1 * near0.numel()
    G = torch.where(near0, torch.zeros_like(G), G)
    with torch_ref.operand_rounding(torch.float16 if precision == "fp16" else torch.bfloat16):
        (oy * G).sum().backward()
    y = m(xq, xkv, rs)
    (y * G.to(dev)).sum().backward()
    # same arithmetic, same forward -- up to operands that sit on a 16-bit rounding boundary and fall to either side
    # depending on the fp32 summation order of the value being rounded (one 16-bit ulp on isolated elements)
    assert rel_err(y.detach().cpu().numpy(), oy.detach().numpy()) < fwd_tol
    bad, n_checked, worst = [], 0, (0.0, "")
    pairs = [("x_kv", xkv.grad.cpu().numpy(), oxkv.grad.numpy())]
    pairs += [(k, p.grad.cpu().numpy(), Pt[k].grad.numpy()) for k, p in m.named_parameters() if Pt[k].grad is not None]
    for k, got, want in pairs:
        if np.abs(want).max() == 0.0:
            continue
        l2, cos = _l2_cos(got, want)
        n_checked += 1
        worst = max(worst, (l2, k))
        if not (l2 < l2_tol and cos > 1 - l2_tol):
            bad.append((l2, cos, k))
    print(f"pgrm {precision}: forward {rel_err(y.detach().cpu().numpy(), oy.detach().numpy()):.2e}, worst gradient rel L2 {worst}")
    assert n_checked > 60 and not bad, sorted(bad, reverse=True)[:10]


def test_image_loss_and_to_mask_kernels_match_reference():
    """dpmn_image_loss (value + gradient in one pass) and dpmn_to_mask (bit-exact integer work) against
    outputs of the unmodified reference functions."""
    import os
    from dpmn_b200.train import image_loss, to_mask
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neighbors.npz"))
    dev = torch.device("cuda")
    for i in range(3):
        o = torch.from_numpy(z[f"loss{i}_out"]).to(dev).requires_grad_(True)
        t = torch.from_numpy(z[f"loss{i}_tgt"]).to(dev)
        v = image_loss(o, t, tuple(float(x) for x in z[f"loss{i}_w"]))
        (v * 100).backward()
        assert abs(float(v.detach()) - float(z[f"loss{i}_val"])) < 2e-6 * abs(float(z[f"loss{i}_val"]))
        assert rel_err(o.grad.cpu().numpy(), z[f"loss{i}_grad"]) < 1e-5
    # channel-slice view of a 4-channel HR batch (images_hr[:, :3], super_resolution.py:212)
    four = torch.cat([torch.from_numpy(z["loss0_tgt"]), torch.zeros(3, 1, 32, 128)], dim=1).to(dev)
    o = torch.from_numpy(z["loss0_out"]).to(dev)
    assert abs(float(image_loss(o, four[:, :3])) - float(z["loss0_val"])) < 2e-6 * abs(float(z["loss0_val"]))
    got = to_mask(torch.from_numpy(z["mask_in"]).to(dev)).cpu().numpy()
    assert np.array_equal(got, z["mask_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("C,windows,B,shifts,p_drop", [(96, [2, 4, 8], 2, [1, 2, 4], 0.0), (96, [2, 4, 8], 3, [1, 2, 4], 0.2),
                                                        (192, [8], 2, [4], 0.0), (192, [4], 1, [2], 0.3), (96, [4, 8], 2, [0, 3], 0.1),
                                                        (96, [2], 1, [0], 0.0)])
def test_attention_backward_tcgen05_against_autograd_of_the_oracle_core(C, windows, B, shifts, p_drop):
    """attn2_bwd_tc.cu (five tcgen05 contractions per 64-row half: S, dP, dQ, dK, dV; bias-table gradient through shared-memory
    atomics) against torch autograd through the oracle's restatement of pgrm.py:197-268 on the same 16-bit operand values and
    the same attn_drop masks.  The kernel rounds dO, dS and P o M to fp16 operands: bar 4e-3 relative L2 per tensor."""
    from dpmn_b200.pgrm import to_window_major, window_attention_windowed_backward
    from oracle import torch_ref
    dev = torch.device("cuda")
    torch.manual_seed(21)
    H, W, heads = 16, 64, 6
    G = len(windows)
    q = torch.randn(B, H * W, C).half()
    kv = torch.randn(B, H * W, 2 * C).half()
    tabs = [torch.randn((2 * w - 1) ** 2, heads // G) * 0.5 for w in windows]
    d_out = (torch.randn(B, H * W, C) * 0.1).half()
    qf, kvf = q.float().requires_grad_(True), kv.float().requires_grad_(True)
    tf = [t.clone().requires_grad_(True) for t in tabs]
    seed, site = 99, 32
    out = torch_ref.window_attention_core(qf, kvf, tf, windows, shifts, H, W, heads // G,
                                          drop=(p_drop, seed) if p_drop > 0 else None, site=site)
    (out * d_out.float()).sum().backward()
    qw = to_window_major(q.to(dev), (H, W), windows, shifts)
    kw = to_window_major(kv[..., :C].contiguous().to(dev), (H, W), windows, shifts)
    vw = to_window_major(kv[..., C:].contiguous().to(dev), (H, W), windows, shifts)
    dq, dkv, dt = window_attention_windowed_backward(qw, kw, vw, d_out.to(dev), [t.to(dev) for t in tabs], B, (H, W), heads, windows,
                                                     shifts, drop=(p_drop, seed, site) if p_drop > 0 else None)
    for name, got, want in [("dq", dq, qf.grad), ("dkv", dkv, kvf.grad)] + [(f"dtable{g}", dt[g], tf[g].grad) for g in range(G)]:
        l2, cos = _l2_cos(got.cpu().numpy(), want.numpy())
        assert l2 < 4e-3 and cos > 0.9999, (name, l2, cos)

