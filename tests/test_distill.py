"""DistillModule (/root/reference/model/distill_module.py): the oracle restatement against fixtures
minted from the unmodified reference class (tests/golden/distill.npz, oracle/make_golden_distill.py) on the CPU, and the
CUDA module (dpmn_distill_forward / dpmn_distill_backward) against the same fixtures on the GPU.
Bars: loss 2e-6 rel, feature 1e-5 of max, gradients 2e-4 of max (L1 sign / ReLU masks make them piecewise constant:
one element within rounding distance of a kink moves a parameter gradient by ~1/(3*B*H*W))."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "distill.npz")
CASES = ["train", "eval", "train_small", "train_feat_only"]
PARAMS = ["conv_cat_feature.weight", "conv_cat_feature.bias", "bn_1.weight", "bn_1.bias",
          "conv_feature.weight", "conv_feature.bias", "bn_2.weight", "bn_2.bias"]


def _case(z, name):
    B, H, W, training, feat_grad, loss_grad = (int(v) for v in z[f"{name}/meta"])
    P = {k.split("p:", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{name}/p:")}
    return B, H, W, bool(training), bool(feat_grad), bool(loss_grad), P


def _rel(got, want, floor=0.0):
    scale = max(float(np.abs(want).max()), floor)
    return float(np.abs(np.asarray(got, np.float64) - want).max()) / scale


def _check_grads(z, name, grads, training, tol):
    # conv biases feeding a train-mode BatchNorm have an identically-zero gradient (the reference holds rounding noise
    # ~1e-10 there): compare those on the scale of the BatchNorm bias gradients instead
    floor = max(float(np.abs(z[f"{name}/g:bn_1.bias"]).max()), float(np.abs(z[f"{name}/g:bn_2.bias"]).max()), 1e-12)
    for k, got in grads.items():
        want = z[f"{name}/g:{k}"]
        assert got.shape == want.shape, (k, got.shape, want.shape)
        fl = floor if (training and k in ("conv_cat_feature.bias", "conv_feature.bias")) else 1e-30
        if float(np.abs(want).max()) == 0.0:
            assert float(np.abs(got).max()) <= tol * floor, (name, k)
            continue
        assert _rel(got, want, fl) < tol, (name, k, _rel(got, want, fl))


@pytest.mark.parametrize("name", CASES)
def test_oracle_distill_matches_reference_fixtures(name):
    from oracle.torch_ref import distill_forward, distill_running_update
    z = np.load(GOLD)
    B, H, W, training, feat_grad, loss_grad, P = _case(z, name)
    P = {k: (v.clone().requires_grad_(True) if k in PARAMS else v) for k, v in P.items()}
    xd = torch.from_numpy(z[f"{name}/x_deep"]).requires_grad_(True)
    xs = torch.from_numpy(z[f"{name}/x_shallow"]).requires_grad_(True)
    loss, feat = distill_forward(P, xd, xs, training=training)
    assert abs(float(loss.detach()) - float(z[f"{name}/loss"])) < 2e-6 * abs(float(z[f"{name}/loss"]))
    assert _rel(feat.detach().numpy(), z[f"{name}/feature"]) < 1e-5
    total = (loss * 100 if loss_grad else 0) + ((feat * torch.from_numpy(z[f"{name}/G"])).sum() if feat_grad else 0)
    total.backward()
    grads = {k: (P[k].grad.numpy() if P[k].grad is not None else np.zeros(tuple(P[k].shape), np.float32)) for k in PARAMS}
    grads["x_deep"], grads["x_shallow"] = xd.grad.numpy(), xs.grad.numpy()
    _check_grads(z, name, grads, training, 2e-4)
    if training:
        with torch.no_grad():
            after = distill_running_update({k: v.detach() for k, v in P.items()}, xd.detach(), xs.detach())
        for k, v in after.items():
            assert _rel(v.numpy(), z[f"{name}/after:{k}"]) < 1e-5, k


def test_distill_state_dict_schema_matches_reference():
    """Same keys, shapes and dtypes as the reference module's state_dict (the `p:` entries of the fixture are its dump)."""
    from dpmn_b200.distill import DistillModule
    z = np.load(GOLD)
    want = {k.split("p:", 1)[1]: z[k] for k in z.files if k.startswith("train/p:")}
    sd = DistillModule().state_dict()
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k].shape), k
        assert (v.dtype == torch.int64) == (want[k].dtype == np.int64), k
    with pytest.raises(RuntimeError):          # no CPU path
        DistillModule()(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_distill_cuda_matches_reference_fixtures(name):
    from dpmn_b200.distill import DistillModule
    z = np.load(GOLD)
    B, H, W, training, feat_grad, loss_grad, P = _case(z, name)
    dev = torch.device("cuda")
    m = DistillModule().to(dev)
    m.load_state_dict(P, strict=True)
    m.train(training)
    xd = torch.from_numpy(z[f"{name}/x_deep"]).to(dev).requires_grad_(True)
    xs = torch.from_numpy(z[f"{name}/x_shallow"]).to(dev).requires_grad_(True)
    loss, feat = m(xd, xs)
    assert loss.shape == () and feat.shape == (B, 3, H, W)
    assert abs(float(loss.detach()) - float(z[f"{name}/loss"])) < 2e-6 * abs(float(z[f"{name}/loss"]))
    assert _rel(feat.detach().cpu().numpy(), z[f"{name}/feature"]) < 1e-5
    total = (loss * 100 if loss_grad else 0) + ((feat * torch.from_numpy(z[f"{name}/G"]).to(dev)).sum() if feat_grad else 0)
    total.backward()
    grads = {k: (p.grad.cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
             for k, p in m.named_parameters()}
    grads["x_deep"], grads["x_shallow"] = xd.grad.cpu().numpy(), xs.grad.cpu().numpy()
    _check_grads(z, name, grads, training, 2e-4)
    sd = m.state_dict()
    for k in sd:
        if "running" in k or "num_batches" in k:
            assert _rel(sd[k].cpu().numpy(), z[f"{name}/after:{k}"], 1e-30) < 1e-5, k


@pytest.mark.gpu
def test_distill_cuda_strided_inputs_no_grad_and_stateless_backward():
    """Channel-slice views of 4-channel tensors (batch stride 4*H*W) are read in place; under no_grad the module returns
    the same values; the C-ABI backward without DPMN_DISTILL_WORKSPACE_HOLDS_FORWARD recomputes the forward itself."""
    import ctypes as C
    from dpmn_b200 import _lib
    from dpmn_b200.distill import DistillModule
    z = np.load(GOLD)
    name = "train_small"
    B, H, W, training, _, _, P = _case(z, name)
    dev = torch.device("cuda")
    m = DistillModule().to(dev)
    m.load_state_dict(P, strict=True)
    m.train(True)
    four_d = torch.cat([torch.from_numpy(z[f"{name}/x_deep"]), torch.zeros(B, 1, H, W)], dim=1).to(dev)
    four_s = torch.cat([torch.from_numpy(z[f"{name}/x_shallow"]), torch.zeros(B, 1, H, W)], dim=1).to(dev)
    with torch.no_grad():
        loss, feat = m(four_d[:, :3], four_s[:, :3])
    assert abs(float(loss) - float(z[f"{name}/loss"])) < 2e-6 * abs(float(z[f"{name}/loss"]))
    assert _rel(feat.cpu().numpy(), z[f"{name}/feature"]) < 1e-5
    # stateless backward straight through the C ABI
    lib = _lib.load()
    xd, xs = four_d[:, :3], four_s[:, :3]
    d = m._descriptor(B, H, W, True, False, 4 * H * W, 4 * H * W)
    ws = torch.empty(int(lib.dpmn_distill_workspace_bytes(C.byref(d))), dtype=torch.uint8, device=dev)
    g = _lib.DistillGrads()
    bufs = {k: torch.zeros_like(p) for k, p in m.named_parameters()}
    g.conv_cat_w, g.conv_cat_b = bufs["conv_cat_feature.weight"].data_ptr(), bufs["conv_cat_feature.bias"].data_ptr()
    g.conv_w, g.conv_b = bufs["conv_feature.weight"].data_ptr(), bufs["conv_feature.bias"].data_ptr()
    g.bn_1.w, g.bn_1.b = bufs["bn_1.weight"].data_ptr(), bufs["bn_1.bias"].data_ptr()
    g.bn_2.w, g.bn_2.b = bufs["bn_2.weight"].data_ptr(), bufs["bn_2.bias"].data_ptr()
    gxd, gxs = torch.empty(B, 3, H, W, device=dev), torch.empty(B, 3, H, W, device=dev)
    g.x_deep, g.x_shallow = gxd.data_ptr(), gxs.data_ptr()
    gl = torch.full((), 100.0, device=dev)
    rc = lib.dpmn_distill_backward(C.byref(d), xd.data_ptr(), xs.data_ptr(), gl.data_ptr(), None, C.byref(g), ws.data_ptr(),
                                   ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    grads = {k: v.cpu().numpy() for k, v in bufs.items()}
    grads["x_deep"], grads["x_shallow"] = gxd.cpu().numpy(), gxs.cpu().numpy()
    _check_grads(z, name, grads, True, 2e-4)
    assert lib.dpmn_distill_backward(C.byref(d), xd.data_ptr(), xs.data_ptr(), gl.data_ptr(), None, C.byref(g), ws.data_ptr(),
                                     16, torch.cuda.current_stream().cuda_stream) == -3
