"""Recogniser-input resizes: TextBase.parse_crnn_data / parse_visionlan_data
(/root/reference/interfaces/base.py:419-425, 473-478).  Fixtures minted from the unmodified reference methods
(tests/golden/resize.npz, oracle/make_golden_resize.py).  Bars: the bicubic + luma path is floating point -> 5e-6 of the
output maximum (summation order); the ToPILImage -> cv2.resize -> ToTensor path is integer work -> bit-exact."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize.npz")
TOL_BICUBIC = 5e-6


def test_oracle_resizes_match_reference_fixtures():
    from oracle.torch_ref import parse_crnn_data, parse_visionlan_data
    z = np.load(GOLD)
    for i in range(3):
        got = parse_crnn_data(z[f"crnn{i}_in"][:, :3])
        want = z[f"crnn{i}_out"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < TOL_BICUBIC * np.abs(want).max(), i
    for name in ("vl32", "vl16", "vlodd"):
        for img, want in zip(z[f"{name}_in"], z[f"{name}_out_u8"]):
            got = parse_visionlan_data(img)
            assert got.shape == (1, 3, 64, 256)
            assert np.array_equal(got[0], want.astype(np.float32) / np.float32(255)), name


@pytest.mark.gpu
def test_cuda_resizes_match_reference_fixtures():
    from dpmn_b200.train import parse_crnn_data, parse_visionlan_data
    z = np.load(GOLD)
    dev = torch.device("cuda")
    for i in range(3):
        x = torch.from_numpy(z[f"crnn{i}_in"]).to(dev)
        got = parse_crnn_data(x[:, :3, :, :]).cpu().numpy()          # channel-slice view, as at the call sites
        want = z[f"crnn{i}_out"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < TOL_BICUBIC * np.abs(want).max(), (i, np.abs(got - want).max())
    for name in ("vl32", "vl16", "vlodd"):
        x = torch.from_numpy(z[f"{name}_in"]).to(dev)
        got = parse_visionlan_data(x).cpu().numpy()                   # the whole batch in one launch
        assert np.array_equal(got, z[f"{name}_out_u8"].astype(np.float32) / np.float32(255)), name
        one = parse_visionlan_data(x[0]).cpu().numpy()                # the reference's per-image signature
        assert one.shape == (1, 3, 64, 256) and np.array_equal(one[0], got[0])
    with pytest.raises(RuntimeError):
        parse_crnn_data(torch.zeros(1, 3, 16, 64))                    # no CPU path


def test_oracle_resizes_against_live_libraries_on_random_shapes():
    """Beyond the committed fixtures: the numpy restatements against the libraries the reference calls, for random source
    sizes (up- and down-scaling, odd sizes; out-of-range values are covered by the committed fixture, whose float -> uint8
    wrap-around does not depend on the host CPU).  Skipped where OpenCV / torchvision are absent."""
    cv2 = pytest.importorskip("cv2")
    from oracle.torch_ref import parse_crnn_data, parse_visionlan_data
    r = np.random.default_rng(77)
    for _ in range(12):
        h, w = int(r.integers(5, 90)), int(r.integers(9, 300))
        img = r.uniform(0, 1, (3, h, w)).astype(np.float32)
        u8 = (torch.from_numpy(img).mul(255).byte().numpy()).transpose(1, 2, 0)          # ToPILImage: mul(255).byte()
        want = cv2.resize(np.ascontiguousarray(u8), (256, 64)).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
        assert np.array_equal(parse_visionlan_data(img)[0], want), (h, w)
    for _ in range(6):
        b, h, w = int(r.integers(1, 4)), int(r.integers(8, 70)), int(r.integers(16, 200))
        x = r.uniform(0, 1, (b, 3, h, w)).astype(np.float32)
        t = torch.nn.functional.interpolate(torch.from_numpy(x), (32, 100), mode="bicubic")
        want = (0.299 * t[:, 0:1] + 0.587 * t[:, 1:2] + 0.114 * t[:, 2:3]).numpy()
        assert np.abs(parse_crnn_data(x) - want).max() < TOL_BICUBIC * np.abs(want).max(), (b, h, w)
