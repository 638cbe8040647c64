"""Host-side training logic that needs no GPU: the image loss restatement against the reference's definition."""
import numpy as np
import torch


def test_image_loss_matches_reference_definition():
    """ImageLoss(gradient=True, loss_weight=[1,1]) = MSE + L1(gradient_map) (loss/image_loss.py:15-43), written out
    independently here with explicit shifts."""
    from oracle.torch_ref import gradient_map, image_loss
    r = np.random.default_rng(0)
    a = torch.from_numpy(r.uniform(0, 1, (2, 3, 8, 12)).astype(np.float32))
    b = torch.from_numpy(r.uniform(0, 1, (2, 3, 8, 12)).astype(np.float32))

    def gmap(x):
        x = x.numpy().astype(np.float64)
        rr = np.zeros_like(x); rr[..., :-1] = x[..., 1:]
        ll = np.zeros_like(x); ll[..., 1:] = x[..., :-1]
        tt = np.zeros_like(x); tt[..., 1:, :] = x[..., :-1, :]
        bb = np.zeros_like(x); bb[..., :-1, :] = x[..., 1:, :]
        return np.sqrt(((rr - ll) * 0.5) ** 2 + ((tt - bb) * 0.5) ** 2 + 1e-6)
    assert np.allclose(gradient_map(a).numpy(), gmap(a), atol=1e-6)
    want = ((a.numpy().astype(np.float64) - b.numpy()) ** 2).mean() + np.abs(gmap(a) - gmap(b)).mean()
    assert abs(float(image_loss(a, b)) - want) < 1e-6


def test_numpy_mask_hash_matches_the_library():
    """oracle/torch_ref.mask_hash restates dpmn_mask_hash (the host-callable hash behind the train-mode Dropout /
    DropPath masks); no GPU needed."""
    from dpmn_b200 import _lib
    from oracle.torch_ref import mask_hash
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for _ in range(200):
        seed = int(rng.integers(0, 2 ** 62))
        site = int(rng.integers(0, 64))
        idx = int(rng.integers(0, 2 ** 40))
        assert int(mask_hash(seed, site, np.array([idx]))[0]) == lib.dpmn_mask_hash(seed, site, idx)
    # keep fraction of a Bernoulli(0.1) drop
    h = mask_hash(7, 17, np.arange(200000))
    u = (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    assert abs(float((u >= 0.1).mean()) - 0.9) < 5e-3


def test_oracle_image_loss_and_to_mask_match_reference_fixtures():
    """oracle/torch_ref.image_loss / to_mask against outputs of the unmodified reference functions
    (tests/golden/neighbors.npz, minted by oracle/make_golden_neighbors.py)."""
    import os
    from oracle.torch_ref import image_loss, to_mask
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "neighbors.npz"))
    for i in range(3):
        o = torch.from_numpy(z[f"loss{i}_out"]).requires_grad_(True)
        v = image_loss(o, torch.from_numpy(z[f"loss{i}_tgt"]), tuple(float(x) for x in z[f"loss{i}_w"]))
        (v * 100).backward()
        assert abs(float(v.detach()) - float(z[f"loss{i}_val"])) < 1e-6 * abs(float(z[f"loss{i}_val"]))
        assert np.abs(o.grad.numpy() - z[f"loss{i}_grad"]).max() < 1e-6 * np.abs(z[f"loss{i}_grad"]).max()
    for img, want in zip(z["mask_in"], z["mask_out"]):
        assert np.array_equal(to_mask(img), want)


def test_distill_chain_follows_the_reference_loop():
    """HotPathTrainer.distill_loss walks the cascade outputs exactly like interfaces/super_resolution.py:245-263: per branch
    from the deepest image back, module distill_list[k-1] (branch 1) / distill_list[k+b1-2] (branch 2), x_deep = the feature
    handed on by the previous module, x_shallow = cascade[k-1].  Host logic only: stand-in modules record their calls."""
    import types
    from dpmn_b200.train import HotPathTrainer

    for b1, b2 in ((3, 3), (2, 4)):
        calls = []

        def make(idx):
            def module(x_deep, x_shallow):
                calls.append((idx, x_deep, x_shallow))
                return torch.tensor(float(idx + 1)), f"feat{idx}"
            return module
        outs = [f"b1_{k}" for k in range(b1)] + [f"b2_{k}" for k in range(b2)] + ["cmm"]
        fake = types.SimpleNamespace(model=types.SimpleNamespace(b1=b1, b2=b2), distill=[make(i) for i in range(b1 + b2 - 2)])
        # `outs[0].new_zeros(())` is the only tensor method the loop needs from the images
        class Img(str):
            def new_zeros(self, shape):
                return torch.zeros(shape)
        outs = [Img(o) for o in outs]
        total = HotPathTrainer.distill_loss(fake, outs)
        # the reference's loop, written out
        want, expect_total = [], 0.0
        feature = outs[b1 - 1]
        for k in range(b1 - 1, 0, -1):
            want.append((k - 1, feature, outs[k - 1]))
            feature = f"feat{k - 1}"
            expect_total += (k - 1 + 1) * 100
        feature = outs[b1 + b2 - 1]
        for k in range(b2 - 1, 0, -1):
            want.append((k + b1 - 2, feature, outs[b1 + k - 1]))
            feature = f"feat{k + b1 - 2}"
            expect_total += (k + b1 - 2 + 1) * 100
        assert [(i, str(d), str(s)) for i, d, s in calls] == [(i, str(d), str(s)) for i, d, s in want]
        assert abs(float(total) - expect_total) < 1e-6


def test_bench_algorithmic_byte_model_matches_design():
    """bench.py's BYTES_IMG (the roofline's algorithmic bytes) against the per-block figure DESIGN.md section 5 states."""
    import bench
    assert bench.BYTES_IMG["gemm_tc"] == 12 * 6_684_672
    assert bench.BYTES_IMG["window_attn_tc"] == 12 * 786_432          # 4 * L * C * 2 B per image-block
    assert bench.FLOPS_IMG["total"] == 11_267_776_512                 # DPMN hot path forward per image
    assert abs(bench.FLOPS_IMG["gemm_tc"] / bench.BYTES_IMG["gemm_tc"] - 80.0) < 0.5


def test_oracle_to_mask_against_live_pillow_on_random_images():
    """oracle/torch_ref.to_mask against the library calls toMask makes (utils/util.py:27-35: ToPILImage, convert('L'),
    threshold at the mean, invert, ToTensor, repeat) on fresh random images of several sizes.  Skipped without torchvision."""
    import pytest
    tv = pytest.importorskip("torchvision")
    from torchvision import transforms
    from oracle.torch_ref import to_mask
    r = np.random.default_rng(5)
    for _ in range(10):
        h, w = int(r.integers(4, 48)), int(r.integers(4, 160))
        img = r.uniform(0, 1, (3, h, w)).astype(np.float32)
        if _ % 3 == 0:
            img[:, :, : w // 2] *= 0.3                         # bimodal: both sides of the mean populated unevenly
        grey = transforms.ToPILImage()(torch.from_numpy(img)).convert("L")
        thres = np.array(grey).mean()
        want = transforms.ToTensor()(grey.point(lambda v: 0 if v > thres else 255)).repeat(3, 1, 1).numpy()
        assert np.array_equal(to_mask(img), want), (h, w)


def test_roofline_traffic_is_read_from_the_committed_ncu_summaries():
    import bench
    t = bench.ncu_traffic_per_launch("gemm_tc")
    assert t is not None and 5e6 < t < 60e6            # mean DRAM bytes per launch of the PGRM GEMM class (ncu --set full)
    ta = bench.ncu_traffic_per_launch("window_attn_tc")
    assert ta is not None and 20e6 < ta < 40e6         # attention: ~28 MB of DRAM traffic per launch at batch 48
    assert bench.ncu_traffic_per_launch("no_such_class") is None


def test_attention_roofline_constants_match_survey_8d():
    """north_star's "fraction of the attention-FLOP roofline" (bench roofline.attention_roofline) rests on SURVEY 8d's figures:
    4 L (C/G) sum ws^2 = 11 010 048 FLOP per image per block, 12 blocks per forward = 132 120 576; stand-alone attention
    traffic 4 L C s = 786 432 B per image per block in 16 bits."""
    import bench
    per_block = 4 * 1024 * 32 * (4 + 16 + 64)
    assert per_block == 11_010_048
    assert bench.FLOPS_IMG["window_attn_tc"] == 12 * per_block == 132_120_576
    assert bench.MIN_BYTES_IMG["window_attn_tc"] == 12 * 786_432
    assert bench.FLOPS_IMG["total"] == 11_267_776_512          # 6 x 1 134.78 + 4 459.07 MFLOP per image
