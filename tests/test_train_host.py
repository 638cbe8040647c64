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
