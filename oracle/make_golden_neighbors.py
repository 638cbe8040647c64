"""Golden fixtures for the two 'next' rows from the UNMODIFIED reference functions:
ImageLoss (loss/image_loss.py) value + autograd gradient, and toMask (utils/util.py:27-35).
Run in the build container only:   python -m oracle.make_golden_neighbors"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def to_mask_reference():
    """utils/util.py imports modules that are absent here; toMask itself only needs torchvision + numpy + PIL, so the
    function's SOURCE is executed from the reference file (lines 27-35 located by name) in a minimal namespace."""
    import ast
    src = open(os.path.join(REF, "utils", "util.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "toMask")
    ns = {}
    from torchvision import transforms
    ns.update(transforms=transforms, np=np, torch=torch)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "util.py:toMask", "exec"), ns)
    return ns["toMask"]


def main():
    loss_mod = _load(os.path.join(REF, "loss", "image_loss.py"), "ref_image_loss")
    r = np.random.default_rng(123)
    save = {}
    for i, (B, C, H, W, w) in enumerate([(3, 3, 32, 128, (1.0, 1.0)), (2, 4, 16, 24, (20.0, 1e-4)), (1, 3, 32, 128, (1.0, 1.0))]):
        out = r.uniform(0, 1, (B, C, H, W)).astype(np.float32)
        tgt = r.uniform(0, 1, (B, C, H, W)).astype(np.float32)
        if i == 2:
            out[:, :, 5:9, 10:40] = tgt[:, :, 5:9, 10:40]          # flat agreement region: sign(0) terms
        crit = loss_mod.ImageLoss(gradient=True, loss_weight=list(w))
        o = torch.from_numpy(out).requires_grad_(True)
        val = crit(o, torch.from_numpy(tgt))
        (val * 100).backward()
        save[f"loss{i}_out"], save[f"loss{i}_tgt"] = out, tgt
        save[f"loss{i}_w"] = np.asarray(w, np.float32)
        save[f"loss{i}_val"] = np.asarray(float(val), np.float64)
        save[f"loss{i}_grad"] = o.grad.numpy()          # gradient of 100 * loss
    to_mask = to_mask_reference()
    imgs = r.uniform(0, 1, (6, 3, 32, 128)).astype(np.float32)
    imgs[1] = np.round(imgs[1] * 255) / 255            # exactly representable bytes
    imgs[2, :, :, :64] *= 0.2                           # bimodal
    imgs[3] = 0.5                                       # constant image: nothing is above the mean
    imgs[4] = r.uniform(-0.3, 1.4, (3, 32, 128)).astype(np.float32)   # a cascade image that left [0, 1]: uint8 wrap-around
    masks = np.concatenate([to_mask(torch.from_numpy(im)).numpy() for im in imgs], axis=0)
    save["mask_in"], save["mask_out"] = imgs, masks.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "neighbors.npz"), **save)
    print("wrote neighbors.npz", {k: v.shape for k, v in save.items()})


if __name__ == "__main__":
    main()
