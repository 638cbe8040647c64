"""Golden fixtures for the recogniser-input resizes from the UNMODIFIED reference methods
TextBase.parse_crnn_data / parse_visionlan_data (interfaces/base.py:419-425, 473-478).  interfaces/base.py imports modules
that are absent here, so the two methods' SOURCE is taken from the reference file by name and executed in a minimal
namespace (torch.nn, torchvision.transforms, cv2, numpy) with a stand-in `self` that only carries `.device`.
Run in the build container only:   python -m oracle.make_golden_resize"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def reference_methods():
    import cv2
    from torch import nn
    from torchvision import transforms
    tree = ast.parse(open(os.path.join(REF, "interfaces", "base.py")).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TextBase")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("parse_crnn_data", "parse_visionlan_data")]
    assert len(fns) == 2
    ns = dict(nn=nn, transforms=transforms, cv2=cv2, np=np, torch=torch)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "base.py:TextBase", "exec"), ns)
    me = types.SimpleNamespace(device=torch.device("cpu"))
    return (lambda x: ns["parse_crnn_data"](me, x)), (lambda x: ns["parse_visionlan_data"](me, x))


def main():
    crnn, vlan = reference_methods()
    r = np.random.default_rng(321)
    save = {}
    for i, shape in enumerate([(3, 4, 16, 64), (2, 4, 32, 128), (2, 3, 20, 36)]):
        x = r.uniform(0, 1, shape).astype(np.float32)
        save[f"crnn{i}_in"] = x
        save[f"crnn{i}_out"] = crnn(torch.from_numpy(x)[:, :3, :, :]).numpy()           # call sites pass images[:, :3]
    imgs = [r.uniform(0, 1, (3, 32, 128)).astype(np.float32) for _ in range(3)]
    imgs[1] = np.round(imgs[1] * 255) / 255
    imgs[2] = r.uniform(-0.3, 1.4, (3, 32, 128)).astype(np.float32)                      # left [0, 1]: uint8 wrap-around
    small = [r.uniform(0, 1, (3, 16, 64)).astype(np.float32) for _ in range(2)]
    odd = [r.uniform(0, 1, (3, 23, 77)).astype(np.float32)]
    for name, group in (("vl32", imgs), ("vl16", small), ("vlodd", odd)):
        outs = np.concatenate([vlan(torch.from_numpy(im)).numpy() for im in group], axis=0)    # (n, 3, 64, 256)
        u8 = np.rint(outs * 255).astype(np.uint8)
        assert np.array_equal(u8.astype(np.float32) / np.float32(255), outs)                     # stored losslessly as bytes
        save[f"{name}_in"], save[f"{name}_out_u8"] = np.stack(group), u8
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **save)
    print("wrote resize.npz", os.path.getsize(os.path.join(OUT, "resize.npz")), {k: v.shape for k, v in save.items()})


if __name__ == "__main__":
    main()
