"""Golden fixtures for DistillModule from the UNMODIFIED reference class
(/root/reference/model/distill_module.py): forward (loss, feature), BatchNorm buffers after a train-mode forward, and the
reference's own autograd gradients of   100 * loss + sum(feature * G)   w.r.t. every parameter and both inputs
(the feature gradient is what the next DistillModule of the chain sends back, interfaces/super_resolution.py:245-263).
Run in the build container only:   python -m oracle.make_golden_distill"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

CASES = [  # name, B, H, W, training, feature gradient, loss gradient
    ("train", 3, 32, 128, True, True, True),
    ("eval", 2, 32, 128, False, True, True),
    ("train_small", 2, 8, 12, True, False, True),
    ("train_feat_only", 2, 16, 24, True, True, False),
]


def main():
    spec = importlib.util.spec_from_file_location("ref_distill", os.path.join(REF, "model", "distill_module.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    save = {}
    for ci, (name, B, H, W, training, feat_grad, loss_grad) in enumerate(CASES):
        torch.manual_seed(100 + ci)
        r = np.random.default_rng(500 + ci)
        m = mod.DistillModule()
        with torch.no_grad():      # non-trivial affine parameters and running statistics
            for bn in (m.bn_1, m.bn_2):
                bn.weight.copy_(torch.from_numpy(r.uniform(0.5, 1.5, 3).astype(np.float32)))
                bn.bias.copy_(torch.from_numpy(r.uniform(-0.3, 0.3, 3).astype(np.float32)))
                bn.running_mean.copy_(torch.from_numpy(r.uniform(-0.2, 0.2, 3).astype(np.float32)))
                bn.running_var.copy_(torch.from_numpy(r.uniform(0.05, 0.3, 3).astype(np.float32)))
        m.train(training)
        for k, v in m.state_dict().items():
            save[f"{name}/p:{k}"] = v.detach().numpy().copy()
        xd = r.uniform(0, 1, (B, 3, H, W)).astype(np.float32)
        xs = r.uniform(0, 1, (B, 3, H, W)).astype(np.float32)
        G = r.normal(0, 1, (B, 3, H, W)).astype(np.float32) / (B * H * W)
        td, ts = torch.from_numpy(xd).requires_grad_(True), torch.from_numpy(xs).requires_grad_(True)
        loss, feat = m(td, ts)
        total = 0
        if loss_grad:
            total = total + loss.sum() * 100
        if feat_grad:
            total = total + (feat * torch.from_numpy(G)).sum()
        total.backward()
        save[f"{name}/x_deep"], save[f"{name}/x_shallow"], save[f"{name}/G"] = xd, xs, G
        save[f"{name}/meta"] = np.asarray([B, H, W, int(training), int(feat_grad), int(loss_grad)], np.int64)
        save[f"{name}/loss"] = np.asarray(float(loss.detach()), np.float64)
        save[f"{name}/feature"] = feat.detach().numpy()
        save[f"{name}/g:x_deep"], save[f"{name}/g:x_shallow"] = td.grad.numpy(), ts.grad.numpy()
        for k, p in m.named_parameters():
            save[f"{name}/g:{k}"] = p.grad.numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        for k, v in m.state_dict().items():
            if "running" in k or "num_batches" in k:
                save[f"{name}/after:{k}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "distill.npz"), **save)
    print("wrote distill.npz:", len(save), "arrays,", os.path.getsize(os.path.join(OUT, "distill.npz")), "bytes")


if __name__ == "__main__":
    main()
