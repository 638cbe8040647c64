"""One configuration of the stand-alone window attention kernel, a few launches (for ncu).  usage: attn_one.py [B] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpmn_b200.pgrm import window_attention_windowed  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
H, W, L, C, heads = 16, 64, 1024, 96, 6
windows, shifts = [2, 4, 8], [1, 2, 4]
dev = torch.device("cuda")
G = len(windows)
tabs = [torch.randn((2 * w - 1) ** 2, heads // G, device=dev) * 0.5 for w in windows]
bufs = [[torch.randn(G, B * L, C // G, device=dev, dtype=torch.float16) for _ in range(3)] for _ in range(4)]
for i in range(n):
    q, k, v = bufs[i % 4]
    window_attention_windowed(q, k, v, tabs, B, (H, W), heads, windows, shifts)
torch.cuda.synchronize()
