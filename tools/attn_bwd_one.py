"""One configuration of the stand-alone tcgen05 attention backward, a few launches (for ncu / timing).
usage: attn_bwd_one.py [B] [n] [p_drop]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpmn_b200.pgrm import window_attention_windowed_backward  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
p_drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
H, W, L, C, heads = 16, 64, 1024, 96, 6
windows, shifts = [2, 4, 8], [1, 2, 4]
dev = torch.device("cuda")
G = len(windows)
tabs = [torch.randn((2 * w - 1) ** 2, heads // G, device=dev) * 0.5 for w in windows]
bufs = [[torch.randn(G, B * L, C // G, device=dev, dtype=torch.float16) for _ in range(3)] + [torch.randn(B, L, C, device=dev, dtype=torch.float16) * 0.1]
        for _ in range(4)]
drop = (p_drop, 7, 16) if p_drop > 0 else None
for i in range(3):
    window_attention_windowed_backward(*bufs[i % 4], tabs, B, (H, W), heads, windows, shifts, drop=drop)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    window_attention_windowed_backward(*bufs[i % 4], tabs, B, (H, W), heads, windows, shifts, drop=drop)
e1.record()
torch.cuda.synchronize()
print(f"B={B} p_drop={p_drop}: {e0.elapsed_time(e1) * 1e3 / n:.1f} us per call (incl. host launch overhead + output allocation)")
