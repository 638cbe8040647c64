"""BASELINE configs[4] / SURVEY 8d config 5: stand-alone window attention roofline sweep on one GPU.
window {2,4,8,16} x embed_dim {96,192} (6 heads: head_dim 16 / 32), shift {0, ws//2}, batch 8..256, fp16 storage,
plus the production [2,4,8] window list.  Reports per-launch time (CUDA events, L2 flushed between launches by
rotating over buffers larger than L2), achieved HBM GB/s on the algorithmic bytes 4*L*C*2 B per image and TFLOP/s
on 4*L*C*ws^2 FLOP per image, as fractions of MEASURED_PEAKS.json.   python tools/attn_sweep.py > profiles/...md"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpmn_b200.pgrm import window_attention, window_attention_windowed  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6540.0)
TF = peaks.get("bf16_tflops", 1644.0)
H, W, L = 16, 64, 1024
dev = torch.device("cuda")
print(f"# window attention sweep, fp16 storage, grid {H}x{W}, peaks: HBM {HBM:.0f} GB/s, dense 16-bit {TF:.0f} TFLOP/s (MEASURED_PEAKS.json)")
print("| windows | C | head_dim | shift | B | us/launch | GB/s | % HBM | TFLOP/s | % tensor | kernel |")
print("|---|---|---|---|---|---|---|---|---|---|---|")


def bench(windows, C, shifted, B):
    G = len(windows)
    heads = 6
    d = C // heads
    shifts = [(w // 2 if (shifted and w < min(H, W)) else 0) for w in windows]
    hpg = heads // G
    tabs = [torch.randn((2 * w - 1) ** 2, hpg, device=dev) * 0.5 for w in windows]
    nbuf = max(2, int(300e6 // (B * L * C * 2 * 4)) + 1)      # rotate so that consecutive launches miss L2
    tc = all(w in (2, 4, 8, 16) for w in windows) and d in (16, 32)
    if tc:      # the production kernel: window-major operands (content is irrelevant for timing)
        qs = [torch.randn(G, B * L, C // G, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        ks = [torch.randn(G, B * L, C // G, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        vs = [torch.randn(G, B * L, C // G, device=dev, dtype=torch.float16) for _ in range(nbuf)]

        def run(i):
            window_attention_windowed(qs[i % nbuf], ks[i % nbuf], vs[i % nbuf], tabs, B, (H, W), heads, windows, shifts)
    else:
        qs = [torch.randn(B, L, C, device=dev, dtype=torch.float16) for _ in range(nbuf)]
        kvs = [torch.randn(B, L, 2 * C, device=dev, dtype=torch.float16) for _ in range(nbuf)]

        def run(i):
            window_attention(qs[i % nbuf], kvs[i % nbuf], tabs, (H, W), heads, windows, shifts)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    # the launches are replayed from a CUDA graph: enqueueing one from Python (ctypes + three tensor-map encodes) costs
    # ~23 us, more than the kernel itself at small batch
    n = 20
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(n):
                run(i)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (n * reps)
    bytes_ = 4 * L * C * 2 * B
    flops = 4 * L * (C // G) * sum(w * w for w in windows) * B
    gbs, tfl = bytes_ / us / 1e3, flops / us / 1e6
    print(f"| {windows} | {C} | {d} | {shifts} | {B} | {us:.1f} | {gbs:.0f} | {100 * gbs / HBM:.1f} | {tfl:.2f} | {100 * tfl / TF:.2f} | "
          f"{('attn_tc v1' if os.environ.get('DPMN_ATTN_V1') == '1' else 'attn2_tc') + ' (tcgen05)' if tc else 'window_attn_simt'} |", flush=True)


QUICK = "--quick" in sys.argv
for C in (() if QUICK else (96, 192)):
    for ws in (2, 4, 8, 16):
        for shifted in (False, True):
            if ws == 16 and shifted:
                continue     # window == min(H, W): the reference forces shift 0 (pgrm.py:148-150)
            for B in (8, 32, 128, 256):
                bench([ws], C, shifted, B)
for B in (8, 16, 32, 48, 64, 128, 256):
    bench([2, 4, 8], 96, True, B)
