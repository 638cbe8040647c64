#!/usr/bin/env python
"""bench.py -- SR images/sec of the DPMN hot path (6 x PGRM + CMM, batch 48/GPU) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|fp16|bf16] [--impl ours|reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definitions of every key.
  value          inference images/s with inputs resident in HBM (CUDA events, max over ranks)     [BASELINE configs[1]]
  e2e            the same through the public module API from pinned HOST buffers, H2D + D2H inside the timed region
  train          BASELINE configs[2] in the same line: HotPathTrainer.step at batch 48/GPU -- forward, 7 image losses +
                 4 distill terms, backward, the flat NCCL all-reduce of the 228.7 MB gradient bucket (N > 1), clip + Adam;
                 images/s, ms/step, all-reduce ms (on the communication stream / exposed on the compute stream), by_kernel_ms
  roofline       SURVEY 8d accounting: the dominant kernel class against the TENSOR roofline (the fused block is
                 tensor-bound, AI ~ 940 FLOP/B); its DRAM traffic against 8d's minimum bytes; a per-class table with both
                 fractions (HBM GB/s and TFLOP/s); kernels timed per launch with CUDA events by the library's profile hook
  parity         the output of the timed configuration diffed against the CPU port on the same inputs (8d)
  cpu_baseline   the torch-CPU port of the reference (oracle/torch_ref.py) on this host's cores, FULL batch-48 forwards
--impl reference times that same CPU port as the reference arm (the reference is pure PyTorch and /root/reference does
not exist on the GPU box; oracle/torch_ref.py is pinned to its outputs), on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 48
ALPHA = 0.5          # --alpha of every DPMN run (README.md:42): image_sr = alpha * CMM + (1 - alpha) * PSN image
METRIC = "SR images/sec"
WORKLOAD = "DPMN hot path forward: 6xPGRM cascade (b1=b2=3, win 2/4/8, dim 96, 6 heads) + CMM + alpha blend, 16x64 -> 32x128, " \
           "batch 48/GPU, synthetic PSN output + priors (configs[1] without the frozen TATT backbone)"

# algorithmic FLOPs per image (2*MAC, matmul/conv only; SURVEY.md 8d / BASELINE.md section 3)
_PGRM_GEMM = 6 * 2 * (56_623_104 + 25_165_824 + 460_062_720 - 7_077_888)   # q/kv + SK proj + Mlp GEMMs (dw conv excluded)
_FC1 = 6 * 2 * 75_497_472                                                  # fc1: 2 * 1024 * 96 * 384 per block
FLOPS_IMG = {"gemm": _PGRM_GEMM, "gemm_tc": _PGRM_GEMM,
             "mlp_fc1_dw_tc": _FC1 + 6 * 2 * 7_077_888,                     # fc1 + depthwise 3x3 (fused kernel A)
             "window_attn": 6 * 2 * 11_010_048, "window_attn_tc": 6 * 2 * 11_010_048,
             "conv": 4_459_069_440, "conv_tc": 4_459_069_440 - 14_155_776 * 2 - 42_467_328,   # minus stem / de_1 (own kernels)
             "total": 11_267_776_512}

_L, _C, _HID = 1024, 96, 384
# SURVEY 8d minimum HBM bytes per image: a fused block reads x_q, x_kv and writes x_kv = 3 L C s (16-bit, s = 2); the
# stand-alone attention reads q, k, v and writes out = 4 L C s; the CMM reads two images and writes one (fp32)
MIN_BYTES_IMG = {"block": 3 * _L * _C * 2, "pgrm_blocks": 12 * 3 * _L * _C * 2, "window_attn_tc": 12 * 4 * _L * _C * 2,
                 "cmm": 3 * 3 * 32 * 128 * 4}
# bytes the UNFUSED launch sequence moves per image (every intermediate in and out once): kept so that
# the distance between what is launched today and 8d's minimum stays visible
BYTES_IMG = {
    "gemm_tc": 12 * (10 * _L * _C * 2 + 4 * _L * _C * 4 + 4 * _L * _HID * 2),
    "window_attn_tc": 12 * 4 * _L * _C * 2,
    "dwconv": 12 * 2 * _L * _HID * 2,
    "mlp_fc1_dw_tc": 12 * (_L * _C * 2 + _L * _HID * 2),
}


def ncu_traffic_per_launch(kernel_class):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel class, read from the committed summaries of
    the `ncu --set full` captures (profiles/*.txt, written by tools/ncu_summary.py); None when the class was not captured."""
    files = {"gemm_tc": (("r02_ncu_full_block_v5.txt", "r02_ncu_full_block_v4.txt", "r02_ncu_full_block.txt", "r01_ncu_full_fp16_v14.txt"), ("gemm_tc_kernel", "gemm_res_ln_kernel")),
             "conv_tc": (("r01_ncu_full_conv_tc_v21.txt",), ("conv_tc_kernel<64>", "conv_tc_kernel<128>")),
             "dwconv": (("r01_ncu_full_dwconv_v2_en1_v23.txt",), ("dwconv16_v2_kernel",)),
             "mlp_fc1_dw_tc": (("r02_ncu_full_block_v5.txt", "r02_ncu_full_block_v4.txt", "r02_ncu_full_block.txt"), ("mlp_fc1_dw_kernel",)),
             "window_attn_tc": (("r02_ncu_full_block_v5.txt", "r02_ncu_full_block_v4.txt", "r02_ncu_full_block.txt", "r01_ncu_full_fp16_v14.txt"), ("attn2_tc_kernel", "attn_tc_kernel"))}
    if kernel_class not in files:
        return None
    names, kernels = files[kernel_class]
    for name in names:
        try:
            rows = [l.split() for l in open(os.path.join(ROOT, "profiles", name)) if not l.startswith("#")]
            vals = []
            for r in rows[1:]:
                line = " ".join(r)
                if any(k in line for k in kernels):
                    nums = [x for x in r if x.replace(".", "", 1).isdigit()]
                    # numeric columns: id, time_us, dram_rd_MB, dram_wr_MB, ...
                    if len(nums) >= 4:
                        vals.append((float(nums[2]) + float(nums[3])) * 1e6)
            if vals:
                return float(np.mean(vals))
        except Exception:           # evidence lookup must never take the measurement down
            continue
    return None


def synth_inputs(seed, B):
    from dpmn_b200 import synth as gen
    r = np.random.default_rng([seed, 99])
    psn = r.uniform(0, 1, size=(B, 4, 32, 128)).astype(np.float32)
    psn[:, 3] = (psn[:, 3] > 0.5)
    p1 = [gen.prior_branch1(seed + k, B) for k in range(3)]
    p2 = [gen.prior_branch2(seed + k, B) for k in range(3)]
    return psn, p1, p2


def synth_weights(seed):
    """Deterministic non-trivial weights for the 6 PGRMs + CMM (random-init architecture, BASELINE `data`)."""
    from dpmn_b200.schema import PGRMConfig, cmm_schema, pgrm_schema
    from dpmn_b200.synth import synth_params
    pg = [synth_params(pgrm_schema(PGRMConfig(iter=k, mode=(k >= 3))), seed + k) for k in range(6)]
    cm = synth_params(cmm_schema(3, 64), seed + 50)
    return pg, cm


def load_weights(model, pg, cm):
    for k, m in enumerate(model.pgrm):
        sd = m.state_dict()
        m.load_state_dict({n: (torch.from_numpy(pg[k][n]) if n in pg[k] else v) for n, v in sd.items()}, strict=True)
    model.cmm.load_state_dict({n: torch.from_numpy(np.asarray(cm[n])) for n in model.cmm.state_dict()}, strict=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_throughput(B, budget_s=25.0, steps=None, warmup=1, keep_output=False):
    """Time oracle/torch_ref.hot_path_forward (with the alpha blend) on the host cores at the FULL per-GPU batch; returns
    (images/s, description, ms/step, output of the last forward | None).  Inputs: synth_inputs(0, B) = rank 0's input set 0."""
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    pg, cm = synth_weights(2)
    pg = [{k: torch.from_numpy(np.asarray(v)) for k, v in p.items()} for p in pg]
    cm = {k: torch.from_numpy(np.asarray(v)) for k, v in cm.items()}
    psn, p1, p2 = synth_inputs(0, B)
    args = (torch.from_numpy(psn), [torch.from_numpy(a) for a in p1], [torch.from_numpy(a) for a in p2])
    times, out = [], None
    with torch.no_grad():
        for _ in range(warmup):
            out = torch_ref.hot_path_forward(pg, cm, *args, alpha=ALPHA)
        t_end = time.perf_counter() + budget_s
        n = 0
        while (steps is None and time.perf_counter() < t_end and n < 50) or (steps is not None and n < steps):
            t0 = time.perf_counter()
            out = torch_ref.hot_path_forward(pg, cm, *args, alpha=ALPHA)
            times.append(time.perf_counter() - t0)
            n += 1
    med = float(np.median(times))
    desc = f"{len(times)} timed full forwards of the batch-{B} workload (median; {warmup} warm-up), torch {torch.__version__} " \
           f"CPU fp32, {torch.get_num_threads()} threads"
    return B / med, desc, med * 1e3, (out if keep_output else None)


def infer_config(B, world):
    return {"workload": WORKLOAD.replace("batch 48/GPU", f"batch {B}/GPU"), "global_batch": B * world,
            "parallelism": f"replicas x{world} (inference: no collective; the `train` object carries the dp{world} all-reduce)",
            "l2": "inputs rotate over 4 sets; per-step working set (activations/workspace) >> 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    B = args.batch
    val, sample, ms, _ = cpu_port_throughput(B, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": infer_config(B, world),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def collect_profile(lib, n_prof):
    import ctypes as C
    cap = 40000
    tags, nk, ms = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_float * cap)()
    n = lib.dpmn_profile_collect(tags, nk, ms, cap)
    agg = {}
    for j in range(n):
        name = lib.dpmn_profile_tag_name(tags[j]).decode()
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms[j] / n_prof
        a[1] += nk[j]
    return {k: {"ms_per_step": v[0], "launches_per_step": v[1] // n_prof} for k, v in agg.items()}


def run_train_leg(args, dev, world, rank, lib, dev_sets, host_sets, barrier, max_over_ranks, steps, warmup, want_e2e):
    """BASELINE configs[2]: HotPathTrainer.step at batch B/GPU, the reference's drop rates (0.1), train-mode BatchNorm."""
    from dpmn_b200.pipeline import DPMNHotPath, HostFeeder
    from dpmn_b200.train import HotPathTrainer
    B = args.batch
    n_sets = len(dev_sets)
    model = DPMNHotPath(precision=args.precision, drop=args.train_drop)
    pg, cm = synth_weights(2)
    load_weights(model, pg, cm)
    model = model.to(dev).train()
    trainer = HotPathTrainer(model, overlap=not args.no_overlap)
    with torch.enable_grad():
        for i in range(warmup):
            trainer.step(*dev_sets[i % n_sets])
        barrier()
        launches0 = lib.dpmn_launch_count()
        trainer.timing = {}
        trainer.collect_timing()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            trainer.step(*dev_sets[i % n_sets])
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = lib.dpmn_launch_count() - launches0
        comm_ms = {k: v / steps for k, v in trainer.collect_timing().items()}
        trainer.timing = None
        # end to end: pinned host batches, H2D of step i+1 overlaps step i, the loss is read back every step
        e2e = None
        if want_e2e:
            feeder = HostFeeder(dev)
            out_host = torch.empty((), dtype=torch.float32).pin_memory()

            def run_e2e(n):
                ticket = feeder.stage(list(host_sets[0]))
                for i in range(n):
                    nxt = feeder.stage(list(host_sets[(i + 1) % n_sets])) if i + 1 < n else None
                    y = trainer.step(*feeder.get(ticket))
                    feeder.release(ticket)
                    feeder.fetch(y, out_host)
                    ticket = nxt
                feeder.drain()
            run_e2e(2)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            run_e2e(steps)
            e1.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
            h2d = sum(t.numel() * 4 for t in [host_sets[0][0]] + host_sets[0][1] + host_sets[0][2] + [host_sets[0][3]])
            e2e = {"value": B * world / (ms_e2e / steps / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
        # per-kernel-class timing, outside the throughput regions (every rank takes part: the step has a collective)
        n_prof = 2
        if rank == 0:
            lib.dpmn_profile_enable(1)
        two_streams = model.concurrent_train
        model.concurrent_train = False            # one stream: per-launch event times are not inflated by overlap
        for i in range(n_prof):
            trainer.step(*dev_sets[i % n_sets])
        torch.cuda.synchronize()
        model.concurrent_train = two_streams
        prof = None
        if rank == 0:
            prof = collect_profile(lib, n_prof)
            lib.dpmn_profile_enable(0)
    ms_step = ms_total / steps
    res = {"value": B * world / (ms_step / 1e3), "unit": "images/s", "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
           "global_batch": B * world, "gpu_launches": int(launches),
           "parallelism": f"dp{world}: one flat fp32 gradient bucket ({trainer.state.n_params} parameters, "
                          f"{trainer.state.n_params * 4 / 1e6:.1f} MB), NCCL all-reduce through dpmn_allreduce_bucket"
                          + ("" if world > 1 else " (N=1: no collective issued)"),
           "allreduce": {"world": world, "bucket_bytes": trainer.state.flat_grads.numel() * 4,
                         "overlapped_with_backward": bool(trainer.comm is not None and trainer.overlap),
                         "ms_per_step": {k: round(v, 4) for k, v in comm_ms.items()}},
           "optimizer": "fused clip_grad_norm_(0.25 per module) + Adam over the flat buffers (dpmn_clip_adam_step)" if trainer.fused
                        else "torch clip_grad_norm_ + torch.optim.Adam",
           "drop_rates": args.train_drop, "e2e": e2e,
           "streams": "the two PGRM cascades (forward and backward) on two CUDA streams, CMM / losses on the main stream"
                      if model.concurrent_train else "one stream"}
    if prof is not None:
        res["by_kernel_ms"] = {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])}
        res["whole_step_tflops"] = 3 * FLOPS_IMG["total"] * B / (ms_step / 1e3) / 1e12    # fwd + 2x bwd algorithmic FLOPs
    if trainer.comm is not None:
        trainer.comm.close()
    del trainer, model
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch.distributed as dist
    from dpmn_b200 import _lib
    from dpmn_b200.pipeline import DPMNHotPath
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the hot path has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    assert lib.dpmn_check_device() == 0

    B = args.batch
    n_sets = 4   # rotate input sets; the per-step working set (workspaces ~0.7 GB) is far larger than the 126 MB L2
    host_sets, dev_sets = [], []
    for s in range(n_sets):
        psn, p1, p2 = synth_inputs(1000 * rank + s, B)
        hr = torch.from_numpy(np.random.default_rng([1000 * rank + s, 7]).uniform(0, 1, (B, 4, 32, 128)).astype(np.float32))
        hs = (torch.from_numpy(psn).pin_memory(), [torch.from_numpy(a).pin_memory() for a in p1],
              [torch.from_numpy(a).pin_memory() for a in p2], hr.pin_memory())
        host_sets.append(hs)
        dev_sets.append((hs[0].to(dev), [a.to(dev) for a in hs[1]], [a.to(dev) for a in hs[2]], hs[3].to(dev)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tensor_burst = peaks.get("bf16_tflops", 1640.0)
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    dtype = {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.precision]

    if args.mode == "train":
        tr = run_train_leg(args, dev, world, rank, lib, dev_sets, host_sets, barrier, max_over_ranks, args.steps, args.warmup, True)
        clk = clocks.stop() if rank == 0 else None
        if rank == 0:
            prof = tr.get("by_kernel_ms", {})
            dom = max((k for k in prof if k.startswith("bwd_conv") or k.startswith("bwd_gemm") or k in ("conv_tc", "gemm_tc", "conv", "gemm")),
                      key=lambda k: prof[k], default=None)
            fl = {"bwd_gemm": 2 * _PGRM_GEMM, "bwd_gemm_tc": 2 * _PGRM_GEMM, "bwd_conv": 2 * FLOPS_IMG["conv"],
                  "conv_tc": 2 * FLOPS_IMG["conv"], "gemm_tc": 2 * _PGRM_GEMM, "conv": 2 * FLOPS_IMG["conv"], "gemm": 2 * _PGRM_GEMM}
            roof = None
            if dom:
                ach = fl[dom] * B / (prof[dom] / 1e3) / 1e12
                roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                        "traffic": None, "peak_source": peak_src + "; class time includes the im2col gathers / 16-bit staging copies",
                        "whole_step_tflops": tr.get("whole_step_tflops"), "by_kernel_ms": prof}
            line = {"metric": "SR images/sec (training step)", "value": tr["value"], "unit": "images/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                    "config": {"workload": "DPMN hot path TRAINING step (configs[2] without the frozen TATT backbone and recognisers): "
                                           "6xPGRM (drop rates " + str(args.train_drop) + ") + CMM (train-mode BN), 7 image losses + 4 "
                                           "DistillModule terms, backward, gradient all-reduce, per-module clip 0.25, Adam; batch "
                                           + str(B) + "/GPU", "global_batch": B * world, "parallelism": tr["parallelism"],
                               "l2": "inputs rotate over 4 sets; per-step working set >> 126 MB L2"},
                    "e2e": tr["e2e"], "gpu_launches": tr["gpu_launches"], "clocks": clk, "roofline": roof, "cpu_baseline": None,
                    "train": tr}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ================= inference (configs[1]) =================
    model = DPMNHotPath(precision=args.precision)
    pg, cm = synth_weights(2)
    load_weights(model, pg, cm)
    model = model.to(dev).eval()
    model.alpha = ALPHA
    h2d_bytes = sum(t.numel() * 4 for t in [host_sets[0][0]] + host_sets[0][1] + host_sets[0][2])
    d2h_bytes = B * 3 * 32 * 128 * 4

    with torch.no_grad():
        for i in range(args.warmup):
            model(*dev_sets[i % n_sets][:3])
        from dpmn_b200.pipeline import GraphedHotPath
        n_slots = int(os.environ.get("DPMN_BENCH_SLOTS", "2"))
        if os.environ.get("DPMN_BENCH_ONE_STREAM", "0") == "1":       # A/B: one stream per slot instead of three
            model.concurrent_branches = False
        graphed = GraphedHotPath(model, B, dev, slots=n_slots)      # capture once, outside the timed regions
        for i in range(args.warmup):                                 # warm replays
            graphed.launch(i % n_slots)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        # ---------- value: inputs resident in HBM.  Each batch = a device-to-device copy of its inputs into the slot's static
        # buffers + one CUDA-graph replay; two slots are in flight, the timed region ends when the last is complete
        def run_graphed_resident(n):
            t_host = time.perf_counter()
            begin = torch.cuda.Event()
            begin.record(torch.cuda.current_stream(dev))
            for i in range(n):
                slot = graphed.slots[i % n_slots]
                ds = dev_sets[i % n_sets]
                slot["stream"].wait_event(begin)
                with torch.cuda.stream(slot["stream"]):
                    slot["psn"].copy_(ds[0], non_blocking=True)
                    for d_, s_ in zip(slot["p1"] + slot["p2"], ds[1] + ds[2]):
                        d_.copy_(s_, non_blocking=True)
                graphed.launch(i % n_slots)
            ms = (time.perf_counter() - t_host) * 1e3 / n
            for sl in graphed.slots:
                torch.cuda.current_stream(dev).wait_event(sl["done"])
            return ms
        e0.record()
        host_enqueue_ms = run_graphed_resident(args.steps)
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = args.steps * graphed.kernels_per_replay

        # ---------- e2e: pinned host buffers, H2D + D2H inside the timed region
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        outs_host = [torch.empty((B, 3, 32, 128), dtype=torch.float32).pin_memory() for _ in range(4)]

        def run_e2e(n):
            """n steps from pinned host memory: H2D of step i+1 and D2H of step i-1 overlap the kernels of step i."""
            fetched = [None] * n_slots
            for i in range(n):
                k = i % n_slots
                slot = graphed.slots[k]
                hs = host_sets[i % n_sets]
                with torch.cuda.stream(h2d):
                    if slot["done"] is not None:
                        h2d.wait_event(slot["done"])          # the previous replay of this slot has read its inputs
                    slot["psn"].copy_(hs[0], non_blocking=True)
                    for d_, s_ in zip(slot["p1"] + slot["p2"], hs[1] + hs[2]):
                        d_.copy_(s_, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(h2d)
                if fetched[k] is not None:
                    slot["stream"].wait_event(fetched[k])     # its previous output has been read back
                done = graphed.launch(k, after=ready)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done)
                    outs_host[k].copy_(slot["out"], non_blocking=True)
                    fetched[k] = torch.cuda.Event()
                    fetched[k].record(d2h)
            h2d.synchronize()
            d2h.synchronize()
        run_e2e(min(args.warmup, 3))
        barrier()
        t0 = time.perf_counter()
        e0.record()
        run_e2e(args.steps)          # ends with the last result in host memory
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))

        # ---------- the output of the timed configuration, for the parity diff (rank 0's input set 0 through the graph)
        timed_out = None
        if rank == 0:
            slot = graphed.slots[0]
            ds = dev_sets[0]
            with torch.cuda.stream(slot["stream"]):
                slot["psn"].copy_(ds[0], non_blocking=True)
                for d_, s_ in zip(slot["p1"] + slot["p2"], ds[1] + ds[2]):
                    d_.copy_(s_, non_blocking=True)
            graphed.launch(0).synchronize()
            timed_out = slot["out"].detach().float().cpu().numpy().copy()

        # ---------- "PSN included" (SURVEY 8d config 2): the frozen TATT backbone in front of the hot path.  Per batch: TATT
        # (torch / cuDNN operators replayed from a CUDA graph, dpmn_b200/psn.py) on the LR images + text prior -> its output is
        # the hot path's image stream -> the hot-path graph.  Both run on the slot's stream; two slots are in flight.
        psn_obj = None
        if not args.no_psn:
            from dpmn_b200.psn import TATT
            from dpmn_b200.synth import synth_value
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            tatt = TATT()
            tatt.load_state_dict({k: (v if k.endswith("pe.pe") else torch.from_numpy(np.asarray(synth_value(90, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape))
                                  for k, v in tatt.state_dict().items()}, strict=True)
            tatt = tatt.to(dev)
            lr_sets = []
            for s in range(n_sets):
                r = np.random.default_rng([1000 * rank + s, 41])
                lr = r.uniform(0, 1, size=(B, 4, 16, 64)).astype(np.float32)
                lr[:, 3] = lr[:, 3] > 0.5
                tp = torch.softmax(torch.from_numpy(r.standard_normal((B, 37, 1, 26)).astype(np.float32)), dim=1)
                lr_sets.append((torch.from_numpy(lr).to(dev), tp.to(dev)))

            def run_with_psn(n, with_hot_path=True):
                begin = torch.cuda.Event()
                begin.record(torch.cuda.current_stream(dev))
                for i in range(n):
                    k = i % n_slots
                    slot = graphed.slots[k]
                    ds = dev_sets[i % n_sets]
                    slot["stream"].wait_event(begin)
                    with torch.cuda.stream(slot["stream"]):
                        sr, _ = tatt.graphed(*lr_sets[i % n_sets], slot=k)
                        if with_hot_path:
                            slot["psn"].copy_(sr, non_blocking=True)
                            for d_, s_ in zip(slot["p1"] + slot["p2"], ds[1] + ds[2]):
                                d_.copy_(s_, non_blocking=True)
                        else:
                            slot["done"] = torch.cuda.Event()
                            slot["done"].record(slot["stream"])
                    if with_hot_path:
                        graphed.launch(k)
                for sl in graphed.slots:
                    torch.cuda.current_stream(dev).wait_event(sl["done"])
            run_with_psn(4)
            run_with_psn(4, False)
            barrier()
            n_psn = max(10, args.steps // 2)
            e0.record()
            run_with_psn(n_psn, False)
            e1.record()
            barrier()
            ms_psn = max_over_ranks(e0.elapsed_time(e1)) / n_psn
            e0.record()
            run_with_psn(n_psn)
            e1.record()
            barrier()
            ms_both = max_over_ranks(e0.elapsed_time(e1)) / n_psn
            psn_obj = {"model": "TATT = TSRN_TL_TRANS(scale 2, 5 SRBs, TP interpreter), frozen, eval, fp32 (TF32 off): torch / cuDNN "
                                "operators replayed from a CUDA graph (dpmn_b200/psn.py; parity vs the reference: tests/test_psn.py)",
                       "psn_ms_per_batch": ms_psn, "ms_per_step_psn_included": ms_both,
                       "value_psn_included": B * world / (ms_both / 1e3), "unit": "images/s", "steps": n_psn,
                       "inputs": "LR images U[0,1) (B,4,16,64) with a {0,1} mask channel + text prior softmax(N(0,1)) (B,37,1,26), resident"}
            del tatt

        # ---------- per-kernel-class timing (roofline leg), outside the throughput regions
        prof = None
        if rank == 0:
            n_prof = 2
            model.concurrent_branches = False     # one stream: per-launch event times are not inflated by overlap
            lib.dpmn_profile_enable(1)
            for i in range(n_prof):
                model(*dev_sets[i % n_sets][:3])
            torch.cuda.synchronize()
            prof = collect_profile(lib, n_prof)
            lib.dpmn_profile_enable(0)
    del graphed
    torch.cuda.empty_cache()

    # ================= training (configs[2]) in the same line =================
    train = None
    if not args.no_train:
        t_steps = args.train_steps if args.train_steps > 0 else max(3, min(args.steps, 20))
        train = run_train_leg(args, dev, world, rank, lib, dev_sets, host_sets, barrier, max_over_ranks, t_steps, 3, False)
    clk = clocks.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_val = B * world / (ms_e2e / args.steps / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------- roofline, SURVEY 8d accounting
    flops_img = dict(FLOPS_IMG)
    ridge = tensor_peak * 1e12 / (hbm_peak * 1e9)
    classes = {}
    for k, v in prof.items():
        ms_k = v["ms_per_step"]
        row = {"ms_per_step": round(ms_k, 4), "launches_per_step": v["launches_per_step"]}
        if k in flops_img:
            tf = flops_img[k] * B / (ms_k / 1e3) / 1e12
            row.update({"tflops": round(tf, 2), "tensor_frac_sustained": round(tf / tensor_peak, 4), "tensor_frac_burst": round(tf / tensor_burst, 4)})
        if k in BYTES_IMG:
            gb = BYTES_IMG[k] * B / (ms_k / 1e3) / 1e9
            row.update({"launched_gbs": round(gb, 1), "hbm_frac_of_launched_bytes": round(gb / hbm_peak, 4)})
        classes[k] = row
    total_prof_ms = sum(v["ms_per_step"] for v in prof.values())
    dom = max((k for k in prof if k in flops_img and k != "total"), key=lambda k: prof[k]["ms_per_step"])
    dom_ms = prof[dom]["ms_per_step"]
    achieved = flops_img[dom] * B / (dom_ms / 1e3) / 1e12
    # the PGRM block classes together (every launch between the patch embed and the head) against 8d's fused-block figures
    block_keys = [k for k in ("gemm_tc", "gemm", "mlp_fc1_dw_tc", "dwconv", "window_attn_tc", "window_attn", "sk_gate", "layernorm") if k in prof]
    block_ms = sum(prof[k]["ms_per_step"] for k in block_keys)
    block_flops = 12 * 552_867_840
    traffic = ncu_traffic_per_launch(dom)
    traffic_step = None
    if traffic is not None:
        traffic_step = traffic * prof[dom]["launches_per_step"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "frac_of_burst_peak": achieved / tensor_burst,
                "traffic": traffic, "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                "accounting": "SURVEY 8d: the fused PGRM block is tensor-bound (AI = 552.9 MFLOP / 589 824 B = 937 FLOP/B vs ridge "
                              f"{ridge:.0f}); achieved = algorithmic FLOPs of the class / its summed launch time",
                "share_of_step": dom_ms / total_prof_ms, "launches_per_step": prof[dom]["launches_per_step"],
                "avg_launch_ms": dom_ms / max(1, prof[dom]["launches_per_step"]),
                "pgrm_blocks": {"classes": block_keys, "ms_per_step": round(block_ms, 4),
                                "tflops": round(block_flops * B / (block_ms / 1e3) / 1e12, 2),
                                "tensor_frac_sustained": round(block_flops * B / (block_ms / 1e3) / 1e12 / tensor_peak, 4),
                                "min_bytes_per_step_8d": MIN_BYTES_IMG["pgrm_blocks"] * B,
                                "hbm_frac_at_min_bytes": round(MIN_BYTES_IMG["pgrm_blocks"] * B / (block_ms / 1e3) / 1e9 / hbm_peak, 4)},
                "dram_traffic_per_step": traffic_step,
                "traffic_over_8d_minimum": (traffic_step / (MIN_BYTES_IMG["pgrm_blocks"] * B)) if traffic_step else None,
                "unfused_bytes_per_step": BYTES_IMG.get(dom, 0) * B,
                "whole_step_tflops": flops_img["total"] * B / (ms_step / 1e3) / 1e12,
                "whole_step_tensor_frac": flops_img["total"] * B / (ms_step / 1e3) / 1e12 / tensor_peak,
                "attention": None,
                "by_kernel_ms": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])},
                "classes": classes}
    if "window_attn_tc" in prof:    # the stand-alone attention is HBM-bound (AI 14 FLOP/B): its own roofline is the HBM one
        a_ms = prof["window_attn_tc"]["ms_per_step"]
        gbs = MIN_BYTES_IMG["window_attn_tc"] * B / (a_ms / 1e3) / 1e9
        roofline["attention"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "tflops": FLOPS_IMG["window_attn_tc"] * B / (a_ms / 1e3) / 1e12,
                                 "traffic": ncu_traffic_per_launch("window_attn_tc"),
                                 "avg_launch_ms": a_ms / max(1, prof["window_attn_tc"]["launches_per_step"])}

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        v, sample, _, ref_out = cpu_port_throughput(B, budget_s=25.0, keep_output=True)
        cpu = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample}
        if timed_out is not None and ref_out is not None:
            ref = ref_out.numpy()
            err = float(np.abs(timed_out - ref).max() / np.abs(ref).max())
            bar = {"fp32": 1e-4, "fp16": 1e-3, "bf16": 1e-2}[args.precision]   # BASELINE north_star: 1e-3 rel fp16
            parity = {"max_abs_err_over_max_ref": err, "bar": bar, "ok": bool(err < bar),
                      "what": f"output of the timed CUDA-graph configuration on input set 0 (batch {B}, 6 PGRM + CMM + alpha blend) vs "
                              "oracle/torch_ref.hot_path_forward on the same inputs and weights (fp32 CPU)"}
            if not parity["ok"]:
                print(f"bench.py: PARITY FAILURE in the timed configuration: {err:.3e} >= {bar}", file=sys.stderr)

    # north_star: images/s "as fraction of the attention-FLOP roofline": what the N GPUs could deliver if the step were nothing but
    # its window-attention contractions (132.1 MFLOP and, for the HBM-bound stand-alone kernel, 9.4 MB per image)
    attn_flops_img, attn_bytes_img = FLOPS_IMG["window_attn_tc"], MIN_BYTES_IMG["window_attn_tc"]
    ips_tensor = world * tensor_peak * 1e12 / attn_flops_img
    ips_hbm = world * hbm_peak * 1e9 / attn_bytes_img
    attention_roofline = {"attention_flops_per_image": attn_flops_img, "attention_min_bytes_per_image": attn_bytes_img,
                          "images_per_s_at_tensor_peak": ips_tensor, "frac_of_attention_flop_roofline": value / ips_tensor,
                          "images_per_s_at_hbm_peak": ips_hbm, "frac_of_attention_hbm_roofline": value / ips_hbm,
                          "note": "whole hot-path images/s over the images/s the attention contractions alone would allow on "
                                  f"{world} GPU(s) at the measured tensor peak / at the measured HBM peak (the stand-alone kernel is HBM-bound)"}
    roofline["attention_roofline"] = attention_roofline
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": infer_config(B, world),
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clk, "roofline": roofline,
            "cpu_baseline": cpu, "parity": parity, "psn": psn_obj, "train": train}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--precision", default=os.environ.get("DPMN_PRECISION", "fp16"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] headline line with the configs[2] `train` object inside; train = the training step alone")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-psn", action="store_true", help="skip the 'PSN included' leg (frozen TATT backbone in front of the hot path)")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[2] training leg of the default line")
    ap.add_argument("--no-overlap", action="store_true", help="training: all-reduce the whole bucket after the backward (no overlap)")
    ap.add_argument("--train-steps", type=int, default=0, help="timed training steps of the default line (0 = min(steps, 20))")
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step (48 = configs[1]/[2]; 64 = configs[3])")
    ap.add_argument("--train-drop", type=float, default=0.1, help="drop / attn_drop / drop_path rate of the training leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
