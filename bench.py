#!/usr/bin/env python
"""bench.py -- SR images/sec of the DPMN hot path (6 x PGRM + CMM forward, batch 48/GPU) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|fp16|bf16] [--impl ours|reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definitions of every key.
  value          images/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e            images/s through the public module API from pinned HOST buffers, H2D + D2H inside the timed region
  roofline       the dominant kernel class, timed per launch with CUDA events by the library's profile hook
  cpu_baseline   the torch-CPU port of the reference (oracle/torch_ref.py) on this host's cores, bounded sample
--impl reference times that same CPU port as the reference arm (the reference is pure PyTorch and
/root/reference does not exist on the GPU box; oracle/torch_ref.py is pinned to its outputs).
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
METRIC = "SR images/sec"
WORKLOAD = "DPMN hot path forward: 6xPGRM cascade (b1=b2=3, win 2/4/8, dim 96, 6 heads) + CMM, 16x64 -> 32x128, " \
           "batch 48/GPU, synthetic PSN output + priors (configs[1] without the frozen TATT backbone)"

# algorithmic FLOPs per image (2*MAC, matmul/conv only; SURVEY.md 8d / BASELINE.md section 3)
_PGRM_GEMM = 6 * 2 * (56_623_104 + 25_165_824 + 460_062_720 - 7_077_888)   # q/kv + SK proj + Mlp GEMMs (dw conv excluded)
FLOPS_IMG = {"gemm": _PGRM_GEMM, "gemm_tc": _PGRM_GEMM,
             "window_attn": 6 * 2 * 11_010_048,
             "conv": 4_459_069_440, "conv_tc": 4_459_069_440 - 14_155_776 * 2 - 42_467_328,   # minus stem / de_1 (own kernels)
             "total": 11_267_776_512}

# algorithmic HBM bytes per image of the bandwidth-bound kernel classes in the 16-bit inference mode:
# 16-bit activations (s = 2), fp32 residual stream (r = 4), weights excluded (L2-resident, < 0.6 MB per block).
_L, _C, _HID = 1024, 96, 384
BYTES_IMG = {
    # per block: q (in, out), kv (in, 2 out), SK pooled pass (in), SK output GEMM (in, LN out) and fc2 (LN out) = 10 L*C*s;
    # the two residual GEMMs read and write the fp32 stream = 4 L*C*r; fc1 out, pointwise in + out, fc2 in = 4 L*hid*s
    "gemm_tc": 12 * (10 * _L * _C * 2 + 4 * _L * _C * 4 + 4 * _L * _HID * 2),
    "window_attn_tc": 12 * 4 * _L * _C * 2,          # q, k, v in, out: SURVEY 8d
    "dwconv": 12 * 2 * _L * _HID * 2,
}


def ncu_traffic_per_launch(kernel_class):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel class, read from the committed summaries of
    the `ncu --set full` captures (profiles/*.txt, written by tools/ncu_summary.py); None when the class was not captured."""
    files = {"gemm_tc": ("r01_ncu_full_fp16_v14.txt", ("gemm_tc_kernel", "gemm_res_ln_kernel")),
             "conv_tc": ("r01_ncu_full_conv_tc_v21.txt", ("conv_tc_kernel<64>", "conv_tc_kernel<128>")),
             "dwconv": ("r01_ncu_full_dwconv_v2_en1_v23.txt", ("dwconv16_v2_kernel",)),
             "window_attn_tc": ("r01_ncu_full_fp16_v14.txt", ("attn_tc_kernel",))}
    if kernel_class not in files:
        return None
    name, kernels = files[kernel_class]
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
        return float(np.mean(vals)) if vals else None
    except Exception:           # evidence lookup must never take the measurement down
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
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


def cpu_port_throughput(budget_s=20.0, steps=None, warmup=1):
    """Time oracle/torch_ref.hot_path_forward on the host cores; returns (images/s, description, ms/step, sample B)."""
    from oracle import torch_ref
    torch.set_num_threads(os.cpu_count() or 1)
    pg, cm = synth_weights(2)
    pg = [{k: torch.from_numpy(np.asarray(v)) for k, v in p.items()} for p in pg]
    cm = {k: torch.from_numpy(np.asarray(v)) for k, v in cm.items()}
    Bs = 8   # bounded sample of the batch-48 workload (same per-image work; the path has no cross-image coupling)
    psn, p1, p2 = synth_inputs(1, Bs)
    args = (torch.from_numpy(psn), [torch.from_numpy(a) for a in p1], [torch.from_numpy(a) for a in p2])
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            torch_ref.hot_path_forward(pg, cm, *args)
        t_end = time.perf_counter() + budget_s
        n = 0
        while (steps is None and time.perf_counter() < t_end and n < 50) or (steps is not None and n < steps):
            t0 = time.perf_counter()
            torch_ref.hot_path_forward(pg, cm, *args)
            times.append(time.perf_counter() - t0)
            n += 1
    med = float(np.median(times))
    return Bs / med, f"{len(times)} timed forwards of a batch-{Bs} slice of the workload (median), torch {torch.__version__} " \
                     f"CPU fp32, {torch.get_num_threads()} threads", med * 1e3, Bs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sample, ms, Bs = cpu_port_throughput(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_batch": Bs, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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

    train = args.mode == "train"
    if train:
        # configs[2]: the three drop rates at the reference's 0.1 (README.md:42) -> every PGRM forward is the fp32
        # training sequence with Dropout / DropPath masks; CMM train-mode BatchNorm; fp32 storage everywhere, the GEMMs and
        # convs of both modules on tcgen05 with 16-bit staged operands
        model = DPMNHotPath(precision=args.precision, drop=args.train_drop)
    else:
        model = DPMNHotPath(precision=args.precision)
    pg, cm = synth_weights(2)
    load_weights(model, pg, cm)
    model = model.to(dev)
    model.train(train)
    trainer = None
    if train:
        from dpmn_b200.train import HotPathTrainer
        trainer = HotPathTrainer(model)

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
    h2d_bytes = sum(t.numel() * 4 for t in [host_sets[0][0]] + host_sets[0][1] + host_sets[0][2] + ([host_sets[0][3]] if train else []))
    out_host = torch.empty((), dtype=torch.float32).pin_memory()      # training: the loss
    d2h_bytes = 4 if train else B * 3 * 32 * 128 * 4

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

    def step_resident(i):
        ds = dev_sets[i % n_sets]
        if train:
            return trainer.step(*ds)
        return model(*ds[:3])

    with torch.set_grad_enabled(train):
        # ---------- value: inputs resident in HBM
        for i in range(args.warmup):
            step_resident(i)
        graphed = None
        if not train:
            from dpmn_b200.pipeline import GraphedHotPath
            n_slots = int(os.environ.get("DPMN_BENCH_SLOTS", "2"))
            graphed = GraphedHotPath(model, B, dev, slots=n_slots)      # capture once, outside the timed regions
            for i in range(args.warmup):                          # warm replays
                graphed.launch(i % n_slots)
        barrier()
        launches0 = lib.dpmn_launch_count()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_enqueue_ms = None
        e0.record()
        if train:
            for i in range(args.steps):
                step_resident(i)
        else:
            # throughput: each batch = a device-to-device copy of its (HBM-resident) inputs into the slot's static
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
            host_enqueue_ms = run_graphed_resident(args.steps)
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = (lib.dpmn_launch_count() - launches0) if train else args.steps * graphed.kernels_per_replay
        clk = clocks.stop() if rank == 0 else None
        # ---------- e2e: pinned host buffers, H2D + D2H inside the timed region
        from dpmn_b200.pipeline import HostFeeder
        feeder = HostFeeder(dev)
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        outs_host = [torch.empty((B, 3, 32, 128), dtype=torch.float32).pin_memory() for _ in range(4)]

        def run_e2e(n):
            """n steps from pinned host memory: H2D of step i+1 and D2H of step i-1 overlap the kernels of step i."""
            if train:
                def batch(i):
                    hs = host_sets[i % n_sets]
                    return [hs[0], hs[1], hs[2], hs[3]]
                ticket = feeder.stage(batch(0))
                for i in range(n):
                    nxt = feeder.stage(batch(i + 1)) if i + 1 < n else None
                    y = trainer.step(*feeder.get(ticket))
                    feeder.release(ticket)
                    feeder.fetch(y, out_host)
                    ticket = nxt
                feeder.drain()
                return
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
        # ---------- per-kernel-class timing (roofline leg), outside the throughput regions
        prof = None
        n_prof = 2
        if rank != 0 and train and world > 1:
            for i in range(n_prof):       # the training step contains a collective: every rank has to take part
                step_resident(i)
            torch.cuda.synchronize()
        if rank == 0:
            if not train:
                model.concurrent_branches = False     # one stream: per-launch event times are not inflated by overlap
            lib.dpmn_profile_enable(1)
            for i in range(n_prof):
                step_resident(i)
            torch.cuda.synchronize()
            import ctypes as C
            cap = 20000
            tags, nk, ms = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_float * cap)()
            n = lib.dpmn_profile_collect(tags, nk, ms, cap)
            lib.dpmn_profile_enable(0)
            agg = {}
            for j in range(n):
                name = lib.dpmn_profile_tag_name(tags[j]).decode()
                a = agg.setdefault(name, [0.0, 0])
                a[0] += ms[j] / n_prof
                a[1] += nk[j]
            prof = {k: {"ms_per_step": v[0], "launches_per_step": v[1] // n_prof} for k, v in agg.items()}

    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_val = B * world / (ms_e2e / args.steps / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks else "fallback"
    flops_img = dict(FLOPS_IMG)
    if train:
        # backward classes: data + weight gradient = 2x the forward contraction FLOPs of the same layers
        flops_img = {"bwd_gemm": 2 * _PGRM_GEMM, "bwd_conv": 2 * FLOPS_IMG["conv"], "gemm": 2 * _PGRM_GEMM + 0,
                     "conv": 2 * FLOPS_IMG["conv"], "total": 3 * FLOPS_IMG["total"] + FLOPS_IMG["total"]}
        # ("gemm"/"conv": fp32 forward recompute inside backward (+ the CMM's own fp32 forward); total = fwd + recompute + bwd)
    dom = max((k for k in prof if k in flops_img and k != "total"), key=lambda k: prof[k]["ms_per_step"])
    dom_ms = prof[dom]["ms_per_step"]
    achieved = flops_img[dom] * B / (dom_ms / 1e3) / 1e12
    total_prof_ms = sum(v["ms_per_step"] for v in prof.values())
    train_bound = "tensor"
    if train and args.precision == "fp32":
        tensor_peak = 75.0   # B200 fp32 FFMA peak (148 SMs x 128 lanes x 2 x 1.965 GHz): the fp32 mode is all SIMT
        peak_src = "nominal fp32 FFMA peak (precision fp32: every contraction is an FFMA kernel)"
        train_bound = "fp32-simt"
    elif train:
        # 16-bit modes: the contractions of this class run on tcgen05 through 16-bit staged operands (lin_*_tc, the CMM's
        # im2col + NT GEMM), so the class -- gathers, staging copies and GEMMs together -- is quoted against the tensor peak
        peak_src += "; class time includes the im2col gathers / 16-bit staging copies around the tcgen05 GEMMs"
    # DRAM bytes per launch of the dominant class from the committed `ncu --set full` capture (mean over its launches in
    # profiles/r01_ncu_full_fp16_v14.txt: dram__bytes_read.sum + dram__bytes_write.sum; writes mostly stay in the 126 MB L2)
    traffic = ncu_traffic_per_launch(dom) if not train else None
    roofline = {"bound": "tensor" if not train else train_bound, "kernel": dom, "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "traffic": traffic, "peak_source": peak_src,
                "share_of_step": dom_ms / total_prof_ms,
                "launches_per_step": prof[dom]["launches_per_step"],
                "avg_launch_ms": dom_ms / max(1, prof[dom]["launches_per_step"]),
                "whole_step_tflops": flops_img["total"] * B / (ms_step / 1e3) / 1e12,
                "by_kernel_ms": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms_per_step"])}}
    if not train and dom in BYTES_IMG:
        # The dominant class is bandwidth-bound by arithmetic intensity (FLOPs / algorithmic bytes below the ridge
        # tensor_peak / hbm_peak): its roofline is the HBM one.  achieved = algorithmic bytes per launch / launch time.
        hbm_peak = peaks.get("hbm_gbs", 6500.0)
        ai = flops_img[dom] / BYTES_IMG[dom]
        if ai < tensor_peak * 1e12 / (hbm_peak * 1e9):
            gbs = BYTES_IMG[dom] * B / (dom_ms / 1e3) / 1e9
            roofline.update({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "peak_source": ("measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback"),
                             "algorithmic_bytes_per_launch": BYTES_IMG[dom] * B / max(1, prof[dom]["launches_per_step"]),
                             "arithmetic_intensity_flop_per_byte": ai,
                             "tensor_view": {"achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                                             "frac": achieved / tensor_peak}})

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not train:
        v, sample, _, _ = cpu_port_throughput(budget_s=15.0)
        cpu = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample}

    workload = WORKLOAD.replace("batch 48/GPU", f"batch {B}/GPU") if not train else (
        "DPMN hot path TRAINING step: 6xPGRM cascade (drop / attn_drop / drop_path " + str(args.train_drop) + ") + CMM forward (train-mode BN), 7 image "
        "losses + 4 DistillModule terms, backward through dpmn_pgrm_backward / dpmn_cmm_backward (fp32), flat gradient all-reduce, per-module "
        "clip 0.25, Adam; 16x64 -> 32x128, batch " + str(B) + "/GPU, synthetic PSN output / priors / HR (configs[2] without the "
        "frozen TATT backbone and recognisers)")
    line = {"metric": METRIC if not train else "SR images/sec (training step)", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.precision],
            "data": "synthetic",
            "config": {"workload": workload, "global_batch": B * world,
                       "parallelism": f"replicas x{world} (no collective: inference)" if not train else
                       f"dp{world}: one flat NCCL all-reduce over the 57.2M-parameter fp32 gradient bucket per step",
                       "l2": f"inputs rotate over {n_sets} sets; per-step working set (activations/workspace) >> 126 MB L2"},
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
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
                    help="infer = BASELINE configs[1] (the headline line); train = configs[2] training step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step (48 = configs[1]/[2]; 64 = configs[3])")
    ap.add_argument("--train-drop", type=float, default=0.1, help="drop / attn_drop / drop_path rate of --mode train")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
