"""The DPMN hot path as its caller runs it: two cascades of three PGRMs over the frozen PSN output, then
the Complementation Modulation Module (interfaces/super_resolution.py:174-265 train, :418-448 eval,
:674-704 test).  Host-side glue only: every module call goes to libdpmn_b200 through the C-ABI."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from .cmm import ComplementationModulationModule
from .pgrm import PGRM

# README.md:42 of the reference: the flags every DPMN run passes (stu_iter_b1 = stu_iter_b2 = 3)
N_ITER = 6


def build_pgrm_stack(precision: str = None, stu_iter_b1: int = 3, stu_iter_b2: int = 3,
                     drop: float = 0.1) -> List[PGRM]:
    """generator_init (interfaces/base.py:150-155) for k = 0..5: branch 1 gets the 2-channel rendered-text
    prior (mode=False), branch 2 the 3-channel mask prior (mode=True); hidden_size = 3.  `drop` is the value of
    --drop_rate / --attn_drop_rate / --drop_path_rate (README.md:42: 0.1).  In train() with non-zero rates every PGRM
    forward runs the fp32 training sequence with Dropout / DropPath masks (see include/dpmn_b200.h); eval() ignores them."""
    n = stu_iter_b1 + stu_iter_b2
    mods = []
    for k in range(n):
        mods.append(PGRM(patch_size=[2] * n, embed_dim=[96] * n, depths=[1] * n, num_heads=[[6]] * n,
                         window_size=[[2, 4, 8]] * n, mlp_ratio=[4.] * n, drop_rate=[drop] * n,
                         attn_drop_rate=[drop] * n, drop_path_rate=[drop] * n, iter=k, mode=(k >= stu_iter_b1),
                         hidden_size=3, precision=precision))
    return mods


class DPMNHotPath(nn.Module):
    def __init__(self, precision: str = None, stu_iter_b1: int = 3, stu_iter_b2: int = 3, drop: float = 0.1,
                 cmm_precision: str = None):
        super().__init__()
        self.b1, self.b2 = stu_iter_b1, stu_iter_b2
        self.pgrm = nn.ModuleList(build_pgrm_stack(precision, stu_iter_b1, stu_iter_b2, drop))
        self.cmm = ComplementationModulationModule(precision=cmm_precision or precision)
        self.concurrent_branches = True       # the two PGRM cascades run on two CUDA streams (the CMM on a third)
        # ... also under autograd (training): every backward node runs on the stream of its forward, so the two cascades'
        # backwards overlap the same way (DPMN_TRAIN_STREAMS=0: one stream, the A/B switch)
        import os
        self.concurrent_train = os.environ.get("DPMN_TRAIN_STREAMS", "1") != "0"
        self._streams = None
        # --alpha of the reference (main.py:66, README.md:42: 0.5): eval / test blend the CMM output with the PSN image
        # (super_resolution.py:449,705).  None = the plain CMM output (what the training loss looks at).
        self.alpha = None

    def forward(self, psn_out: torch.Tensor, priors_b1: Sequence[torch.Tensor], priors_b2: Sequence[torch.Tensor]):
        """psn_out (B,4,32,128) frozen-backbone output; priors_b1[k] (B,2,32,128) rendered-text maps;
        priors_b2[k] (B,3,32,128) binary masks -> fused SR image (B,3,32,128)."""
        return self.forward_all(psn_out, priors_b1, priors_b2)[-1]

    def _cmm_call(self, psn_out, b1_out, b2_out):
        if self.alpha is None or self.training:
            return self.cmm(b1_out, b2_out)                                                   # :265
        return self.cmm(b1_out, b2_out, blend_with=psn_out[:, :3], alpha=self.alpha)          # :448-449, :704-705

    def _branch_fn(self, psn_out):
        """One cascade of PGRMs (super_resolution.py:174-240) as a closure over the PSN output; shared by forward_all and
        submit so that both accept priors=None for branch 2 (device toMask of the cascade image, :218-226)."""
        from .train import to_mask

        def branch(first, count, priors):
            cascade = psn_out[:, :3, :]                   # channel-slice view, super_resolution.py:196
            done: List[torch.Tensor] = []
            for j in range(count):
                k = first + j
                # branch 2 without given priors: x_q = toMask(current cascade image), on the device (:218-226)
                x_q = priors[j] if priors is not None else to_mask(cascade)
                # residual_list = the earlier outputs of this branch (super_resolution.py:207,234)
                y = self.pgrm[k](x_q, cascade, done[:j])
                done.append(y)
                cascade = y
            return done
        return branch

    def forward_all(self, psn_out, priors_b1, priors_b2) -> List[torch.Tensor]:
        """All seven images the training loss looks at (super_resolution.py:212,239,267): the six PGRM outputs in
        call order, then the CMM output.  priors_b2=None: the branch-2 priors are computed from the cascade images
        with the device toMask, as the reference does on the host."""
        branch = self._branch_fn(psn_out)

        if psn_out.is_cuda and self.concurrent_branches and (not torch.is_grad_enabled() or self.concurrent_train):
            outs, done = self._submit(psn_out, priors_b1, priors_b2, branch)
            main = torch.cuda.current_stream(psn_out.device)
            main.wait_event(done)                 # ordinary stream semantics for the caller: results are ready on `main`
            if not torch.cuda.is_current_stream_capturing():
                for t in outs:
                    t.record_stream(main)
            return outs
        done1 = branch(0, self.b1, priors_b1)
        done = branch(self.b1, self.b2, priors_b2)
        return list(done1) + list(done) + [self._cmm_call(psn_out, done1[-1], done[-1])]                  # :265

    def _submit(self, psn_out, priors_b1, priors_b2, branch):
        """The two cascades are independent until the CMM (super_resolution.py:174-240): they run on two side streams
        (the tail of one branch's kernel overlaps the launch + prologue of the other's) and the CMM on a third, so a
        caller that keeps several batches in flight (`submit`) also overlaps the CMM's small deep layers with the next
        batch's PGRMs.  Workspaces are per stream (pgrm.workspace) and staged weights per module: nothing is shared."""
        dev = psn_out.device
        main = torch.cuda.current_stream(dev)
        if self._streams is None or self._streams[0].device != dev:
            self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
        s_cmm = self._streams[2]
        start = torch.cuda.Event()
        start.record(main)
        results = []
        for st, (first, count, priors) in zip(self._streams[:2], ((0, self.b1, priors_b1), (self.b1, self.b2, priors_b2))):
            st.wait_event(start)
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                for t in [psn_out] + list(priors or []):
                    t.record_stream(st)
            with torch.cuda.stream(st):
                done = branch(first, count, priors)
            if not capturing:
                for t in list(done) + [psn_out]:      # the CMM stream reads the branch outputs and (alpha blend) the PSN image
                    t.record_stream(s_cmm)
            end = torch.cuda.Event()
            end.record(st)
            s_cmm.wait_event(end)
            results.append(done)
        with torch.cuda.stream(s_cmm):
            y = self._cmm_call(psn_out, results[0][-1], results[1][-1])                                   # :265
        finished = torch.cuda.Event()
        finished.record(s_cmm)
        return list(results[0]) + list(results[1]) + [y], finished

    @torch.no_grad()
    def submit(self, psn_out, priors_b1, priors_b2):
        """Asynchronous inference: enqueue one batch without making the current stream wait for it.  Returns
        (sr, done): `sr` (B,3,32,128) is valid once the CUDA event `done` has fired -- make the consuming stream
        `wait_event(done)` (HostFeeder.fetch does).  Batches are processed in submission order."""
        outs, finished = self._submit(psn_out, priors_b1, priors_b2, self._branch_fn(psn_out))
        return outs[-1], finished


class GraphedHotPath:
    """Inference through CUDA graphs: the ~170 kernel launches of one forward (three internal streams) are captured
    once per slot and replayed with one cudaGraphLaunch, which takes the host out of the critical path (enqueueing
    the launches from Python costs ~3.2 ms per batch, about as much as the GPU needs to run them).

    Each of the `slots` graphs owns static input / output tensors and its own workspaces, and replays on its own
    stream, so consecutive batches overlap on the GPU exactly as with `DPMNHotPath.submit`.  Usage per batch:
    write the inputs of slot s (`inputs(s)`, e.g. by H2D copies), `launch(s, after=<event the inputs are ready>)`,
    read `output(s)` once the returned event has fired; reuse slot s only after that.  Weights are baked in by
    address: re-capture (`GraphedHotPath(model, ...)`) after loading new weights."""

    def __init__(self, model: DPMNHotPath, batch: int, device: torch.device, slots: int = 2, img=(32, 128)):
        from . import pgrm as _pgrm
        self.model, self.device = model, device
        H, W = img
        self.slots = []
        with torch.no_grad():
            warm = (torch.zeros(batch, 4, H, W, device=device), [torch.zeros(batch, 2, H, W, device=device) for _ in range(model.b1)],
                    [torch.zeros(batch, 3, H, W, device=device) for _ in range(model.b2)])
            for _ in range(2):                      # stage the 16-bit weights, set kernel attributes, size workspaces
                model(*warm)
            torch.cuda.synchronize(device)
            for _ in range(slots):
                psn = torch.zeros(batch, 4, H, W, device=device)
                p1 = [torch.zeros(batch, 2, H, W, device=device) for _ in range(model.b1)]
                p2 = [torch.zeros(batch, 3, H, W, device=device) for _ in range(model.b2)]
                stream = torch.cuda.Stream(device)
                graph = torch.cuda.CUDAGraph()
                saved = dict(_pgrm._WORKSPACES)
                _pgrm._WORKSPACES.clear()           # this graph gets private workspaces (allocated in its own pool)
                from . import _lib
                n0 = _lib.load().dpmn_launch_count()
                try:
                    with torch.cuda.graph(graph, stream=stream):
                        out = model(psn, p1, p2)
                finally:
                    self.kernels_per_replay = int(_lib.load().dpmn_launch_count() - n0)
                    _pgrm._WORKSPACES.clear()
                    _pgrm._WORKSPACES.update(saved)
                self.slots.append({"psn": psn, "p1": p1, "p2": p2, "out": out, "graph": graph, "stream": stream,
                                   "done": None})
        torch.cuda.synchronize(device)

    def inputs(self, slot: int):
        s = self.slots[slot]
        return s["psn"], s["p1"], s["p2"]

    def output(self, slot: int) -> torch.Tensor:
        return self.slots[slot]["out"]

    def launch(self, slot: int, after: torch.cuda.Event = None) -> torch.cuda.Event:
        s = self.slots[slot]
        st = s["stream"]
        if after is None:
            after = torch.cuda.Event()
            after.record(torch.cuda.current_stream(self.device))
        st.wait_event(after)
        with torch.cuda.stream(st):
            s["graph"].replay()
            done = torch.cuda.Event()
            done.record(st)
        s["done"] = done
        return done


class HostFeeder:
    """Host <-> device staging for callers that hold their batches in (pinned) host memory: inputs of step i+1 are
    copied on a side stream while step i computes, and the result of step i is read back on another side stream, so
    PCIe traffic overlaps the kernels instead of serialising with them.  Plumbing only (streams + events)."""

    def __init__(self, device: torch.device, depth: int = 2):
        self.device = device
        self.depth = depth
        self.h2d = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)
        self.slots = [None] * depth          # device tensors per slot
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth       # event: the compute that read the slot has been enqueued and finished
        self.n = 0
        self.out_done = None

    @staticmethod
    def _flatten(x):
        return [x] if torch.is_tensor(x) else [t for y in x for t in HostFeeder._flatten(y)]

    @staticmethod
    def _rebuild(x, it):
        return next(it) if torch.is_tensor(x) else [HostFeeder._rebuild(y, it) for y in x]

    def stage(self, host_batch):
        """Enqueue the H2D copies of a (nested list of) host tensors; returns a ticket for `get`."""
        slot = self.n % self.depth
        self.n += 1
        flat = self._flatten(host_batch)
        if self.slots[slot] is None or [t.shape for t in self.slots[slot]] != [t.shape for t in flat]:
            self.slots[slot] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in flat]
        with torch.cuda.stream(self.h2d):
            if self.consumed[slot] is not None:
                self.h2d.wait_event(self.consumed[slot])       # do not overwrite inputs a running step still reads
            for dst, src in zip(self.slots[slot], flat):
                dst.copy_(src, non_blocking=True)
            self.ready[slot].record(self.h2d)
        return slot, host_batch

    def get(self, ticket):
        """Device tensors of a staged batch, ordered after their copies on the current stream."""
        slot, structure = ticket
        torch.cuda.current_stream(self.device).wait_event(self.ready[slot])
        return self._rebuild(structure, iter(self.slots[slot]))

    def release(self, ticket, after: torch.cuda.Event = None):
        """Call after the step that consumed the batch has been enqueued (`after`: the event that marks its end)."""
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self.consumed[ticket[0]] = ev

    def fetch(self, result: torch.Tensor, host_out: torch.Tensor, after: torch.cuda.Event = None):
        """Read a result back into pinned host memory on the D2H stream, after the event `after` (default: everything
        enqueued on the current stream so far)."""
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ev)
            result.record_stream(self.d2h)
            host_out.copy_(result, non_blocking=True)
            self.out_done = torch.cuda.Event()
            self.out_done.record(self.d2h)

    def drain(self):
        self.h2d.synchronize()
        self.d2h.synchronize()
