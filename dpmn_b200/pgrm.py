"""Drop-in `PGRM` module: the reference's constructor, forward signature and state_dict schema
(/root/reference/model/pgrm.py:460-565) over the libdpmn_b200 C-ABI.

The module owns parameters only; every FLOP of `forward` runs in the CUDA library.  There is no
PyTorch/CPU compute path: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .schema import PGRMConfig, pgrm_schema, resolve_pgrm_config

_WORKSPACES = {}   # (device index) -> uint8 tensor, shared by all modules (launches are stream-ordered)


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


class ParamTree(nn.Module):
    """Bare container: parameters/buffers are attached under the dotted names of a state_dict schema."""

    def attach(self, dotted: str, value, kind: str):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamTree())
            node = node._modules[p]
        if kind == "param":
            node.register_parameter(parts[-1], nn.Parameter(value))
        else:
            node.register_buffer(parts[-1], value)

    def fetch(self, dotted: str) -> torch.Tensor:
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            node = node._modules[p]
        t = node._parameters.get(parts[-1])
        return t if t is not None else node._buffers[parts[-1]]


def relative_position_index(ws: int) -> np.ndarray:
    """(N, N) int64: idx(n, m) = (i_n - i_m + ws-1)(2ws-1) + (j_n - j_m + ws-1)   (pgrm.py:133-145).
    State-dict compatibility only -- the kernels compute this in closed form."""
    i, j = np.divmod(np.arange(ws * ws), ws)
    return ((i[:, None] - i[None, :] + ws - 1) * (2 * ws - 1) + (j[:, None] - j[None, :] + ws - 1)).astype(np.int64)


def shift_mask(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """(nW, N, N) fp32 in {0, -100} (pgrm.py:153-173).  State-dict compatibility only."""
    def region(x, n):
        return (x >= n - ws).astype(np.int64) + (x >= n - shift).astype(np.int64)
    lab = 3 * region(np.arange(H), H)[:, None] + region(np.arange(W), W)[None, :]
    lab = lab.reshape(H // ws, ws, W // ws, ws).transpose(0, 2, 1, 3).reshape(-1, ws * ws)
    return np.where(lab[:, None, :] != lab[:, :, None], -100.0, 0.0).astype(np.float32)


def _initial_value(name: str, shape, cfg: PGRMConfig) -> torch.Tensor:
    """Reference initialisation (pgrm.py:130,496-497,524-533): trunc_normal(.02) for Linear weights and
    bias tables, zeros for Linear/LayerNorm biases, ones for LayerNorm weights and weight_list_*,
    xavier_uniform for conv weights, torch's default uniform for conv biases."""
    leaf = name.rsplit(".", 1)[-1]
    t = torch.empty(tuple(shape), dtype=torch.float32)
    is_conv = any(k in name for k in ("prior_fusion", "patch_embed.proj", "depthwise_conv", "pointwise_conv",
                                      "conv_before_upsample"))
    if name.startswith("weight_list_"):
        return t.fill_(1.0)
    if "relative_position_bias_table" in name:
        return nn.init.trunc_normal_(t, std=.02)
    if is_conv:
        if leaf == "weight":
            return nn.init.xavier_uniform_(t)
        fan_in = {"prior_fusion": 2 * 9, "patch_embed.proj": cfg.in_chans * cfg.patch_size ** 2,
                  "depthwise_conv": 9, "pointwise_conv": cfg.mlp_hidden}
        fi = next((v for k, v in fan_in.items() if k in name), None)
        if fi is None:   # conv_before_upsample.{0,1}
            fi = (cfg.embed_dim if name.startswith("conv_before_upsample.0") else
                  cfg.hidden_size * cfg.patch_size ** 2) * 9
        bound = 1.0 / math.sqrt(fi)
        return t.uniform_(-bound, bound)
    if len(shape) == 1:   # LayerNorm weight / any bias
        return t.fill_(1.0) if (leaf == "weight") else t.zero_()
    return nn.init.trunc_normal_(t, std=.02)   # Linear weight


def resolve_precision(precision):
    """None -> $DPMN_PRECISION -> "fp16" (the fast, 1e-3-parity mode); explicit values are validated."""
    import os
    p = precision if precision is not None else os.environ.get("DPMN_PRECISION", "fp16")
    if p not in _lib.PREC:
        raise ValueError(f"dpmn_b200: precision must be one of {sorted(_lib.PREC)}, got {p!r}")
    return p


def grad_sink_views(module: nn.Module, names, params):
    """`module._grad_sink` (set by dpmn_b200.train.FlatTrainState): name -> fp32 view of the caller's flat gradient bucket for
    EVERY trainable parameter.  The C backward accumulates into its `grads` pointers, so with a sink the gradients land in
    the bucket directly -- no temporary flat buffer, no autograd AccumulateGrad pass over 57 M elements.  None = off."""
    sink = getattr(module, "_grad_sink", None)
    if sink is None:
        return None
    for n, p in zip(names, params):
        v = sink.get(n)
        if v is None or v.device != p.device or v.numel() != p.numel():
            return None               # incomplete / stale sink: fall back to the autograd-accumulated path
    return sink


class PreparedWeights:
    """Cache of the tensor-core modes' staged 16-bit weights (include/dpmn_b200.h `prepared`): re-staged only
    when a parameter / buffer was modified in place (torch's version counter) or re-allocated."""

    def __init__(self):
        self.buf = None
        self.key = None

    def attach(self, module: nn.Module, d, nbytes: int, device):
        if nbytes == 0:
            return None
        # `_weights_epoch`: bumped by callers that update the parameters through raw pointers (the fused clip + Adam
        # kernel of dpmn_b200.train), which torch's version counters cannot see
        key = (getattr(module, "_weights_epoch", 0),) + tuple(
            (t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self.key = None
        d.prepared = self.buf.data_ptr()
        d.prepared_valid = int(self.key == key)
        return key


_BLOCK_FIELDS = (("norm1_q", "norm1_q"), ("norm1_kv", "norm1_kv"), ("norm2", "norm2"), ("q", "attn.q"), ("kv", "attn.kv"),
                 ("sk_proj", "attn.sknet.proj"), ("sk_fc1", "attn.sknet.fc1"), ("sk_fc2", "attn.sknet.fc2"),
                 ("sk_head", "attn.sknet.proj_head"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2"),
                 ("dw", "mlp.depthwise_conv"), ("pw", "mlp.pointwise_conv"))


class _PGRMFunction(torch.autograd.Function):
    """Autograd node of PGRM.forward: forward = dpmn_pgrm_forward, backward = dpmn_pgrm_backward (which recomputes
    the fp32 forward internally, so only the INPUTS are saved).  Differentiable inputs: x_kv, residual_list[1:],
    every parameter -- the leaves the reference's loss.backward() reaches (super_resolution.py:245-275)."""

    @staticmethod
    def forward(ctx, module, x_q, x_kv, n_res, *rest):
        residuals = list(rest[:n_res])
        seed = module._new_seed() if module._stochastic() else None    # train-mode Dropout / DropPath masks
        keep_ws = []
        with torch.no_grad():
            # a stochastic forward is the fp32 training sequence: run it in a private buffer laid out as the backward's
            # workspace, which then finds every intermediate in place (DPMN_PGRM_WORKSPACE_HOLDS_FORWARD)
            out = module._run(x_q, x_kv, residuals, probe=False, seed=seed, keep_ws=keep_ws if seed is not None else None)
        ctx.module, ctx.n_res, ctx.seed = module, n_res, seed
        ctx.fwd_ws = keep_ws[0] if keep_ws else None
        ctx.save_for_backward(x_q, x_kv, *residuals)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_q, x_kv, *residuals = ctx.saved_tensors
        m = ctx.module
        need_xkv = ctx.needs_input_grad[2]
        need_res = [ctx.needs_input_grad[4 + i] for i in range(ctx.n_res)]
        d_xkv, d_res, d_params = m._backward(x_q, x_kv, residuals, d_out, need_xkv, need_res, seed=ctx.seed,
                                             fwd_ws=ctx.fwd_ws)
        ctx.fwd_ws = None
        return (None, None, d_xkv, None, *d_res, *d_params)


class PGRM(ParamTree):
    """Prior-Guided Refinement Module, signature-compatible with the reference (pgrm.py:462-467).

    Extra keyword: `precision` in {"fp32", "fp16", "bf16"} selects the arithmetic of the contractions
    (fp32 = FFMA, 1e-5 parity; fp16/bf16 = tcgen05 tensor cores with fp32 accumulation, 1e-3 parity).  Default (None): the
    environment variable DPMN_PRECISION, else "fp16" -- the two-import swap of INTEGRATION.md lands on the tensor-core path."""

    def __init__(self, img_size=[32, 128], patch_size=[2], in_chans=3, embed_dim=[96], depths=[1], num_heads=[[6]],
                 window_size=[[2, 4, 8]], mlp_ratio=[4.], qkv_bias=True, qk_scale=None, drop_rate=[0.],
                 attn_drop_rate=[0.], drop_path_rate=[0.1], iter=0, norm_layer=nn.LayerNorm, ape=False,
                 patch_norm=True, mode=True, use_checkpoint=False, hidden_size=64, precision=None, **kwargs):
        super().__init__()
        if not qkv_bias or qk_scale is not None or ape or not patch_norm or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("dpmn_b200.PGRM supports the configuration DPMN instantiates "
                                      "(qkv_bias=True, qk_scale=None, ape=False, patch_norm=True, LayerNorm)")
        self.cfg = resolve_pgrm_config(img_size, patch_size, in_chans, embed_dim, depths, num_heads, window_size,
                                       mlp_ratio, iter, mode, hidden_size)
        self.iter = iter
        self.mode = mode
        self.drop_rate = float(drop_rate[iter])
        self.attn_drop_rate = float(attn_drop_rate[iter])
        # stochastic depth schedule (pgrm.py:499,512): linspace over sum(depths)*2, sliced per layer
        n = sum(depths) * 2
        dpr = [float(x) for x in np.linspace(0, drop_path_rate[iter], n)]
        lo = sum(depths[:iter]) * 2
        self.drop_path = dpr[lo: lo + 2] or [0.0]
        self.precision = resolve_precision(precision)
        H, W = self.cfg.grid
        for name, shape, kind in pgrm_schema(self.cfg):
            if kind == "param":
                self.attach(name, _initial_value(name, shape, self.cfg), "param")
            elif "relative_position_index" in name:
                g = int(name.rsplit("_", 1)[-1])
                self.attach(name, torch.from_numpy(relative_position_index(self.cfg.window_size[g])), "buffer")
            elif "attn_mask" in name:
                g = int(name.rsplit("_", 1)[-1])
                blk = int(name.split(".")[3])
                ws_eff, sh_eff = self.cfg.effective_windows(blk)
                self.attach(name, torch.from_numpy(shift_mask(H, W, ws_eff[g], sh_eff[g])), "buffer")
            else:
                raise AssertionError(name)
        self._prepared = PreparedWeights()
        _lib.load()   # fail at construction, not at first forward, if the CUDA library is absent

    # ------------------------------------------------------------------------------------------
    def _ptr(self, name: str) -> int:
        t = self.fetch(name)
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"dpmn_b200.PGRM: parameter {name} must be a contiguous fp32 CUDA tensor "
                               f"(got {t.device}, {t.dtype}); there is no CPU path")
        return t.data_ptr()

    def _stochastic(self) -> bool:
        """train() with a non-zero Dropout / DropPath rate: the forward then draws masks (fp32 training sequence)."""
        return self.training and (self.drop_rate > 0 or self.attn_drop_rate > 0 or max(self.drop_path) > 0)

    def _new_seed(self) -> int:
        """Seed of one forward's Dropout / DropPath masks: torch's CPU generator (torch.manual_seed reproduces it) mixed with
        `_seed_salt` -- the trainer sets it to the data-parallel rank, so ranks that seeded torch identically (to build
        identical replicas) still draw different masks for their different shards."""
        base = int(torch.randint(0, 2 ** 62, (1,)).item())
        salt = int(getattr(self, "_seed_salt", 0))
        return (base ^ ((salt * 0x9E3779B97F4A7C15) & (2 ** 62 - 1))) if salt else base

    def _descriptor(self, B: int, q_chans: int, n_mix: int, seed=None) -> _lib.PgrmDesc:
        cfg = self.cfg
        d = _lib.PgrmDesc()
        if seed is not None:    # pgrm.py:24,180,310,494 -- masks are a function of (seed, site, index), see the header
            d.drop_rate, d.attn_drop_rate, d.seed = self.drop_rate, self.attn_drop_rate, seed
            for b, r in enumerate(self.drop_path[:_lib.MAX_BLOCKS]):
                d.drop_path_rate[b] = r
        d.batch, d.img_h, d.img_w, d.patch = B, cfg.img_size[0], cfg.img_size[1], cfg.patch_size
        d.q_chans, d.embed_dim, d.num_heads, d.n_groups = q_chans, cfg.embed_dim, cfg.num_heads, cfg.groups
        for g, ws in enumerate(cfg.window_size):
            d.window[g] = ws
        d.mlp_hidden, d.hidden_size = cfg.mlp_hidden, cfg.hidden_size
        d.precision = _lib.PREC[self.precision]
        d.n_mix = n_mix
        if q_chans == 2:
            if cfg.mode:
                raise RuntimeError("PGRM(mode=True) has no prior_fusion conv: x_q must have 3 channels (pgrm.py:470-471)")
            d.prior_fusion_w, d.prior_fusion_b = self._ptr("prior_fusion.weight"), self._ptr("prior_fusion.bias")
        d.pe_w, d.pe_b = self._ptr("patch_embed.proj.weight"), self._ptr("patch_embed.proj.bias")
        d.pe_norm_w, d.pe_norm_b = self._ptr("patch_embed.norm.weight"), self._ptr("patch_embed.norm.bias")
        for b in range(cfg.depth):
            pre = f"layers.0.blocks.{b}."
            bw = d.blocks[b]
            for field, key in _BLOCK_FIELDS:
                setattr(bw, field + "_w", self._ptr(pre + key + ".weight"))
                setattr(bw, field + "_b", self._ptr(pre + key + ".bias"))
            for g in range(cfg.groups):
                bw.rpb_table[g] = self._ptr(pre + f"attn.relative_position_bias_table_{g}")
        d.head0_w, d.head0_b = self._ptr("conv_before_upsample.0.weight"), self._ptr("conv_before_upsample.0.bias")
        d.head1_w, d.head1_b = self._ptr("conv_before_upsample.1.weight"), self._ptr("conv_before_upsample.1.bias")
        for i in range(n_mix):
            d.mix_weight[i] = self._ptr(f"weight_list_{i}")
        return d

    @staticmethod
    def _image_arg(t: torch.Tensor, name: str):
        """Accept dense NCHW or a channel-slice view (dense within an image, arbitrary batch stride)."""
        if not t.is_cuda:
            raise RuntimeError(f"dpmn_b200.PGRM: {name} is on {t.device}; the hot path only exists on CUDA")
        if t.dtype != torch.float32:
            raise RuntimeError(f"dpmn_b200.PGRM: {name} must be fp32 (the reference's dtype), got {t.dtype}")
        _, c, h, w = t.shape
        if t.stride(3) == 1 and t.stride(2) == w and t.stride(1) == h * w and (t.shape[0] == 1 or t.stride(0) >= c * h * w):
            return t, (t.stride(0) if t.shape[0] > 1 else c * h * w)
        t = t.contiguous()
        return t, c * h * w

    def forward(self, x_q: torch.Tensor, x_kv: torch.Tensor, residual_list: Sequence[torch.Tensor]):
        residual_list = list(residual_list)
        params = [p for _, p in self.named_parameters()]
        if torch.is_grad_enabled() and (x_kv.requires_grad or any(r.requires_grad for r in residual_list)
                                        or any(p.requires_grad for p in params)):
            return _PGRMFunction.apply(self, x_q, x_kv, len(residual_list), *residual_list, *params)
        return self._run(x_q, x_kv, residual_list, probe=False)

    def _backward(self, x_q, x_kv, residuals, d_out, need_xkv=True, need_res=None, seed=None, fwd_ws=None):
        """d_out (B, hs, H, W) -> (d x_kv | None, [d residual_i | None], [d param | None in named_parameters order]).
        `seed`: the seed of the forward's Dropout / DropPath masks (None = none were drawn)."""
        lib = _lib.load()
        cfg = self.cfg
        B = x_q.shape[0]
        n_mix = max(1, len(residuals))
        need_res = list(need_res) if need_res is not None else [True] * len(residuals)
        x_q, q_bs = self._image_arg(x_q, "x_q")
        x_kv, kv_bs = self._image_arg(x_kv, "x_kv")
        d = self._descriptor(B, x_q.shape[1], n_mix, seed=seed)
        d.x_q_batch_stride, d.x_kv_batch_stride = q_bs, kv_bs
        keep = []
        for i in range(1, n_mix):
            r, bs = self._image_arg(residuals[i], f"residual_list[{i}]")
            keep.append(r)
            d.mix_input[i] = r.data_ptr()
            d.mix_input_batch_stride[i] = bs
        dev = x_kv.device
        d_out = d_out.contiguous().float()
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        sizes = [p.numel() for p in params]
        sink = grad_sink_views(self, names, params)
        with torch.cuda.device(dev):
            if sink is not None:
                views = sink          # the caller's gradient bucket: dpmn_pgrm_backward accumulates straight into it
            else:
                flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)   # one flat bucket, views per parameter
                views, off = {}, 0
                for n, p, sz in zip(names, params, sizes):
                    views[n] = flat[off: off + sz].view_as(p)
                    off += sz
            g = _lib.PgrmGrads()
            used = set()

            def gp(name):
                used.add(name)
                return views[name].data_ptr()
            if x_q.shape[1] == 2:
                g.prior_fusion_w, g.prior_fusion_b = gp("prior_fusion.weight"), gp("prior_fusion.bias")
            g.pe_w, g.pe_b = gp("patch_embed.proj.weight"), gp("patch_embed.proj.bias")
            g.pe_norm_w, g.pe_norm_b = gp("patch_embed.norm.weight"), gp("patch_embed.norm.bias")
            for b in range(cfg.depth):
                pre = f"layers.0.blocks.{b}."
                bg = g.blocks[b]
                for field, key in _BLOCK_FIELDS:
                    setattr(bg, field + "_w", gp(pre + key + ".weight"))
                    setattr(bg, field + "_b", gp(pre + key + ".bias"))
                for gi in range(cfg.groups):
                    bg.rpb_table[gi] = gp(pre + f"attn.relative_position_bias_table_{gi}")
            g.head0_w, g.head0_b = gp("conv_before_upsample.0.weight"), gp("conv_before_upsample.0.bias")
            g.head1_w, g.head1_b = gp("conv_before_upsample.1.weight"), gp("conv_before_upsample.1.bias")
            for i in range(n_mix):
                g.mix_weight[i] = gp(f"weight_list_{i}")
            img = (B, cfg.hidden_size, cfg.img_size[0], cfg.img_size[1])
            d_xkv = torch.empty((B, 3, cfg.img_size[0], cfg.img_size[1]), dtype=torch.float32, device=dev) if need_xkv else None
            if d_xkv is not None:
                g.x_kv = d_xkv.data_ptr()
            d_res = [None] * len(residuals)
            for i in range(1, n_mix):    # residual_list[0] never enters the output (pgrm.py:563)
                if need_res[i]:
                    d_res[i] = torch.empty(img, dtype=torch.float32, device=dev)
                    g.mix_input[i] = d_res[i].data_ptr()
            nbytes = lib.dpmn_pgrm_backward_workspace_bytes(C.byref(d))
            if nbytes == 0:
                raise RuntimeError("dpmn_pgrm_backward_workspace_bytes: configuration rejected")
            if fwd_ws is not None and fwd_ws.numel() >= nbytes:
                ws = fwd_ws
                d.flags = _lib.PGRM_WORKSPACE_HOLDS_FORWARD
            else:
                ws = workspace(dev, nbytes)
            rc = lib.dpmn_pgrm_backward(C.byref(d), x_q.data_ptr(), x_kv.data_ptr(), d_out.data_ptr(), C.byref(g),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "dpmn_pgrm_backward")
        # parameters the output does not depend on get no gradient, as in the reference (unused weight_list_i,
        # prior_fusion when x_q already has 3 channels)
        if sink is not None:
            d_params = [None] * len(params)       # already accumulated in place
        else:
            d_params = [views[n] if (n in used and p.requires_grad) else None for n, p in zip(names, params)]
        hook = getattr(self, "_after_backward", None)
        if hook is not None:
            hook(self)
        return d_xkv, d_res, d_params

    def forward_probe(self, x_q, x_kv, residual_list):
        """forward + the per-block tensors the parity tests compare: (out, attn_core[2], block_out[2])."""
        return self._run(x_q, x_kv, residual_list, probe=True)

    def _run(self, x_q, x_kv, residual_list, probe: bool, seed=None, keep_ws=None):
        if seed is None and self._stochastic():
            seed = self._new_seed()             # train-mode forward outside autograd (torch.no_grad())
        lib = _lib.load()
        cfg = self.cfg
        if x_q.dim() != 4 or x_kv.dim() != 4 or x_q.shape[0] != x_kv.shape[0]:
            raise ValueError("PGRM.forward: x_q (B,2|3,H,W) and x_kv (B,3,H,W) expected")
        if tuple(x_kv.shape[1:]) != (3, cfg.img_size[0], cfg.img_size[1]) or tuple(x_q.shape[2:]) != tuple(cfg.img_size):
            # pgrm.py:421: "Input image size doesn't match model"
            raise AssertionError(f"Input image size ({tuple(x_kv.shape)}) doesn't match model {cfg.img_size}")
        B = x_q.shape[0]
        n_mix = max(1, len(residual_list))
        if n_mix > self.iter + 1:
            raise AttributeError(f"PGRM(iter={self.iter}) has no weight_list_{n_mix - 1} (pgrm.py:564)")
        x_q, q_bs = self._image_arg(x_q, "x_q")
        x_kv, kv_bs = self._image_arg(x_kv, "x_kv")
        d = self._descriptor(B, x_q.shape[1], n_mix, seed=seed)
        d.x_q_batch_stride, d.x_kv_batch_stride = q_bs, kv_bs
        keep = []
        for i in range(1, n_mix):   # residual_list[0] is skipped by the reference (pgrm.py:563)
            r, bs = self._image_arg(residual_list[i], f"residual_list[{i}]")
            if tuple(r.shape) != (B, cfg.hidden_size, cfg.img_size[0], cfg.img_size[1]):
                raise ValueError(f"residual_list[{i}] has shape {tuple(r.shape)}")
            keep.append(r)
            d.mix_input[i] = r.data_ptr()
            d.mix_input_batch_stride[i] = bs
        dev = x_kv.device
        with torch.cuda.device(dev):
            prep_key = self._prepared.attach(self, d, lib.dpmn_pgrm_prepared_bytes(C.byref(d)), dev)
            nbytes = lib.dpmn_pgrm_workspace_bytes(C.byref(d))
            if nbytes == 0:
                raise RuntimeError("dpmn_pgrm_workspace_bytes: configuration rejected (see DPMN_E_UNSUPPORTED rules "
                                   "in include/dpmn_b200.h)")
            if keep_ws is not None:
                ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
                keep_ws.append(ws)
            else:
                ws = workspace(dev, nbytes)
            out = torch.empty((B, cfg.hidden_size, cfg.img_size[0], cfg.img_size[1]), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            if not probe:
                rc = lib.dpmn_pgrm_forward(C.byref(d), x_q.data_ptr(), x_kv.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                           ws.numel(), stream)
                _lib.check(rc, "dpmn_pgrm_forward")
                if seed is None:     # the stochastic training sequence never stages `prepared`: it must not validate it
                    self._prepared.key = prep_key
                return out
            L, Cc = cfg.tokens, cfg.embed_dim
            cores = [torch.empty((B, L, Cc), dtype=torch.float32, device=dev) for _ in range(2)]
            blocks = [torch.empty((B, L, Cc), dtype=torch.float32, device=dev) for _ in range(2)]
            a = (C.c_void_p * 2)(*[t.data_ptr() for t in cores])
            b = (C.c_void_p * 2)(*[t.data_ptr() for t in blocks])
            rc = lib.dpmn_pgrm_forward_probe(C.byref(d), x_q.data_ptr(), x_kv.data_ptr(), out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), stream, C.byref(a), C.byref(b))
            _lib.check(rc, "dpmn_pgrm_forward_probe")
            self._prepared.key = prep_key
            return out, cores, blocks


def window_attention(q: torch.Tensor, kv: torch.Tensor, tables: List[torch.Tensor], grid, num_heads: int,
                     windows: Sequence[int], shifts: Sequence[int]) -> torch.Tensor:
    """Stand-alone windowed attention core (pgrm.py:197-268) on projected q (B,L,C), kv (B,L,2C).
    dtype fp32 / fp16 / bf16 selects the storage precision; output rows are window-major (quirk 1)."""
    lib = _lib.load()
    if not (q.is_cuda and kv.is_cuda and q.is_contiguous() and kv.is_contiguous()):
        raise RuntimeError("window_attention: contiguous CUDA tensors required")
    prec = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[q.dtype]
    B, L, Cc = q.shape
    out = torch.empty_like(q)
    G = len(windows)
    tabs = (C.c_void_p * _lib.MAX_GROUPS)(*[t.data_ptr() for t in tables])
    wv = (C.c_int32 * _lib.MAX_GROUPS)(*windows)
    sv = (C.c_int32 * _lib.MAX_GROUPS)(*shifts)
    with torch.cuda.device(q.device):
        ws = workspace(q.device, 256)
        rc = lib.dpmn_window_attn_forward(q.data_ptr(), kv.data_ptr(), out.data_ptr(), C.byref(tabs), B, grid[0], grid[1],
                                          Cc, num_heads, G, C.byref(wv), C.byref(sv), prec, ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream(q.device).cuda_stream)
    _lib.check(rc, "dpmn_window_attn_forward")
    return out


def window_order(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """window-major row -> original token index of one group (roll + window_partition, pgrm.py:209-221,43-52)."""
    p = torch.arange(H * W)
    w_idx, n = p // (ws * ws), p % (ws * ws)
    nWw = W // ws
    hp = (w_idx // nWw) * ws + n // ws
    wp = (w_idx % nWw) * ws + n % ws
    return ((hp + shift) % H) * W + (wp + shift) % W


def to_window_major(x: torch.Tensor, grid, windows: Sequence[int], shifts: Sequence[int]) -> torch.Tensor:
    """(B, L, C) token order -> [G][B*L][C/G] window-major per group: the layout the projection epilogue writes and
    `window_attention_windowed` reads.  Host-side helper for tests / the stand-alone sweep (a torch gather)."""
    B, L, C = x.shape
    G = len(windows)
    cg = C // G
    parts = []
    for g, (ws, sh) in enumerate(zip(windows, shifts)):
        order = window_order(grid[0], grid[1], ws, sh).to(x.device)
        parts.append(x[:, order, g * cg:(g + 1) * cg].reshape(B * L, cg))
    return torch.stack(parts).contiguous()


def window_attention_windowed(qw: torch.Tensor, kw: torch.Tensor, vw: torch.Tensor, tables: List[torch.Tensor], batch: int,
                              grid, num_heads: int, windows: Sequence[int], shifts: Sequence[int],
                              drop: Optional[Tuple[float, int, int]] = None) -> torch.Tensor:
    """The tcgen05 window-attention kernel on window-major fp16 / bf16 operands [G][B*L][C/G] -> (B, L, C).
    drop = (attn_drop rate, seed, site): train-mode attn_drop (pgrm.py:248) with the library's counter-hash masks."""
    lib = _lib.load()
    if not (qw.is_cuda and qw.is_contiguous() and kw.is_contiguous() and vw.is_contiguous()):
        raise RuntimeError("window_attention_windowed: contiguous CUDA tensors required")
    prec = {torch.float16: 1, torch.bfloat16: 2}[qw.dtype]
    G, rows, cg = qw.shape
    L = grid[0] * grid[1]
    out = torch.empty((batch, L, G * cg), dtype=qw.dtype, device=qw.device)
    tabs = (C.c_void_p * _lib.MAX_GROUPS)(*[t.data_ptr() for t in tables])
    wv = (C.c_int32 * _lib.MAX_GROUPS)(*windows)
    sv = (C.c_int32 * _lib.MAX_GROUPS)(*shifts)
    with torch.cuda.device(qw.device):
        p_drop, seed, site = drop if drop is not None else (0.0, 0, 0)
        rc = lib.dpmn_window_attn_forward_windowed_train(qw.data_ptr(), kw.data_ptr(), vw.data_ptr(), out.data_ptr(),
                                                         C.byref(tabs), batch, grid[0], grid[1], G * cg, num_heads, G,
                                                         C.byref(wv), C.byref(sv), prec, float(p_drop), int(seed), int(site),
                                                         torch.cuda.current_stream(qw.device).cuda_stream)
    _lib.check(rc, "dpmn_window_attn_forward_windowed_train")
    return out


def window_attention_windowed_backward(qw: torch.Tensor, kw: torch.Tensor, vw: torch.Tensor, d_out: torch.Tensor,
                                       tables: List[torch.Tensor], batch: int, grid, num_heads: int, windows: Sequence[int],
                                       shifts: Sequence[int], drop: Optional[Tuple[float, int, int]] = None):
    """Backward of `window_attention_windowed` on tcgen05: d_out (B, L, C) 16-bit in the output's window-major row order ->
    (dq (B, L, C), dkv (B, L, 2C)) fp32 in token order, and the list of relative-position-table gradients."""
    lib = _lib.load()
    for t in (qw, kw, vw, d_out):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("window_attention_windowed_backward: contiguous CUDA tensors required")
    prec = {torch.float16: 1, torch.bfloat16: 2}[qw.dtype]
    G, rows, cg = qw.shape
    L = grid[0] * grid[1]
    dq = torch.empty((batch, L, G * cg), dtype=torch.float32, device=qw.device)
    dkv = torch.empty((batch, L, 2 * G * cg), dtype=torch.float32, device=qw.device)
    dtabs = [torch.zeros_like(t) for t in tables]
    tabs = (C.c_void_p * _lib.MAX_GROUPS)(*[t.data_ptr() for t in tables])
    dts = (C.c_void_p * _lib.MAX_GROUPS)(*[t.data_ptr() for t in dtabs])
    wv = (C.c_int32 * _lib.MAX_GROUPS)(*windows)
    sv = (C.c_int32 * _lib.MAX_GROUPS)(*shifts)
    p_drop, seed, site = drop if drop is not None else (0.0, 0, 0)
    with torch.cuda.device(qw.device):
        rc = lib.dpmn_window_attn_backward_windowed(qw.data_ptr(), kw.data_ptr(), vw.data_ptr(), d_out.data_ptr(), dq.data_ptr(),
                                                    dkv.data_ptr(), C.byref(tabs), C.byref(dts), batch, grid[0], grid[1], G * cg,
                                                    num_heads, G, C.byref(wv), C.byref(sv), prec, float(p_drop), int(seed), int(site),
                                                    torch.cuda.current_stream(qw.device).cuda_stream)
    _lib.check(rc, "dpmn_window_attn_backward_windowed")
    return dq, dkv, dtabs

