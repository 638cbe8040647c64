"""Drop-in `ComplementationModulationModule`: the reference's constructor, forward signature and state_dict
schema (/root/reference/model/cmm.py:80-161) over the libdpmn_b200 C-ABI.  No PyTorch/CPU compute path."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from .pgrm import ParamTree, PreparedWeights, grad_sink_views, workspace
from .schema import cmm_schema


def _initial_value(name: str, shape) -> torch.Tensor:
    """torch defaults, which is all the reference uses for this module (no custom init in cmm.py):
    kaiming_uniform(a=sqrt(5)) conv/linear weights, U(+-1/sqrt(fan_in)) biases, BN (1, 0, mean 0, var 1)."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    t = torch.empty(tuple(shape), dtype=torch.float32)
    if leaf in ("running_mean",):
        return t.zero_()
    if leaf in ("running_var",):
        return t.fill_(1.0)
    if len(shape) >= 2:
        return nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    return t   # 1-D: resolved by the caller (needs the sibling weight's fan-in)


class _CMMFunction(torch.autograd.Function):
    """Autograd node of CMM.forward: backward = dpmn_cmm_backward (recomputes the fp32 forward internally)."""

    @staticmethod
    def forward(ctx, module, x1, x2, *params):
        with torch.no_grad():
            # fp32 forward: run it inside a private buffer laid out as the backward's workspace, so that the backward
            # finds every activation in place and skips its recompute (DPMN_CMM_WORKSPACE_HOLDS_FORWARD)
            keep = module.precision == "fp32" or module.training   # the fp32-structured forward (see include/dpmn_b200.h)
            res = module._forward_impl(x1, x2, keep_workspace=keep)
            out, ws = res if keep else (res, None)
        ctx.module = module
        ctx.training = module.training
        ctx.fwd_ws = ws
        ctx.save_for_backward(x1, x2)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x1, x2 = ctx.saved_tensors
        d_x1, d_x2, d_params = ctx.module._backward(x1, x2, d_out, ctx.training, ctx.needs_input_grad[1],
                                                    ctx.needs_input_grad[2], fwd_ws=ctx.fwd_ws)
        ctx.fwd_ws = None
        return (None, d_x1, d_x2, *d_params)


class ComplementationModulationModule(ParamTree):
    """cmm.py:80-81 signature.  Extra keyword `precision` as in `PGRM`."""

    def __init__(self, c_img=3, norm='batch', act_en='leaky_relu', act_de='relu', cnum=64, precision=None):
        super().__init__()
        if norm != 'batch' or act_en != 'leaky_relu' or act_de != 'relu':
            raise NotImplementedError("dpmn_b200 CMM implements the configuration DPMN instantiates "
                                      "(norm='batch', act_en='leaky_relu', act_de='relu'; super_resolution.py:72)")
        from .pgrm import resolve_precision
        self.c_img, self.cnum, self.precision = int(c_img), int(cnum), resolve_precision(precision)
        schema = cmm_schema(self.c_img, self.cnum)
        shapes = {n: s for n, s, _ in schema}
        for name, shape, kind in schema:
            leaf = name.rsplit(".", 1)[-1]
            v = _initial_value(name, shape)
            if len(shape) == 1 and kind == "param":
                stem = name.rsplit(".", 1)[0]
                wshape = shapes[stem + ".weight"]
                if len(wshape) == 1:          # BatchNorm affine
                    v = v.fill_(1.0) if leaf == "weight" else v.zero_()
                else:                         # conv / linear bias: fan_in of the sibling weight
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    bound = 1.0 / math.sqrt(fan_in)
                    v = v.uniform_(-bound, bound)
            self.attach(name, v, kind)
        self._prepared = PreparedWeights()
        _lib.load()

    def _ptr(self, name: str) -> int:
        t = self.fetch(name)
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"dpmn_b200 CMM: {name} must be a contiguous fp32 CUDA tensor (got {t.device}, "
                               f"{t.dtype}); there is no CPU path")
        return t.data_ptr()

    def _bn(self, dst: _lib.Bn, stem: str):
        dst.w, dst.b = self._ptr(stem + ".weight"), self._ptr(stem + ".bias")
        dst.running_mean, dst.running_var = self._ptr(stem + ".running_mean"), self._ptr(stem + ".running_var")

    def _stage(self, dst: _lib.CmmStage, stem: str):
        dst.conv_a_w, dst.conv_a_b = self._ptr(stem + "1.weight"), self._ptr(stem + "1.bias")
        self._bn(dst.bn_a, stem + "2")
        dst.conv_b_w, dst.conv_b_b = self._ptr(stem + "4.weight"), self._ptr(stem + "4.bias")
        self._bn(dst.bn_b, stem + "5")

    def _descriptor(self, B, H, W) -> _lib.CmmDesc:
        d = _lib.CmmDesc()
        d.batch, d.img_h, d.img_w, d.c_img, d.cnum = B, H, W, self.c_img, self.cnum
        d.precision = _lib.PREC[self.precision]
        d.training = int(self.training)
        d.update_running_stats = int(self.training)
        for br in (0, 1):
            d.en1_w[br], d.en1_b[br] = self._ptr(f"en_1_{br + 1}.weight"), self._ptr(f"en_1_{br + 1}.bias")
            for l, lvl in enumerate((2, 3, 4, 5)):
                self._stage(d.enc[br][l], f"en_{lvl}_{br + 1}.encode.")
            d.en6_w[br], d.en6_b[br] = self._ptr(f"en_6_{br + 1}.1.weight"), self._ptr(f"en_6_{br + 1}.1.bias")
        d.fc1_w, d.fc1_b = self._ptr("fc_1.weight"), self._ptr("fc_1.bias")
        d.fc2_w, d.fc2_b = self._ptr("fc_2.weight"), self._ptr("fc_2.bias")
        d.de6_w, d.de6_b = self._ptr("de_6.1.weight"), self._ptr("de_6.1.bias")
        self._bn(d.de6_bn, "de_6.2")
        for i, lvl in enumerate((5, 4, 3, 2)):
            self._stage(d.dec[i], f"de_{lvl}.decode.")
        d.de1_w, d.de1_b = self._ptr("de_1.1.weight"), self._ptr("de_1.1.bias")
        return d

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, blend_with: torch.Tensor = None, alpha: float = 0.5) -> torch.Tensor:
        """cmm.py:120 signature.  Extra keywords (eval / test call sites, super_resolution.py:449,705): with `blend_with`
        = images_lr_psn[:, :3] the result is alpha * CMM(x1, x2) + (1 - alpha) * blend_with, the blend fused into the kernel
        that writes the output."""
        params = [p for _, p in self.named_parameters()]
        if torch.is_grad_enabled() and (x1.requires_grad or x2.requires_grad or any(p.requires_grad for p in params)):
            y = _CMMFunction.apply(self, x1, x2, *params)
            # under autograd the blend stays a torch expression (the reference's own line); the training loss never uses it
            return y if blend_with is None else alpha * y + (1 - alpha) * blend_with
        return self._forward_impl(x1, x2, blend_with=blend_with, alpha=alpha)

    def _backward(self, x1, x2, d_out, training, need_x1=True, need_x2=True, fwd_ws=None):
        """d_out (B, c_img, H, W) -> (d x1 | None, d x2 | None, [d param in named_parameters order])."""
        lib = _lib.load()
        x1, x2 = x1.contiguous(), x2.contiguous()
        B, _, H, W = x1.shape
        d = self._descriptor(B, H, W)
        d.training = int(training)
        d.update_running_stats = 0
        dev = x1.device
        d_out = d_out.contiguous().float()
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        sink = grad_sink_views(self, names, params)
        with torch.cuda.device(dev):
            if sink is not None:
                views = sink          # accumulate straight into the caller's gradient bucket
            else:
                flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
                views, off = {}, 0
                for n, p in zip(names, params):
                    views[n] = flat[off: off + p.numel()].view_as(p)
                    off += p.numel()

            def gp(name):
                return views[name].data_ptr()

            def stage(dst, stem):
                dst.conv_a_w, dst.conv_a_b = gp(stem + "1.weight"), gp(stem + "1.bias")
                dst.bn_a.w, dst.bn_a.b = gp(stem + "2.weight"), gp(stem + "2.bias")
                dst.conv_b_w, dst.conv_b_b = gp(stem + "4.weight"), gp(stem + "4.bias")
                dst.bn_b.w, dst.bn_b.b = gp(stem + "5.weight"), gp(stem + "5.bias")
            g = _lib.CmmGrads()
            for br in (0, 1):
                g.en1_w[br], g.en1_b[br] = gp(f"en_1_{br + 1}.weight"), gp(f"en_1_{br + 1}.bias")
                for l, lvl in enumerate((2, 3, 4, 5)):
                    stage(g.enc[br][l], f"en_{lvl}_{br + 1}.encode.")
                g.en6_w[br], g.en6_b[br] = gp(f"en_6_{br + 1}.1.weight"), gp(f"en_6_{br + 1}.1.bias")
            g.fc1_w, g.fc1_b, g.fc2_w, g.fc2_b = gp("fc_1.weight"), gp("fc_1.bias"), gp("fc_2.weight"), gp("fc_2.bias")
            g.de6_w, g.de6_b = gp("de_6.1.weight"), gp("de_6.1.bias")
            g.de6_bn.w, g.de6_bn.b = gp("de_6.2.weight"), gp("de_6.2.bias")
            for i, lvl in enumerate((5, 4, 3, 2)):
                stage(g.dec[i], f"de_{lvl}.decode.")
            g.de1_w, g.de1_b = gp("de_1.1.weight"), gp("de_1.1.bias")
            d_x1 = torch.empty_like(x1) if need_x1 else None
            d_x2 = torch.empty_like(x2) if need_x2 else None
            if d_x1 is not None:
                g.x1 = d_x1.data_ptr()
            if d_x2 is not None:
                g.x2 = d_x2.data_ptr()
            nbytes = lib.dpmn_cmm_backward_workspace_bytes(C.byref(d))
            if nbytes == 0:
                raise RuntimeError("dpmn_cmm_backward_workspace_bytes: configuration rejected")
            if fwd_ws is not None and fwd_ws.numel() >= nbytes:
                ws = fwd_ws
                d.flags = _lib.CMM_WORKSPACE_HOLDS_FORWARD
            else:
                ws = workspace(dev, nbytes)
            rc = lib.dpmn_cmm_backward(C.byref(d), x1.data_ptr(), x2.data_ptr(), d_out.data_ptr(), C.byref(g),
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "dpmn_cmm_backward")
        hook = getattr(self, "_after_backward", None)
        if hook is not None:
            hook(self)           # e.g. the trainer starts the all-reduce of this module's bucket segment (dpmn_b200.train)
        if sink is not None:
            return d_x1, d_x2, [None] * len(params)
        return d_x1, d_x2, [views[n] if p.requires_grad else None for n, p in zip(names, params)]

    def _forward_impl(self, x1: torch.Tensor, x2: torch.Tensor, keep_workspace: bool = False, blend_with=None, alpha=0.5):
        lib = _lib.load()
        for n, t in (("x1", x1), ("x2", x2)):
            if not t.is_cuda:
                raise RuntimeError(f"dpmn_b200 CMM: {n} is on {t.device}; the hot path only exists on CUDA")
            if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != self.c_img:
                raise ValueError(f"dpmn_b200 CMM: {n} must be fp32 (B,{self.c_img},H,W), got {t.dtype} {tuple(t.shape)}")
        if x1.shape != x2.shape:
            raise ValueError("CMM.forward: x1 and x2 must have the same shape")
        x1, x2 = x1.contiguous(), x2.contiguous()
        B, _, H, W = x1.shape
        # (no BatchNorm of the CMM ever sees a single value per channel: the 1x1 bottleneck en_6 has none, the smallest
        #  normalised maps are en_5 / de_6 at (H/16, W/16) >= 2x2 -- the reference accepts B = 1 at 32x32, cmm.py:91-93,107-111)
        d = self._descriptor(B, H, W)
        if blend_with is not None:
            if not blend_with.is_cuda or blend_with.dtype != torch.float32 or tuple(blend_with.shape) != tuple(x1.shape):
                raise ValueError(f"dpmn_b200 CMM: blend_with must be an fp32 CUDA tensor of shape {tuple(x1.shape)}")
            if not (blend_with.stride(3) == 1 and blend_with.stride(2) == W and blend_with.stride(1) == H * W):
                blend_with = blend_with.contiguous()
            d.blend_input = blend_with.data_ptr()
            d.blend_input_batch_stride = blend_with.stride(0) if B > 1 else self.c_img * H * W
            d.blend_alpha = float(alpha)
        dev = x1.device
        with torch.cuda.device(dev):
            prep_key = self._prepared.attach(self, d, lib.dpmn_cmm_prepared_bytes(C.byref(d)), dev)
            nbytes = lib.dpmn_cmm_workspace_bytes(C.byref(d))
            if nbytes == 0:
                raise RuntimeError("dpmn_cmm_workspace_bytes: configuration rejected (image sides must be multiples of 32)")
            if keep_workspace:
                ws = torch.empty(int(lib.dpmn_cmm_backward_workspace_bytes(C.byref(d))), dtype=torch.uint8, device=dev)
            else:
                ws = workspace(dev, nbytes)
            out = torch.empty_like(x1)
            rc = lib.dpmn_cmm_forward(C.byref(d), x1.data_ptr(), x2.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                      ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "dpmn_cmm_forward")
        if not self.training:    # only the eval-mode tensor-core path stages `prepared` (cmm_forward_tc)
            self._prepared.key = prep_key
        if self.training:
            for name, buf in self.named_buffers():
                if name.endswith("num_batches_tracked"):
                    buf += 1
        return (out, ws) if keep_workspace else out
