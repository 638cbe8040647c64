"""Drop-in `DistillModule`: the reference's constructor, forward signature, return value
`(loss, feature_cat)` and state_dict schema (/root/reference/model/distill_module.py:4-31) over the libdpmn_b200 C-ABI
(dpmn_distill_forward / dpmn_distill_backward, csrc/distill.cu).  No PyTorch/CPU compute path: the nn.Conv2d /
nn.BatchNorm2d children below only own the parameters and buffers (same names, shapes and initialisers as the
reference's, which uses torch defaults); their forwards are never called."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


def _dense_or_stride(t: torch.Tensor):
    """(tensor, batch stride in elements) for a (B, 3, H, W) fp32 tensor whose images are dense."""
    B, Ch, H, W = t.shape
    if t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W:
        return t, (t.stride(0) if B > 1 else Ch * H * W)
    return t.contiguous(), Ch * H * W


class _DistillFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x_deep, x_shallow, *params):
        ctx.set_materialize_grads(False)
        loss, feature, ws, desc_args = module._forward_impl(x_deep, x_shallow)
        ctx.module, ctx.ws, ctx.desc_args = module, ws, desc_args
        ctx.save_for_backward(x_deep, x_shallow)
        return loss, feature

    @staticmethod
    def backward(ctx, g_loss, g_feature):
        x_deep, x_shallow = ctx.saved_tensors
        out = ctx.module._backward(x_deep, x_shallow, g_loss, g_feature, ctx.ws, ctx.desc_args,
                                   ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        ctx.ws = None
        d_deep, d_shallow, d_params = out
        return (None, d_deep, d_shallow, *d_params)


class DistillModule(nn.Module):
    """distill_module.py:5-16.  forward(x_deep, x_shallow) -> (L1 loss (0-d), feature_cat (B, 3, H, W))."""

    def __init__(self):
        super().__init__()
        self.conv_cat_feature = nn.Conv2d(6, 3, 3, 1, 1)
        self.bn_1 = nn.BatchNorm2d(3)
        self.act_1 = nn.ReLU(True)
        self.conv_feature = nn.Conv2d(3, 3, 3, 1, 1)
        self.bn_2 = nn.BatchNorm2d(3)
        self.act_2 = nn.ReLU(True)
        self.loss = nn.L1Loss()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        _lib.load()

    # ---- descriptor -------------------------------------------------------------------------------------------------
    def _ptr(self, t: torch.Tensor, name: str) -> int:
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"dpmn_b200 DistillModule: {name} must be a contiguous fp32 CUDA tensor (got {t.device}, "
                               f"{t.dtype}); there is no CPU path")
        return t.data_ptr()

    def _descriptor(self, B, H, W, training, update, deep_bs, shallow_bs) -> _lib.DistillDesc:
        d = _lib.DistillDesc()
        d.batch, d.img_h, d.img_w = B, H, W
        d.training, d.update_running_stats, d.flags = int(training), int(update), 0
        d.bn_eps, d.bn_momentum = float(self.bn_1.eps), float(self.bn_1.momentum)
        d.deep_batch_stride, d.shallow_batch_stride = deep_bs, shallow_bs
        d.conv_cat_w = self._ptr(self.conv_cat_feature.weight, "conv_cat_feature.weight")
        d.conv_cat_b = self._ptr(self.conv_cat_feature.bias, "conv_cat_feature.bias")
        d.conv_w = self._ptr(self.conv_feature.weight, "conv_feature.weight")
        d.conv_b = self._ptr(self.conv_feature.bias, "conv_feature.bias")
        for dst, bn, nm in ((d.bn_1, self.bn_1, "bn_1"), (d.bn_2, self.bn_2, "bn_2")):
            dst.w, dst.b = self._ptr(bn.weight, nm + ".weight"), self._ptr(bn.bias, nm + ".bias")
            dst.running_mean = self._ptr(bn.running_mean, nm + ".running_mean")
            dst.running_var = self._ptr(bn.running_var, nm + ".running_var")
        return d

    @staticmethod
    def _check(x_deep, x_shallow):
        for n, t in (("x_deep", x_deep), ("x_shallow", x_shallow)):
            if not t.is_cuda:
                raise RuntimeError(f"dpmn_b200 DistillModule: {n} is on {t.device}; this module only exists on CUDA")
            if t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != 3:
                raise ValueError(f"dpmn_b200 DistillModule: {n} must be fp32 (B,3,H,W), got {t.dtype} {tuple(t.shape)}")
        if x_deep.shape != x_shallow.shape:
            raise ValueError("DistillModule.forward: x_deep and x_shallow must have the same shape")

    # ---- forward / backward -----------------------------------------------------------------------------------------
    def forward(self, x_deep: torch.Tensor, x_shallow: torch.Tensor):
        self._check(x_deep, x_shallow)
        params = list(self.parameters())
        if torch.is_grad_enabled() and (x_deep.requires_grad or x_shallow.requires_grad or
                                        any(p.requires_grad for p in params)):
            return _DistillFunction.apply(self, x_deep, x_shallow, *params)
        loss, feature, _, _ = self._forward_impl(x_deep, x_shallow)
        return loss, feature

    def _forward_impl(self, x_deep, x_shallow):
        lib = _lib.load()
        B, _, H, W = x_deep.shape
        if self.training and B * H * W == 1:
            raise ValueError("Expected more than 1 value per channel when training")
        xd, d_bs = _dense_or_stride(x_deep.detach())
        xs, s_bs = _dense_or_stride(x_shallow.detach())
        dev = x_deep.device
        training = bool(self.training)
        with torch.cuda.device(dev):
            d = self._descriptor(B, H, W, training, training, d_bs, s_bs)
            ws = torch.empty(int(lib.dpmn_distill_workspace_bytes(C.byref(d))), dtype=torch.uint8, device=dev)
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            feature = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
            rc = lib.dpmn_distill_forward(C.byref(d), xd.data_ptr(), xs.data_ptr(), loss.data_ptr(), feature.data_ptr(),
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "dpmn_distill_forward")
        if training:
            self.bn_1.num_batches_tracked += 1
            self.bn_2.num_batches_tracked += 1
        return loss, feature, ws, (training, xd, d_bs, xs, s_bs)

    def _backward(self, x_deep, x_shallow, g_loss, g_feature, ws, desc_args, need_deep=True, need_shallow=True):
        lib = _lib.load()
        training, xd, d_bs, xs, s_bs = desc_args
        B, _, H, W = x_deep.shape
        dev = x_deep.device
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        with torch.cuda.device(dev):
            d = self._descriptor(B, H, W, training, False, d_bs, s_bs)
            d.flags = _lib.DISTILL_WORKSPACE_HOLDS_FORWARD
            from .pgrm import grad_sink_views
            sink = grad_sink_views(self, names, params)
            if sink is not None:
                views = sink          # accumulate straight into the caller's gradient bucket
            else:
                flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
                views, off = {}, 0
                for n, p in zip(names, params):
                    views[n] = flat[off: off + p.numel()].view_as(p)
                    off += p.numel()
            g = _lib.DistillGrads()
            g.conv_cat_w, g.conv_cat_b = views["conv_cat_feature.weight"].data_ptr(), views["conv_cat_feature.bias"].data_ptr()
            g.conv_w, g.conv_b = views["conv_feature.weight"].data_ptr(), views["conv_feature.bias"].data_ptr()
            g.bn_1.w, g.bn_1.b = views["bn_1.weight"].data_ptr(), views["bn_1.bias"].data_ptr()
            g.bn_2.w, g.bn_2.b = views["bn_2.weight"].data_ptr(), views["bn_2.bias"].data_ptr()
            d_deep = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) if need_deep else None
            d_shallow = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev) if need_shallow else None
            g.x_deep = d_deep.data_ptr() if d_deep is not None else None
            g.x_shallow = d_shallow.data_ptr() if d_shallow is not None else None
            gl = g_loss.detach().reshape(()).float().contiguous() if g_loss is not None else None
            gf = g_feature.detach().float().contiguous() if g_feature is not None else None
            rc = lib.dpmn_distill_backward(C.byref(d), xd.data_ptr(), xs.data_ptr(),
                                           gl.data_ptr() if gl is not None else None,
                                           gf.data_ptr() if gf is not None else None, C.byref(g), ws.data_ptr(), ws.numel(),
                                           torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "dpmn_distill_backward")
        if sink is not None:
            return d_deep, d_shallow, [None] * len(params)
        return d_deep, d_shallow, [views[n] if p.requires_grad else None for n, p in zip(names, params)]
