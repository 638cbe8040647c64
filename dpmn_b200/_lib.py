"""ctypes binding of libdpmn_b200.so (include/dpmn_b200.h).

The shared library is the product: there is NO fallback.  If it is missing or does not export the
symbols of the header, importing the compute modules raises.  Structs mirror the header field by field
and are verified against `dpmn_abi_sizeof` at load time.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpmn_b200.so")

MAX_GROUPS, MAX_MIX, MAX_BLOCKS = 4, 8, 2
CMM_WORKSPACE_HOLDS_FORWARD = 1
DISTILL_WORKSPACE_HOLDS_FORWARD = 1
PGRM_WORKSPACE_HOLDS_FORWARD = 1
PREC = {"fp32": 0, "f32": 0, "fp16": 1, "f16": 1, "bf16": 2}
ERRORS = {-1: "DPMN_E_ARG (null pointer / inconsistent sizes)",
          -2: "DPMN_E_UNSUPPORTED (configuration outside this build or that the reference cannot run)",
          -3: "DPMN_E_WORKSPACE (workspace too small)",
          -4: "DPMN_E_DEVICE (not an sm_100 device)"}

fp = C.c_void_p   # device pointers travel as integers


class BlockWeights(C.Structure):
    _fields_ = ([(n, fp) for n in ("norm1_q_w", "norm1_q_b", "norm1_kv_w", "norm1_kv_b")]
                + [("rpb_table", fp * MAX_GROUPS)]
                + [(n, fp) for n in ("q_w", "q_b", "kv_w", "kv_b", "sk_proj_w", "sk_proj_b", "sk_fc1_w", "sk_fc1_b",
                                     "sk_fc2_w", "sk_fc2_b", "sk_head_w", "sk_head_b", "norm2_w", "norm2_b",
                                     "fc1_w", "fc1_b", "fc2_w", "fc2_b", "dw_w", "dw_b", "pw_w", "pw_b")])


class PgrmDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("patch", C.c_int32),
                ("q_chans", C.c_int32), ("embed_dim", C.c_int32), ("num_heads", C.c_int32),
                ("n_groups", C.c_int32), ("window", C.c_int32 * MAX_GROUPS), ("mlp_hidden", C.c_int32),
                ("hidden_size", C.c_int32), ("precision", C.c_int32), ("n_mix", C.c_int32),
                ("x_q_batch_stride", C.c_int64), ("x_kv_batch_stride", C.c_int64),
                ("prior_fusion_w", fp), ("prior_fusion_b", fp), ("pe_w", fp), ("pe_b", fp),
                ("pe_norm_w", fp), ("pe_norm_b", fp),
                ("blocks", BlockWeights * MAX_BLOCKS),
                ("head0_w", fp), ("head0_b", fp), ("head1_w", fp), ("head1_b", fp),
                ("mix_weight", fp * MAX_MIX), ("mix_input", fp * MAX_MIX),
                ("mix_input_batch_stride", C.c_int64 * MAX_MIX),
                ("prepared", fp), ("prepared_valid", C.c_int32), ("flags", C.c_int32),
                ("drop_rate", C.c_float), ("attn_drop_rate", C.c_float), ("drop_path_rate", C.c_float * MAX_BLOCKS),
                ("seed", C.c_uint64)]


class BlockGrads(C.Structure):       # dpmn_block_grads: same fields as dpmn_block_weights
    _fields_ = list(BlockWeights._fields_)


class PgrmGrads(C.Structure):
    _fields_ = [("prior_fusion_w", fp), ("prior_fusion_b", fp), ("pe_w", fp), ("pe_b", fp),
                ("pe_norm_w", fp), ("pe_norm_b", fp),
                ("blocks", BlockGrads * MAX_BLOCKS),
                ("head0_w", fp), ("head0_b", fp), ("head1_w", fp), ("head1_b", fp),
                ("mix_weight", fp * MAX_MIX), ("x_kv", fp), ("mix_input", fp * MAX_MIX)]


class Bn(C.Structure):
    _fields_ = [("w", fp), ("b", fp), ("running_mean", fp), ("running_var", fp)]


class CmmStage(C.Structure):
    _fields_ = [("conv_a_w", fp), ("conv_a_b", fp), ("bn_a", Bn), ("conv_b_w", fp), ("conv_b_b", fp), ("bn_b", Bn)]


class CmmDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("c_img", C.c_int32),
                ("cnum", C.c_int32), ("precision", C.c_int32), ("training", C.c_int32),
                ("update_running_stats", C.c_int32),
                ("x1_batch_stride", C.c_int64), ("x2_batch_stride", C.c_int64),
                ("en1_w", fp * 2), ("en1_b", fp * 2),
                ("enc", (CmmStage * 4) * 2),
                ("en6_w", fp * 2), ("en6_b", fp * 2),
                ("fc1_w", fp), ("fc1_b", fp), ("fc2_w", fp), ("fc2_b", fp),
                ("de6_w", fp), ("de6_b", fp), ("de6_bn", Bn),
                ("dec", CmmStage * 4),
                ("de1_w", fp), ("de1_b", fp),
                ("prepared", fp), ("prepared_valid", C.c_int32), ("flags", C.c_int32),
                ("blend_input", fp), ("blend_input_batch_stride", C.c_int64), ("blend_alpha", C.c_float),
                ("reserved_", C.c_int32)]


class BnGrads(C.Structure):
    _fields_ = [("w", fp), ("b", fp)]


class CmmStageGrads(C.Structure):
    _fields_ = [("conv_a_w", fp), ("conv_a_b", fp), ("bn_a", BnGrads), ("conv_b_w", fp), ("conv_b_b", fp), ("bn_b", BnGrads)]


class CmmGrads(C.Structure):
    _fields_ = [("en1_w", fp * 2), ("en1_b", fp * 2),
                ("enc", (CmmStageGrads * 4) * 2),
                ("en6_w", fp * 2), ("en6_b", fp * 2),
                ("fc1_w", fp), ("fc1_b", fp), ("fc2_w", fp), ("fc2_b", fp),
                ("de6_w", fp), ("de6_b", fp), ("de6_bn", BnGrads),
                ("dec", CmmStageGrads * 4),
                ("de1_w", fp), ("de1_b", fp), ("x1", fp), ("x2", fp)]


class DistillDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("training", C.c_int32),
                ("update_running_stats", C.c_int32), ("flags", C.c_int32), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
                ("deep_batch_stride", C.c_int64), ("shallow_batch_stride", C.c_int64),
                ("conv_cat_w", fp), ("conv_cat_b", fp), ("bn_1", Bn), ("conv_w", fp), ("conv_b", fp), ("bn_2", Bn)]


class DistillGrads(C.Structure):
    _fields_ = [("conv_cat_w", fp), ("conv_cat_b", fp), ("bn_1", BnGrads), ("conv_w", fp), ("conv_b", fp), ("bn_2", BnGrads),
                ("x_deep", fp), ("x_shallow", fp)]


# every symbol include/dpmn_b200.h declares: (restype, argtypes)
_i32, _sz, _vp = C.c_int32, C.c_size_t, C.c_void_p
SYMBOLS = {
    "dpmn_version": (C.c_char_p, []),
    "dpmn_check_device": (C.c_int, []),
    "dpmn_abi_sizeof": (_sz, [_i32]),
    "dpmn_launch_count": (C.c_uint64, []),
    "dpmn_profile_enable": (C.c_int, [_i32]),
    "dpmn_profile_collect": (_i32, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_float), _i32]),
    "dpmn_profile_tag_name": (C.c_char_p, [_i32]),
    "dpmn_pgrm_workspace_bytes": (_sz, [C.POINTER(PgrmDesc)]),
    "dpmn_pgrm_prepared_bytes": (_sz, [C.POINTER(PgrmDesc)]),
    "dpmn_cmm_prepared_bytes": (_sz, [C.POINTER(CmmDesc)]),
    "dpmn_pgrm_forward": (C.c_int, [C.POINTER(PgrmDesc), _vp, _vp, _vp, _vp, _sz, _vp]),
    "dpmn_pgrm_forward_probe": (C.c_int, [C.POINTER(PgrmDesc), _vp, _vp, _vp, _vp, _sz, _vp,
                                          C.POINTER(_vp * MAX_BLOCKS), C.POINTER(_vp * MAX_BLOCKS)]),
    "dpmn_window_attn_forward": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp * MAX_GROUPS), _i32, _i32, _i32, _i32, _i32,
                                           _i32, C.POINTER(_i32 * MAX_GROUPS), C.POINTER(_i32 * MAX_GROUPS), _i32,
                                           _vp, _sz, _vp]),
    "dpmn_window_attn_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "dpmn_window_attn_forward_windowed": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(_vp * MAX_GROUPS), _i32, _i32, _i32, _i32,
                                                    _i32, _i32, C.POINTER(_i32 * MAX_GROUPS), C.POINTER(_i32 * MAX_GROUPS),
                                                    _i32, _vp]),
    "dpmn_window_attn_forward_windowed_train": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(_vp * MAX_GROUPS), _i32, _i32, _i32, _i32,
                                                          _i32, _i32, C.POINTER(_i32 * MAX_GROUPS), C.POINTER(_i32 * MAX_GROUPS),
                                                          _i32, C.c_float, C.c_uint64, C.c_uint32, _vp]),
    "dpmn_window_attn_backward_windowed": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp * MAX_GROUPS),
                                                     C.POINTER(_vp * MAX_GROUPS), _i32, _i32, _i32, _i32, _i32, _i32,
                                                     C.POINTER(_i32 * MAX_GROUPS), C.POINTER(_i32 * MAX_GROUPS), _i32,
                                                     C.c_float, C.c_uint64, C.c_uint32, _vp]),
    "dpmn_cmm_workspace_bytes": (_sz, [C.POINTER(CmmDesc)]),
    "dpmn_cmm_forward": (C.c_int, [C.POINTER(CmmDesc), _vp, _vp, _vp, _vp, _sz, _vp]),
    "dpmn_cmm_debug_bytes": (_sz, [C.POINTER(CmmDesc), _i32]),
    "dpmn_cmm_debug_copy": (C.c_int, [C.POINTER(CmmDesc), _vp, _i32, _vp, _sz, _vp]),
    "dpmn_gemm_nt": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "dpmn_gemm_nt_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "dpmn_mask_hash": (C.c_uint32, [C.c_uint64, C.c_uint32, C.c_uint64]),
    "dpmn_image_loss": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _i32, _i32, _i32, _i32, C.c_float, C.c_float, C.c_float,
                                  _vp, _vp, _vp]),
    "dpmn_to_mask": (C.c_int, [_vp, C.c_int64, _vp, _i32, _i32, _i32, _vp]),
    "dpmn_crnn_input": (C.c_int, [_vp, C.c_int64, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dpmn_visionlan_input": (C.c_int, [_vp, C.c_int64, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "dpmn_distill_workspace_bytes": (_sz, [C.POINTER(DistillDesc)]),
    "dpmn_distill_forward": (C.c_int, [C.POINTER(DistillDesc), _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dpmn_distill_backward": (C.c_int, [C.POINTER(DistillDesc), _vp, _vp, _vp, _vp, C.POINTER(DistillGrads), _vp, _sz, _vp]),
    "dpmn_pgrm_backward_workspace_bytes": (_sz, [C.POINTER(PgrmDesc)]),
    "dpmn_pgrm_backward": (C.c_int, [C.POINTER(PgrmDesc), _vp, _vp, _vp, C.POINTER(PgrmGrads), _vp, _sz, _vp]),
    "dpmn_cmm_backward_workspace_bytes": (_sz, [C.POINTER(CmmDesc)]),
    "dpmn_cmm_backward": (C.c_int, [C.POINTER(CmmDesc), _vp, _vp, _vp, C.POINTER(CmmGrads), _vp, _sz, _vp]),
    "dpmn_nccl_available": (C.c_int, []),
    "dpmn_nccl_version": (C.c_int, []),
    "dpmn_nccl_unique_id": (C.c_int, [_vp]),
    "dpmn_nccl_comm_init": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "dpmn_nccl_comm_destroy": (C.c_int, [_vp]),
    "dpmn_allreduce_bucket": (C.c_int, [_vp, _vp, _sz, _i32, _vp]),
    "dpmn_clip_adam_workspace_bytes": (_sz, [_i32]),
    "dpmn_clip_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(C.c_int64), _i32, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_float, C.c_float, C.c_int64, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Load (once) and type the C-ABI library.  Raises if it is absent or incomplete: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (first: the library resolves NCCL from the copy torch has already mapped, see optim_nccl.cu)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dpmn_b200: {LIB_PATH} is missing.  Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dpmn_b200/csrc`).  There is no CPU or PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    for which, st in enumerate((BlockWeights, PgrmDesc, Bn, CmmStage, CmmDesc, BlockGrads, PgrmGrads, CmmGrads,
                                DistillDesc, DistillGrads)):
        got = lib.dpmn_abi_sizeof(which)
        if got != C.sizeof(st):
            raise RuntimeError(f"dpmn_b200: ABI mismatch for {st.__name__}: library {got} B, binding {C.sizeof(st)} B")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")
