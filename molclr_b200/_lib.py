"""ctypes binding of libmolclr_b200.so (the C ABI declared in include/molclr_b200.h).

There is no fallback: if the shared object is missing or a call fails, this raises.  PyTorch is used
only for device memory (``tensor.data_ptr()``) and the current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOLCLR_B200_LIB") or os.path.join(_HERE, "libmolclr_b200.so")   # (override: A/B timing of builds)

vp, i64, i32, f32, sz, u32 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t, C.c_uint32
STATUS_FP16_RANGE = 1  # MOLCLR_STATUS_FP16_RANGE
ABI_VERSION = 3        # MOLCLR_ABI_VERSION of include/molclr_b200.h


class GemmArgs(C.Structure):
    """Mirror of ``molclr_gemm_args`` (include/molclr_b200.h)."""
    _fields_ = [
        ("A", vp), ("lda", i64), ("a_mn", C.c_int32),
        ("B", vp), ("ldb", i64), ("b_mn", C.c_int32),
        ("A_lo", vp), ("B_lo", vp),
        ("M", i64), ("N", i64), ("K", i64),
        ("out", vp), ("ldo", i64), ("transpose_out", C.c_int32),
        ("out2", vp), ("ldo2", i64),
        ("out_lo", vp), ("ldo_lo", i64),
        ("bias", vp),
        ("addend", vp), ("ldadd", i64),
        ("mask", vp), ("ldmask", i64),
        ("relu", C.c_int32), ("round_out", C.c_int32),
        ("colstat", vp), ("colstat_mode", C.c_int32),
        ("split_k", C.c_int32),
        ("relu_bits", vp), ("mask_bits", vp), ("ld_bits", i64),
        ("compensate", C.c_int32),
        ("B16", vp), ("ld16", i64), ("rows16", i64), ("status", vp),
    ]


class WeightDesc(C.Structure):
    """Mirror of ``molclr_weight_desc`` (include/molclr_b200.h)."""
    _fields_ = [
        ("src", vp), ("ld_src", i64), ("rows", C.c_int32), ("cols", C.c_int32),
        ("hi", vp), ("lo", vp), ("ld_hi", i64),
        ("hi_t", vp), ("ld_hi_t", i64),
        ("raw", vp), ("ld_raw", i64), ("transpose_raw", C.c_int32),
        ("b16", vp), ("ld16", i64), ("rows16", C.c_int32), ("b16_kind", C.c_int32),
    ]


LAYER_CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p)          # molclr_layer_cb


class GinLayer(C.Structure):
    """Mirror of ``molclr_gin_layer``."""
    _fields_ = [("w1_hi", vp), ("w1_raw", vp), ("w1_b16", vp), ("b1", vp), ("w2_hi", vp), ("w2_raw", vp), ("w2_b16", vp), ("b2", vp),
                ("w1_hi_t", vp), ("w2_hi_t", vp), ("bond_type", vp), ("bond_dir", vp), ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("num_batches_tracked", vp), ("momentum", f32), ("eps", f32)]


class GinModel(C.Structure):
    """Mirror of ``molclr_gin_model``."""
    _fields_ = [("num_layer", C.c_int32), ("emb_dim", C.c_int32), ("feat_dim", C.c_int32), ("x_emb1", vp), ("x_emb2", vp),
                ("layers", C.POINTER(GinLayer)), ("w1_ld16", i64), ("w1_rows16", i64), ("w2_ld16", i64), ("w2_rows16", i64),
                ("wf_hi", vp), ("wf_lo", vp), ("bf", vp), ("w0_hi", vp), ("w0_lo", vp), ("b0", vp), ("w2_hi", vp), ("w2_lo", vp), ("b2", vp),
                ("status", vp)]


class PlanView(C.Structure):
    """Mirror of ``molclr_plan_view``."""
    _fields_ = [("N", i64), ("E", i64), ("G", i64), ("xpacked", vp), ("node2graph", vp), ("rowptr", vp), ("col", vp), ("eattr", vp),
                ("rowptr_t", vp), ("col_t", vp), ("cnt", vp), ("nbr", vp), ("nbr_t", vp), ("gptr", vp), ("gperm", vp)]


# name -> (restype, argtypes); every symbol include/molclr_b200.h declares
SIGNATURES = {
    "molclr_abi_version": (i32, []),
    "molclr_last_error": (C.c_char_p, []),
    "molclr_launch_count": (C.c_uint64, []),
    "molclr_set_pdl": (i32, [i32]),
    "molclr_device_info": (i32, [C.POINTER(i32), C.POINTER(i32)]),
    "molclr_plan_workspace_bytes": (sz, [i64, i64, i64]),
    "molclr_plan_build": (i32, [vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]),
    "molclr_augment_views": (i32, [vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, C.c_uint64, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                   vp, vp]),
    "molclr_subgraph_select": (i32, [vp, vp, vp, vp, i64, vp, i64, vp, vp, C.c_uint64, i32, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                     vp]),
    "molclr_subgraph_fill": (i32, [vp, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, i64, vp, vp, i64, vp]),
    "molclr_embed_nodes_fwd": (i32, [vp, vp, vp, i64, i32, vp, vp]),
    "molclr_embed_nodes_bwd_workspace_bytes": (sz, [i64]),
    "molclr_edge_table_grad_workspace_bytes": (sz, [i32]),
    "molclr_prepare_weights": (i32, [C.POINTER(WeightDesc), i32, vp]),
    "molclr_gemm_dw_workspace_bytes": (sz, [i64, i64, i64]),
    "molclr_gemm_dw_ordered": (i32, [vp, i64, vp, i64, i64, i64, i64, vp, i64, vp, sz, vp]),
    "molclr_embed_nodes_bwd": (i32, [vp, vp, i64, i64, i32, vp, vp, vp]),
    "molclr_gine_aggregate_fwd": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, vp, i64, i32, vp, u32, f32, vp]),
    "molclr_rowwise_max_blocks": (i32, []),
    "molclr_gine_aggregate_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i64, i32, vp, i32, vp, C.POINTER(i32), u32, f32, vp]),
    "molclr_relu_bn_bwd_stats": (i32, [vp, vp, vp, i32, i64, i32, vp, vp, C.POINTER(i32), u32, f32, vp]),
    "molclr_gcn_aggregate_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, i64, vp]),
    "molclr_row_sum": (i32, [vp, i32, i32, vp, vp]),
    "molclr_bn_apply_fwd": (i32, [vp, vp, i32, i64, i32, vp, vp, i64, i32, u32, f32, vp]),
    "molclr_bn_tile_stats": (i32, [vp, i64, i32, i32, vp, vp]),
    "molclr_edge_table_grad": (i32, [vp, i64, vp, i64, i32, vp, vp, vp]),
    "molclr_reduce_partials": (i32, [vp, i32, i32, f32, i32, vp, vp]),
    "molclr_bn_finalize_workspace_bytes": (sz, [i32]),
    "molclr_bn_fwd_finalize": (i32, [vp, i32, i32, i64, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp]),
    "molclr_bn_eval_coef": (i32, [vp, vp, vp, vp, f32, i32, vp, vp]),
    "molclr_bn_bwd_finalize": (i32, [vp, i32, i64, i32, vp, vp, i32, vp, vp, vp, vp]),
    "molclr_bn_bwd_apply": (i32, [vp, vp, vp, vp, i32, vp, vp, vp, i64, i32, vp, i64, i32, vp, vp, u32, f32, vp]),
    "molclr_pool_fwd": (i32, [vp, vp, i32, vp, vp, i32, i64, i32, vp, i64, i32, vp, vp, u32, f32, vp]),
    "molclr_dropout_mask": (i32, [u32, f32, i64, i32, vp, vp]),
    "molclr_pool_bwd_stats": (i32, [vp, vp, vp, i32, vp, vp, vp, i64, i32, vp, C.POINTER(i32), u32, f32, vp]),
    "molclr_gemm_colstat_tiles": (i32, [i64]),
    "molclr_gemm_colstat_tile_rows": (i32, []),
    "molclr_gemm_mask_words": (i32, [i64]),
    "molclr_gemm_tile_count": (i32, [i64, i64, i32]),
    "molclr_gemm_workers": (i32, []),
    "molclr_gemm_tf32": (i32, [C.POINTER(GemmArgs), vp]),
    "molclr_gemm_dw": (i32, [vp, i64, vp, i64, i64, i64, i64, vp, i64, vp]),
    "molclr_gemm_dw_acc": (i32, [vp, i64, vp, i64, i64, i64, i64, vp, i64, vp]),
    "molclr_gin_ctx_bytes": (sz, [C.POINTER(GinModel), i64, i64, i32, i32]),
    "molclr_gin_scratch_bytes": (sz, [C.POINTER(GinModel), i64, i64, i32]),
    "molclr_gin_grad_layout": (i64, [C.POINTER(GinModel), C.POINTER(i64)]),
    "molclr_gin_encoder_fwd": (i32, [C.POINTER(GinModel), C.POINTER(PlanView), i32, i32, i32, C.POINTER(u32), f32, vp, sz, vp, sz, vp]),
    "molclr_gin_ctx_pooled": (i32, [C.POINTER(GinModel), i64, i64, i32, i32, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]),
    "molclr_proj_head_fwd": (i32, [C.POINTER(GinModel), i64, i64, i32, i32, vp, vp, vp, vp]),
    "molclr_proj_head_bwd": (i32, [C.POINTER(GinModel), i64, i64, i32, i32, vp, vp, vp, i32, vp, vp, sz, vp]),
    "molclr_gin_encoder_bwd": (i32, [C.POINTER(GinModel), C.POINTER(PlanView), i32, i32, i32, C.POINTER(u32), f32, vp, vp, i32, vp, vp, sz, LAYER_CB, vp, vp]),
    "molclr_add_inplace": (i32, [vp, vp, i64, vp]),
    "molclr_step_timing": (i32, [i32]),
    "molclr_step_timing_read": (i32, [i32, C.POINTER(C.c_double), C.POINTER(i32)]),
    "molclr_round_tf32": (i32, [vp, vp, vp, i64, vp]),
    "molclr_round_tf32_2d": (i32, [vp, i64, vp, vp, i64, i64, i64, vp]),
    "molclr_copy_2d": (i32, [vp, sz, vp, sz, sz, sz, vp]),
    "molclr_act_fwd": (i32, [vp, i32, i64, vp, vp, vp]),
    "molclr_act_bwd": (i32, [vp, vp, i32, i64, vp, vp]),
    "molclr_l2_normalize_fwd": (i32, [vp, i64, i32, f32, vp, vp, vp]),
    "molclr_l2_normalize_bwd": (i32, [vp, vp, vp, i64, i32, f32, vp, vp]),
    "molclr_l2_normalize_bwd_scaled": (i32, [vp, vp, vp, i64, i32, f32, vp, vp, vp]),
    "molclr_ntxent_rows_fwd": (i32, [vp, vp, i64, i64, i32, f32, i32, vp, vp, vp, vp, i64, vp]),
    "molclr_ntxent_h_supported": (i32, [i32, f32]),
    "molclr_ntxent_fwd_h": (i32, [vp, vp, i64, i64, i64, i32, i64, i64, f32, vp, vp, vp, vp, sz, vp]),
    "molclr_ntxent_bwd_h": (i32, [vp, vp, i64, i64, i64, i32, i64, i64, f32, vp, vp, f32, vp, vp, sz, vp]),
    "molclr_ntxent_workspace_bytes": (sz, [i64, i64, i32]),
    "molclr_ntxent_fwd": (i32, [vp, vp, i64, i64, i32, i64, i64, f32, i32, vp, vp, vp, vp, sz, vp]),
    "molclr_ntxent_bwd": (i32, [vp, vp, i64, i64, i32, i64, i64, f32, i32, vp, vp, f32, vp, vp, sz, vp]),
}

_lib = None


def load():
    """Loads the library once; raises if it has not been built (python -m molclr_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"molclr_b200: {LIB_PATH} not found -- build it with `python -m molclr_b200.build` "
                               "(there is no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)         # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.molclr_abi_version() != ABI_VERSION:
            raise RuntimeError("molclr_b200: ABI version mismatch between _lib.py and libmolclr_b200.so")
        if os.environ.get("MOLCLR_B200_PDL", "1") == "0":      # escape hatch / A-B timing: plain stream-ordered launches
            lib.molclr_set_pdl(0)
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().molclr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"molclr_b200.{what} failed ({rc}): {msg}")


def stream():
    """Raw cudaStream_t of the current stream of the current device (the fast private accessor: torch.cuda.current_stream() costs
    ~15 us per call, and every launch asks)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of the expected dtype (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("molclr_b200: expected a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"molclr_b200: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("molclr_b200: expected a contiguous tensor")
    return t.data_ptr()


def ptr2d(t):
    """Device pointer of a 2-D fp32 CUDA matrix whose rows are contiguous (row stride = leading dimension)."""
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1):
        raise RuntimeError("molclr_b200: expected a 2-D fp32 CUDA matrix with contiguous rows")
    return t.data_ptr()
