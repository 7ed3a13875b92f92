"""ctypes binding of libsemidetr_b200.so (the C ABI declared in include/semidetr_b200.h).

The product path has NO fallback: if the shared library is missing, or an op is called without a
CUDA device, this module raises -- it never routes to a CPU / PyTorch implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsemidetr_b200.so")

c_void_p, c_int, c_float, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double

_MSDA_FWD = [c_void_p] * 6 + [c_int] * 7 + [c_void_p]
_MSDA_BWD = [c_void_p] * 7 + [c_int] * 7 + [c_void_p] * 3

# symbol -> argtypes; every function returns int.  Must list every symbol include/semidetr_b200.h declares
# (tests/test_abi.py parses the header and checks this table and the .so against it).
SIGNATURES = {
    "sdb_abi_version": [],
    "sdb_msda_forward_f32": _MSDA_FWD,
    "sdb_msda_forward_f64": _MSDA_FWD,
    "sdb_msda_backward_f32": _MSDA_BWD,
    "sdb_msda_backward_f64": _MSDA_BWD,
    "sdb_msda_forward_bf16": _MSDA_FWD,
    "sdb_msda_backward_bf16": _MSDA_BWD,
    "sdb_msda_set_variant": [c_int, c_int],
    "sdb_msda_forward_tma_f32": [c_void_p, c_int] + [c_void_p] * 6 + [c_int] + [c_void_p] * 2 + [c_int] * 7 + [c_void_p],
    "sdb_msda_fused_forward_f32": [c_void_p] * 5 + [c_int] + [c_void_p] * 2 + [c_int] * 7 + [c_void_p],
    "sdb_msda_fused_backward_f32": [c_void_p] * 6 + [c_int] + [c_void_p] * 2 + [c_int] * 7 + [c_void_p] * 3,
    "sdb_msda_prologue_forward_f32": [c_void_p] * 4 + [c_int, c_void_p] + [c_int] * 5 + [c_void_p] * 2,
    "sdb_msda_prologue_forward_bf16": [c_void_p] * 4 + [c_int, c_void_p] + [c_int] * 5 + [c_void_p] * 2,
    "sdb_msda_prologue_backward_f32": [c_void_p] * 5 + [c_int, c_void_p] + [c_int] * 5 + [c_void_p] * 2,
    "sdb_msda_prologue_backward_bf16": [c_void_p] * 5 + [c_int, c_void_p] + [c_int] * 5 + [c_void_p] * 2,
    "sdb_match_cost_f32": [c_void_p] * 9 + [c_int] * 3 + [c_float] * 3 + [c_void_p] * 2,
    "sdb_lsap_solve_f32": [c_void_p] * 6 + [c_int] * 3 + [c_void_p] * 3,
    "sdb_hungarian_assign_f32": [c_void_p] * 9 + [c_int] * 4 + [c_float] * 3 + [c_void_p] * 5,
    "sdb_ema_update_f32": [c_void_p, c_void_p, c_int, c_double],
    "sdb_adamw_ema_step_f32": [c_void_p] * 11 + [c_int] + [c_float] * 3 + [c_double],  # seg arrays are HOST pointers
    "sdb_adamw_ema_step_sched_f32": [c_void_p] * 10 + [c_int] + [c_float] * 3 + [c_double],  # seg_bounds is a HOST pointer
    "sdb_detr_loss_forward_f32": [c_void_p] * 10 + [c_int] * 3 + [c_float] * 3 + [c_void_p],
    "sdb_detr_loss_backward_f32": [c_void_p] * 11 + [c_int] * 3 + [c_float] * 3 + [c_void_p] * 2,
    "sdb_pseudo_label_nms_f32": [c_void_p] * 4 + [c_int] * 4 + [c_float] * 2 + [c_int] * 2 + [c_void_p] * 5,
    "sdb_gmm_threshold_f32": [c_void_p] * 3 + [c_int] * 2 + [c_float, c_int, c_double, c_void_p],
    "sdb_mha_forward_f32": [c_void_p] + [c_void_p, ctypes.c_int64, ctypes.c_int64] * 3 + [c_void_p] * 2 + [c_int] * 4 +
                           [c_float, c_void_p, c_void_p],
    "sdb_mha_backward_f32": [c_void_p] + [c_void_p, ctypes.c_int64, ctypes.c_int64] * 3 + [c_void_p] * 6 + [c_int] * 4 +
                            [c_float] + [c_void_p, ctypes.c_int64, ctypes.c_int64] * 3 + [c_void_p],
    "sdb_dp_ctrl_bytes": [],
    "sdb_dp_error_word_offset": [],
    "sdb_dp_adamw_exchange_f32": [c_void_p, c_int, c_int] + [c_void_p] * 10 + [c_int] + [c_float] * 5 + [ctypes.c_int64],
    "sdb_dp_small_allreduce_f32": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int],
    "sdb_colsum_f32": [c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p],
    "sdb_relu_backward_colsum_f32": [c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p, c_void_p],
    "sdb_colsum_bf16": [c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p],
    "sdb_relu_backward_colsum_bf16": [c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p, c_void_p],
    "sdb_gemm_tf32": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "sdb_gemm_tf32_relu_grad": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int],
    "sdb_layernorm_bwd_workspace_floats": [],
    "sdb_layernorm_forward_f32": [c_void_p] * 4 + [ctypes.c_int64, c_int, c_float] + [c_void_p] * 3,
    "sdb_add_layernorm_forward_f32": [c_void_p] * 5 + [ctypes.c_int64, c_int, c_float] + [c_void_p] * 5,
    "sdb_add_layernorm_backward_f32": [c_void_p] * 8 + [ctypes.c_int64, c_int] + [c_void_p] * 4,
    "sdb_layernorm_backward_f32": [c_void_p] * 6 + [ctypes.c_int64, c_int] + [c_void_p] * 4,
}

# include/semidetr_b200_debug.h -> lib/libsemidetr_b200_debug.so (tools/ only; `python -m semi_detr_b200.build --debug`)
DEBUG_LIB_PATH = os.path.join(_HERE, "lib", "libsemidetr_b200_debug.so")
DEBUG_SIGNATURES = {
    "sdb_gemm_tf32": SIGNATURES["sdb_gemm_tf32"],
    "sdb_gemm_tf32_set_trace": [c_void_p],
    "sdb_debug_umma_rate": [c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_debug_tma_rate": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
}

_lib = None
_debug_lib = None


def debug_lib():
    """The instrumentation library (trace stamps, issue-rate microbenchmarks).  Never loaded by the package itself."""
    global _debug_lib
    if _debug_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise RuntimeError(f"{DEBUG_LIB_PATH} is missing: build it with `python -m semi_detr_b200.build --debug`")
        l = ctypes.CDLL(DEBUG_LIB_PATH)
        for name, argtypes in DEBUG_SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        l.sdb_last_error.restype = ctypes.c_char_p
        _debug_lib = l
    return _debug_lib


# kernels of this library launched so far, by entry point (bench.py reports the per-step count)
LAUNCHES = {"msda_forward": 0, "msda_backward": 0, "msda_fused_forward": 0, "msda_fused_backward": 0, "msda_forward_tma": 0, "match_cost": 0, "lsap_solve": 0, "ema_update": 0,
            "layernorm_forward": 0, "layernorm_backward": 0, "adamw_ema_step": 0, "colsum": 0, "gemm_tf32": 0, "relu_backward_colsum": 0, "detr_loss_forward": 0, "detr_loss_backward": 0, "pseudo_label_nms": 0,
            "gmm_threshold": 0, "msda_forward_bf16": 0, "msda_backward_bf16": 0,
            "mha_forward": 0, "mha_backward": 0, "dp_adamw_exchange": 0, "dp_small_allreduce": 0,
            "msda_prologue_forward": 0, "msda_prologue_backward": 0}


class EmaChunk(ctypes.Structure):
    """sdb_ema_chunk"""
    _fields_ = [("teacher", c_void_p), ("student", c_void_p), ("count", ctypes.c_int64)]


def lib():
    """The loaded library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m semi_detr_b200.build` "
                "(semi_detr_b200 has no CPU or PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        l.sdb_last_error.restype = ctypes.c_char_p
        l.sdb_last_error.argtypes = []
        _lib = l
    return _lib


def check(rc, what):
    """Mirror of the reference's AT_ASSERTM / AT_ERROR -> RuntimeError behaviour."""
    if rc != 0:
        msg = lib().sdb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    return None if t is None else t.data_ptr()


def current_stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
