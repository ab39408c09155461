"""ctypes binding of include/rmem_b200.h.  The product path has no CPU fallback: if the CUDA extension
cannot be loaded this module raises, and every entry point raises RmemError on a non-zero status."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librmem_b200.so")

ATTN_DENSE = 0
ATTN_TC = 1
ATTN_TC2 = 2
ATTN_TC3 = 3
ATTN_TC4 = 4
MAX_BANK_FRAMES = 16


class RmemError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("B", C.c_void_p), ("ldb", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("conv", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Cin", C.c_int), ("Wout", C.c_int),
        ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("alpha", C.c_float), ("bias", C.c_void_p), ("bias_along_m", C.c_int),
        ("act", C.c_int), ("act_from_col", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_longlong), ("gate", C.c_void_p), ("ldg", C.c_longlong),
        ("accumulate", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_is_f32", C.c_int),
        ("C2", C.c_void_p), ("ldc2", C.c_longlong), ("c2_is_f32", C.c_int), ("n_split", C.c_int),
        ("pad_n_ok", C.c_int), ("n_images", C.c_int),
    ]


class EngineConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("H", C.c_int), ("W", C.c_int), ("former_mem_len", C.c_int),
                ("latter_mem_len", C.c_int), ("max_engines", C.c_int), ("attn_impl", C.c_int),
                ("long_term_mem_gap", C.c_int), ("no_long_memory", C.c_int), ("reverse_infer", C.c_int),
                ("time_encode", C.c_int), ("gru_memory", C.c_int)]


class WeightEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_size_t), ("nbytes", C.c_size_t)]


# every symbol include/rmem_b200.h declares (tests/test_capi_symbols.py checks header <-> this list <-> .so)
SYMBOLS = [
    "rmem_version", "rmem_last_error", "rmem_operand_dtype", "rmem_gemm_fwd", "rmem_set_gemm_impl", "rmem_long_attn_workspace_bytes", "rmem_long_attn_fwd", "rmem_long_attn_grid_fwd", "rmem_mha_workspace_bytes", "rmem_mha_fwd", "rmem_debug_attn_rescale_counter", "rmem_debug_attn_trace", "rmem_debug_attn_schedule", "rmem_debug_attn_events", "rmem_debug_gemm_trace", "rmem_debug_gemm_force", "rmem_debug_gemm_log", "rmem_debug_gemm_log_count",
    "rmem_qprep_fwd", "rmem_temporal_pe_slots", "rmem_local_attn_fwd", "rmem_local_attn_tc_workspace_bytes", "rmem_local_attn_tc_fwd", "rmem_layernorm_fwd", "rmem_groupnorm_fwd",
    "rmem_dwconv5x5_fwd", "rmem_upsample_bilinear_fwd", "rmem_transpose_fwd", "rmem_maxpool3x3s2_fwd",
    "rmem_pack_image_fwd", "rmem_pack_image_padded_fwd", "rmem_idbank_fwd", "rmem_mask_head_fwd", "rmem_tta_head_fwd", "rmem_preprocess_fwd", "rmem_evict_relevance_fwd", "rmem_evict_pick",
    "rmem_engine_arena_bytes", "rmem_engine_create", "rmem_engine_destroy", "rmem_engine_restart",
    "rmem_engine_set_gap", "rmem_engine_add_reference_frame", "rmem_engine_propagate", "rmem_engine_update_memory",
    "rmem_engine_num_groups", "rmem_engine_long_indexes", "rmem_engine_pred_logits", "rmem_engine_last_evict",
    "rmem_engine_layer_memory",
    "rmem_engine_launch_count", "rmem_engine_prefetch", "rmem_engine_prefetch2", "rmem_engine_prefetch_n", "rmem_engine_set_timing", "rmem_engine_get_timing",
    "rmem_train_loss_workspace_bytes", "rmem_train_loss_fwd_bwd", "rmem_train_predict_mask",
]

_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if needed and possible) the CUDA extension.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except Exception as e:  # missing or stale .so that cannot be rebuilt (no nvcc, compile error) -> fail loudly
            if not os.path.exists(LIB_PATH):
                raise RmemError(f"rmem_b200 CUDA extension is missing and could not be built: {e}") from e
            if not _build.up_to_date() and os.environ.get("RMEM_ALLOW_STALE_LIB") != "1":
                raise RmemError(f"{LIB_PATH} was built from other sources than the ones in this tree and the rebuild "
                                f"failed ({e}); set RMEM_ALLOW_STALE_LIB=1 to load it anyway") from e
    if not os.path.exists(LIB_PATH):
        raise RmemError(f"rmem_b200 CUDA extension not found at {LIB_PATH}; run `python -m rmem_b200.build`")
    lib = C.CDLL(LIB_PATH)
    lib.rmem_last_error.restype = C.c_char_p
    lib.rmem_operand_dtype.restype = C.c_char_p
    lib.rmem_engine_launch_count.restype = C.c_longlong
    lib.rmem_engine_destroy.restype = None
    for s in SYMBOLS:
        getattr(lib, s)  # AttributeError if the .so does not export it
    _lib = lib
    return lib


def op_dtype():
    """torch dtype of the 16-bit tensor-core operands the extension was built for (fp16 default)."""
    import torch
    return torch.float16 if load().rmem_operand_dtype() == b"fp16" else torch.bfloat16


def check(rc: int):
    if rc != 0:
        msg = load().rmem_last_error()
        raise RmemError(f"rmem_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) pointer of a torch tensor as c_void_p; None -> NULL."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
