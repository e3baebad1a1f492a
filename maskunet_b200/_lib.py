"""ctypes binding of libmaskunet_b200.so (the C ABI declared in include/maskunet_b200.h).

There is no fallback: if the library is missing the import fails loudly, and a
CPU tensor reaching an op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# MASKUNET_B200_LIB selects another build of the same library (A/B experiments with tools/); default: in-tree
LIB_PATH = os.environ.get("MASKUNET_B200_LIB") or os.path.join(HERE, "libmaskunet_b200.so")

MU_F32, MU_BF16 = 0, 1

_P, _I, _F = c_void_p, c_int32, c_float

# name -> argtypes, mirrors include/maskunet_b200.h one to one
SIGNATURES = {
    "mu_mask_binarize": [_P, _I, _I, _P, _P, _P, _P, _P],
    "mu_qkv_project": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_attn_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_attn_fwd_cudacore": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_residual_ln_fwd": [_P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_residual_ln_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_attn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _I, _I, _I, _I, _I, _P],
    "mu_attn_bwd_cudacore": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_qkv_project_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mu_transpose": [_P, _P, _I, _I, _I, _I, _P],
    "mu_maxpool2": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_upsample_concat_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_upsample_concat_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_sample_layernorm_fwd": [_P, _P, _P, _F, _P, _P, _P, _P, _I, ctypes.c_int64, _I, _P],
    "mu_sample_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, ctypes.c_int64, _I, _P],
    "mu_cross_entropy_fused": [_P, _P, _P, ctypes.c_int64, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    "mu_argmax_iou": [_P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _F, _I, _P],
    "mu_column_sums": [_P, _P, ctypes.c_int64, _I, _I, _P],
    "mu_conv1x1_prep": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mu_conv1x1_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv1x1_fwd_stats": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv1x1_bwd_data": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv1x1_bwd_weight": [_P, _P, _P, ctypes.c_size_t, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv1x1_bwd_weight_bias": [_P, _P, _P, ctypes.c_size_t, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_bn_act_fwd": [_P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    "mu_bn_act_apply": [_P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    "mu_bn_act_fwd_stats": [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
    "mu_conv_prep_weights": [_P, _P, _P, _I, _I, _I, _P],
    "mu_conv3x3_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv3x3_bwd_data": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv3x3_bwd_data_acc": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_conv3x3_bwd_weight": [_P, _P, _P, ctypes.c_size_t, _P, _I, _I, _I, _I, _I, _I, _P],
    "mu_bn_update_running": [_P, _P, _P, _P, _F, _F, ctypes.c_int64, _I, _P],
    "mu_instance_triplet_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _I, _F, _F, _P, _P, _P, _I, _P],
    "mu_instance_triplet_bwd": [_P, _P, _I, _I, _P, _I, _F, _F, _P, _P, _F, _P, _P, _I, _P],
    "mu_query_mask_bits": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P],
    "mu_query_attn_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "mu_query_attn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "mu_to_tensor_u8": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "mu_resize_linear_to_tensor_u8": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "mu_resize_nearest_u8_i64": [_P, _P, _I, _I, _I, _I, _P],
    "mu_bn_act_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, _I, _I, _I, _P],
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. maskunet_b200 has no CPU or PyTorch fallback: build the CUDA library with "
            "`python maskunet_b200/build.py` (needs nvcc, cross-compiles for sm_100a without a GPU).")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mu_version.restype = c_int32
    lib.mu_version.argtypes = []
    lib.mu_last_error.restype = ctypes.c_char_p
    lib.mu_last_error.argtypes = []
    lib.mu_device_supported.restype = c_int32
    lib.mu_device_supported.argtypes = []
    lib.mu_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    lib.mu_attn_bwd_workspace_bytes.argtypes = [_I, _I, _I, _I]
    lib.mu_conv3x3_workspace_bytes.restype = ctypes.c_size_t
    lib.mu_conv3x3_workspace_bytes.argtypes = [_I, _I]
    lib.mu_conv1x1_workspace_bytes.restype = ctypes.c_size_t
    lib.mu_conv1x1_workspace_bytes.argtypes = [_I, _I]
    lib.mu_query_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    lib.mu_query_attn_bwd_workspace_bytes.argtypes = [_I, _I, _I]
    _u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.mu_tmap_cache_stats.restype = None
    lib.mu_tmap_cache_stats.argtypes = [_u64p, _u64p, _u64p]
    lib.mu_tmap_cache_clear.restype = None
    lib.mu_tmap_cache_clear.argtypes = []
    lib.mu_set_deterministic.restype = None
    lib.mu_set_deterministic.argtypes = [c_int32]
    lib.mu_get_deterministic.restype = c_int32
    lib.mu_get_deterministic.argtypes = []
    lib.mu_set_deterministic_scratch.restype = c_int32
    lib.mu_set_deterministic_scratch.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int32
        fn.argtypes = argtypes
    _lib = lib
    return lib


def tmap_cache_stats() -> dict:
    """Counters of the library's TMA descriptor cache: {'hits', 'misses', 'entries'}."""
    h, m, e = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    load().mu_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m), ctypes.byref(e))
    return {"hits": h.value, "misses": m.value, "entries": e.value}


_det_scratch = {}      # device index -> the registered scratch tensor (kept alive here)
DET_SCRATCH_BYTES = 128 << 20


def set_deterministic(on: bool, device=None) -> None:
    """Process-wide switch: fixed-order accumulation in every kernel that sums across CTAs, so that two runs on the
    same inputs return the same bits (free-running mode uses float atomics: last-bit differences that 39 batch-statistics
    BatchNorms amplify).  Turning it on registers a zeroed 128 MiB scratch tensor for ``device`` (default: the current
    CUDA device); call it once per device / process after ``torch.cuda.set_device``.  One compute stream per device.
    Costs: the attention backward adds its dQ partials through order semaphores; everything else is a
    few microseconds per kernel."""
    lib = load()
    if on:
        import torch
        if torch.cuda.is_available():
            idx = torch.cuda.current_device() if device is None else torch.device(device).index
            if idx not in _det_scratch:
                with torch.cuda.device(idx):
                    buf = torch.zeros(DET_SCRATCH_BYTES, dtype=torch.uint8, device=f"cuda:{idx}")
                    check(lib.mu_set_deterministic_scratch(ctypes.c_void_p(buf.data_ptr()), DET_SCRATCH_BYTES),
                          "mu_set_deterministic_scratch")
                _det_scratch[idx] = buf
    lib.mu_set_deterministic(1 if on else 0)


def is_deterministic() -> bool:
    return bool(load().mu_get_deterministic())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mu_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")
