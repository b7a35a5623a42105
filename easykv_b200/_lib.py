"""ctypes binding of `libeasykv_b200.so` (C ABI: `include/easykv_b200.h`).

The CUDA library is the only compute path of this package: `load()` raises when it is missing or
when its ABI version differs — there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeasykv_b200.so")
ABI_VERSION = 8

F16, BF16, F32 = 0, 1, 2
POLICY_NONE, POLICY_ROCO, POLICY_H2O, POLICY_TOVA, POLICY_RANGE = 0, 1, 2, 3, 4
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3

EXPORTS = ("ekv_abi_version", "ekv_last_error", "ekv_scratch_bytes", "ekv_attend_evict", "ekv_select",
           "ekv_evict_explicit", "ekv_export_logical", "ekv_launch_count", "ekv_debug_set_timeline", "ekv_debug_set_dispatch", "ekv_rope_qk", "ekv_rope_cache", "ekv_sample_top_p", "ekv_token_nll", "ekv_debug_umma_probe", "ekv_debug_set_chunk_variant", "ekv_chunk_entry_limit")


class Step(C.Structure):
    _fields_ = [("policy", C.c_int32), ("accumulate", C.c_int32), ("evict", C.c_int32), ("apply", C.c_int32),
                ("score_offset", C.c_int32), ("counter_add", C.c_float), ("c_new0", C.c_float),
                ("c_new_step", C.c_float), ("k_feasible", C.c_int32), ("protect_last", C.c_int32),
                ("sink_protect", C.c_int32), ("win_lo", C.c_int32), ("win_recent", C.c_int32),
                ("range_start", C.c_int32), ("arith", C.c_int32), ("tova_head_mean", C.c_int32),
                ("raw_colsum", C.c_int32), ("budget_gate", C.c_int32)]


class LayerIO(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k_new", C.c_void_p), ("v_new", C.c_void_p), ("out", C.c_void_p),
                ("K", C.c_void_p), ("V", C.c_void_p), ("S", C.c_void_p), ("SQ", C.c_void_p), ("C", C.c_void_p),
                ("lidx", C.c_void_p), ("new_slots", C.c_void_p), ("victim_slots", C.c_void_p),
                ("victim_lidx", C.c_void_p), ("scratch", C.c_void_p), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
                ("k_new_raw", C.c_void_p), ("seq_n_before", C.c_void_p)]


class Shape(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("Hkv", C.c_int32), ("d", C.c_int32),
                ("q_len", C.c_int32), ("cap", C.c_int32), ("n_before", C.c_int32), ("n_phys", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m easykv_b200.build` (nvcc, sm_100a). "
            "easykv_b200 has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.ekv_abi_version.restype = C.c_int
    lib.ekv_last_error.restype = C.c_char_p
    lib.ekv_launch_count.restype = C.c_int64
    lib.ekv_scratch_bytes.restype = C.c_int64
    lib.ekv_scratch_bytes.argtypes = [C.POINTER(Shape), C.POINTER(Step)]
    lib.ekv_chunk_entry_limit.restype = C.c_int32
    lib.ekv_chunk_entry_limit.argtypes = [C.POINTER(Shape), C.c_int32, C.c_int32]
    lib.ekv_attend_evict.restype = C.c_int
    lib.ekv_attend_evict.argtypes = [C.POINTER(Shape), C.POINTER(LayerIO), C.POINTER(Step), C.c_int32, C.c_void_p]
    lib.ekv_select.restype = C.c_int
    lib.ekv_select.argtypes = [C.POINTER(Shape), C.POINTER(LayerIO), C.POINTER(Step), C.c_void_p]
    lib.ekv_evict_explicit.restype = C.c_int
    lib.ekv_evict_explicit.argtypes = [C.POINTER(Shape), C.POINTER(LayerIO), C.c_void_p, C.c_int32, C.c_void_p]
    lib.ekv_export_logical.restype = C.c_int
    lib.ekv_export_logical.argtypes = [C.POINTER(Shape), C.POINTER(LayerIO), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ekv_rope_qk.restype = C.c_int
    lib.ekv_rope_qk.argtypes = [C.POINTER(Shape)] + [C.c_void_p] * 10
    lib.ekv_rope_cache.restype = C.c_int
    lib.ekv_rope_cache.argtypes = [C.POINTER(Shape), C.POINTER(LayerIO), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ekv_sample_top_p.restype = C.c_int
    lib.ekv_sample_top_p.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ekv_token_nll.restype = C.c_int
    lib.ekv_token_nll.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ekv_debug_umma_probe.restype = C.c_int
    lib.ekv_debug_umma_probe.argtypes = [C.c_int32] + [C.c_void_p] * 7
    lib.ekv_debug_set_chunk_variant.restype = None
    lib.ekv_debug_set_chunk_variant.argtypes = [C.c_int32]
    lib.ekv_debug_set_dispatch.restype = None
    lib.ekv_debug_set_dispatch.argtypes = [C.c_int32, C.c_int32]
    lib.ekv_debug_set_timeline.restype = None
    lib.ekv_debug_set_timeline.argtypes = [C.c_void_p]
    v = lib.ekv_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {v}, this package needs {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc):
    """Map the C status convention onto the reference's exceptions (ValueError on bad shapes,
    `easykv/llama_patch.py:204-228`)."""
    if rc == OK:
        return
    msg = load().ekv_last_error().decode(errors="replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
