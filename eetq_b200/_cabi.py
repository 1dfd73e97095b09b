"""ctypes binding of libeetq_b200.so (the C ABI declared in include/eetq_b200.h).

This is the only place the native library is loaded.  There is NO fallback: if the library is missing or a call
fails, a RuntimeError is raised (the reference surfaces C++ exceptions as RuntimeError through pybind,
/root/reference/csrc/utils/cuda_utils.h:29-51).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# EETQ_B200_LIB: development override (e.g. the timeline build, tools/timeline.py); the product always loads the in-tree library
LIB_PATH = os.environ.get("EETQ_B200_LIB") or os.path.join(_HERE, "libeetq_b200.so")

F16, BF16, F32 = 0, 1, 2
FLAG_DEFAULT, FLAG_FORCE_GEMV, FLAG_FORCE_TC, FLAG_PDL, FLAG_FORCE_MMA = 0, 1, 2, 4, 8
GEMV_MAX_M = 8
GEMV4_MAX_M = 8

_lib: Optional[ctypes.CDLL] = None


class LL(ctypes.Structure):
    """eetq_b200_ll (include/eetq_b200.h): which exchange an LL buffer currently carries"""
    _fields_ = [("step", ctypes.c_void_p), ("per_step", ctypes.c_int), ("index", ctypes.c_int)]


class LLPush(ctypes.Structure):
    """eetq_b200_ll_push"""
    _fields_ = [("world", ctypes.c_int), ("peers", ctypes.POINTER(ctypes.c_uint64)), ("local", ctypes.c_void_p),
                ("elem_off", ctypes.c_int64), ("step", ctypes.c_void_p), ("per_step", ctypes.c_int), ("index", ctypes.c_int)]


class GemvOpts(ctypes.Structure):
    """eetq_b200_gemv_opts"""
    _fields_ = [("norm_weight", ctypes.c_void_p), ("eps", ctypes.c_float), ("xmode", ctypes.c_int), ("epi", ctypes.c_int),
                ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("x_ll", ctypes.POINTER(LL)), ("residual_ll", ctypes.POINTER(LL)),
                ("residual_off", ctypes.c_int64), ("push", ctypes.POINTER(LLPush)), ("next_w", ctypes.c_void_p),
                ("next_n", ctypes.c_int64), ("next_k", ctypes.c_int64)]


_c_i64 = ctypes.c_int64
_c_vp = ctypes.c_void_p
_c_int = ctypes.c_int
_c_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/eetq_b200.h declares (tests/test_cabi.py checks)
SIGNATURES = {
    "eetq_b200_last_error": (ctypes.c_char_p, []),
    "eetq_b200_version": (_c_int, []),
    "eetq_b200_launch_count": (ctypes.c_uint64, []),
    "eetq_b200_quantize": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "eetq_b200_pack": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_unpack": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_from_ref_layout": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_to_ref_layout": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_i64]),
    "eetq_b200_quantize4": (_c_int, [_c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "eetq_b200_pack4": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_unpack4": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_from_ref_layout4": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_to_ref_layout4": (_c_int, [_c_vp, _c_i64, _c_i64, _c_vp, _c_vp]),
    "eetq_b200_w4a16_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_i64]),
    "eetq_b200_w4a16_gemm": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_int, _c_vp, _c_sz,
                                      _c_int, _c_vp]),
    "eetq_b200_w8a16_gemm": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "eetq_b200_w8a16_gemm_ex": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_int,
                                         _c_vp, _c_sz, _c_int, _c_vp]),
    "eetq_b200_w8a16_gemm_residual": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_int,
                                               _c_vp, _c_sz, _c_int, _c_vp]),
    "eetq_b200_w8a16_gemm_trace": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp, _c_sz, _c_vp, _c_sz,
                                            _c_vp]),
    "eetq_b200_w8a16_gemm_trace_info": (_c_int, [_c_i64, _c_i64, _c_i64, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "eetq_b200_set_timeline": (_c_int, [_c_vp, ctypes.c_uint64]),
    "eetq_b200_decode_attention_occupancy": (_c_int, [_c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "eetq_b200_w8a16_gemm_host": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int,
                                           _c_vp, _c_sz, _c_vp]),
    "eetq_b200_w8a16_gemv_fused": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_int,
                                            ctypes.POINTER(GemvOpts), _c_int, _c_vp]),
    "eetq_b200_decode_embed": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, ctypes.POINTER(LL), _c_int, _c_vp]),
    "eetq_b200_rmsnorm": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, ctypes.c_float, _c_int, _c_vp]),
    "eetq_b200_layernorm_forward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i64, ctypes.c_float, _c_vp]),
    "eetq_b200_rotary_embedding_neox": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_i64, _c_vp]),
    "eetq_b200_prefill_rope_kv": (_c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_vp]),
    "eetq_b200_silu_mul": (_c_int, [_c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_vp]),
    "eetq_b200_decode_attention": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, ctypes.POINTER(LLPush),
                                            _c_vp, _c_i64, _c_i64, _c_int, _c_vp]),
    "eetq_b200_lm_head_scratch_bytes": (_c_sz, []),
    "eetq_b200_lm_head_argmax": (_c_int, [_c_vp, ctypes.POINTER(LL), _c_vp, ctypes.c_float, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp,
                                          _c_vp, _c_vp, ctypes.POINTER(LLPush), _c_int, _c_int, _c_vp]),
}


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library; raise if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(eetq_b200 has no CPU or PyTorch fallback path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().eetq_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().eetq_b200_launch_count())
