"""w8a16_oracle.py -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

CPU restatement (vectorised torch/numpy) of the one EETQ hot path this repository replaces:

* Q1  per-output-channel symmetric INT8 quantisation   (csrc/cutlass_kernels/cutlass_preprocessors.cc:581-678)
* Q3  the reference sm75..sm89 weight layout           (cutlass_preprocessors.cc:497-534)
* K1  the w8a16 GEMM arithmetic                         (cutlass_extensions/.../mma_tensorop_dequantizer.h:259-274,
                                                         default_fpA_intB_traits.h:110, fpA_intB_gemm_template.h:133)
* the north-star CPU baseline "dequantise -> torch.matmul" (examples/layers/test_w8a16_gemm.py:44-47,
                                                         python/eetq/modules/qlinear.py:83-86)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  ``eetq_b200`` (the product) never does; it raises if its CUDA library is
missing instead of falling back to anything in here.

Parity pin: the reference has no golden vectors or asserting tests, so this oracle is
pinned against the reference C++ itself -- compiled UNMODIFIED into ``oracle/_ref/libref_oracle.so``
by ``oracle/Makefile`` -- in ``tests/test_oracle.py`` (live, when the .so is present) and through the
fixtures under ``tests/golden/`` (generated from that same library by ``tests/golden/make_golden.py``).
A second, independent scalar restatement lives in ``w8a16_oracle.c``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))

# k offsets inside each group of 16, in storage order
_REF_K_ORDER = torch.tensor([0, 8, 1, 9, 2, 10, 3, 11, 4, 12, 5, 13, 6, 14, 7, 15])


# ------------------------------------------------------------------------------------------------
# Q1 quantiser
# ------------------------------------------------------------------------------------------------
def quantize(w_kn: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``quant_weights`` arithmetic for a CPU ``[K, N]`` (or ``[E, K, N]``) fp16/fp32 tensor.

    Returns ``(q int8 [..,K,N] row-major, scales (dtype of w) [..,N], s32 fp32 [..,N])``.
    Follows cutlass_preprocessors.cc:608-649: fp32 abs-max per output column, ``s32 = amax * (1/128)``,
    ``q = int8(clamp(round_half_away(w / s32), -128, 127))`` dividing by the FP32 scale; an all-zero
    column yields NaN -> 127 (std::min/max comparison order) with scale 0.
    """
    assert w_kn.dtype in (torch.float16, torch.float32) and w_kn.dim() in (2, 3)
    wf = w_kn.float()
    # std::max(acc, NaN) keeps acc (cutlass_preprocessors.cc:626): NaN entries do not contribute to the abs-max
    s32 = torch.nan_to_num(wf.abs(), nan=0.0, posinf=float("inf")).amax(dim=-2) * np.float32(1.0 / 128.0)
    r = wf / s32.unsqueeze(-2)
    r = torch.where(r >= 0, torch.floor(r + 0.5), torch.ceil(r - 0.5))
    # round-half-away via floor(r+0.5) is exact here: |r| <= 128 so r+0.5 is representable in fp32
    # whenever r has a fractional part (|r| < 2^22).
    r = torch.where(torch.isnan(r), torch.full_like(r, 127.0), r)
    q = torch.clamp(r, -128, 127).to(torch.int8)
    return q, s32.to(w_kn.dtype), s32


# ------------------------------------------------------------------------------------------------
# Q3 reference layout  (what EETQ / HF checkpoints store in `qweight` / `weight`)
# ------------------------------------------------------------------------------------------------
def ref_layout(q_kn: torch.Tensor) -> torch.Tensor:
    """Row-major int8 ``[K, N]`` -> the reference's sm80 interleaved bytes, shaped ``[K, N]`` int8.

    Closed form of preprocess_weights_for_mixed_gemm (cutlass_preprocessors.cc:497-534): bytes viewed as
    ``[N/2][K/64][2][4][16]`` hold ``uint8(q[k, n] + 128)`` with ``n = 2*pair + c`` and
    ``k = 64*ktile + 16*g + (p>>1) + 8*(p&1)``.
    """
    K, N = q_kn.shape
    if K % 64 or N % 64:
        raise ValueError("reference layout needs K % 64 == 0 and N % 64 == 0")
    u = (q_kn.to(torch.int16) + 128).to(torch.uint8)
    t = u.t().contiguous().view(N // 2, 2, K // 64, 4, 16)  # [pair, c, ktile, g, j]
    t = t[..., _REF_K_ORDER].permute(0, 2, 1, 3, 4).contiguous()  # [pair, ktile, c, g, p]
    return t.view(K, N).view(torch.int8)


def ref_layout_inv(w_ref: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`ref_layout`: reference bytes (shaped ``[K, N]``) -> row-major int8 ``[K, N]``."""
    K, N = w_ref.shape
    if K % 64 or N % 64:
        raise ValueError("reference layout needs K % 64 == 0 and N % 64 == 0")
    t = w_ref.contiguous().view(torch.uint8).view(N // 2, K // 64, 2, 4, 16).permute(0, 2, 1, 3, 4)
    t = t[..., torch.argsort(_REF_K_ORDER)].contiguous().view(N, K)
    return (t.t().to(torch.int16) - 128).to(torch.int8).contiguous()


# ------------------------------------------------------------------------------------------------
# the layout our kernels consume -- restated so the CUDA packer can be checked
# ------------------------------------------------------------------------------------------------
def b200_layout(q_kn: torch.Tensor) -> torch.Tensor:
    """Row-major int8 ``[K, N]`` -> output-feature-major biased bytes ``out[n*K + k] = uint8(q[k, n] + 128)``,
    returned as int8 *shaped* ``[K, N]`` (the nominal shape the reference API uses for its processed tensor,
    fpA_intB_gemm_wrapper.cu:71)."""
    K, N = q_kn.shape
    u = (q_kn.t().contiguous().to(torch.int16) + 128).to(torch.uint8)
    return u.view(torch.int8).view(K, N)


def b200_layout_inv(w_b200: torch.Tensor) -> torch.Tensor:
    K, N = w_b200.shape
    u = w_b200.contiguous().view(torch.uint8).view(N, K)
    return (u.to(torch.int16) - 128).to(torch.int8).t().contiguous()


# ------------------------------------------------------------------------------------------------
# packed int4 (QuantType::PACKED_INT4_WEIGHT_ONLY) -- quant_weights(w, torch.quint4x2) / preprocess_weights(w, is_int4=True)
# ------------------------------------------------------------------------------------------------
def pack_int4(q_kn: torch.Tensor) -> torch.Tensor:
    """int8 values in [-8, 7], ``[.., K, N]`` -> the reference's packed form ``[.., K, N/2]`` int8: low nibble = even column,
    high nibble = odd column (cutlass_preprocessors.cc:651-669)."""
    u = (q_kn.to(torch.int16) & 0xF).to(torch.uint8)
    return (u[..., 0::2] | (u[..., 1::2] << 4)).view(torch.int8).contiguous()


def unpack_int4(p_kn2: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`pack_int4`: ``[.., K, N/2]`` packed -> sign-extended int8 ``[.., K, N]``."""
    u = p_kn2.contiguous().view(torch.uint8)
    lo = (u & 0xF).to(torch.int16)
    hi = (u >> 4).to(torch.int16)
    q = torch.stack([lo, hi], dim=-1).reshape(*u.shape[:-1], u.shape[-1] * 2)
    return (((q + 8) & 0xF) - 8).to(torch.int8)


def quantize4(w_kn: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``quant_weights(w, torch.quint4x2)`` arithmetic (cutlass_preprocessors.cc:608-669).

    Returns ``(packed int8 [..,K,N/2], scales (dtype of w) [..,N], s32 fp32 [..,N], q int8 [..,K,N] in [-8,7])``.
    ``s32 = amax * (1/8)``; ``q = clamp(int(round_half_away(w / s32)), -8, 7)``.  The reference converts the rounded float to
    ``int`` BEFORE clamping (:658-659); for NaN (an all-zero column: 0/0, or a NaN weight) that conversion yields INT_MIN on
    x86-64 (cvttss2si's "integer indefinite"), which clamps to -8 -- pinned against the compiled reference in tests/test_oracle.py."""
    assert w_kn.dtype in (torch.float16, torch.float32) and w_kn.dim() in (2, 3)
    wf = w_kn.float()
    s32 = torch.nan_to_num(wf.abs(), nan=0.0, posinf=float("inf")).amax(dim=-2) * np.float32(1.0 / 8.0)
    r = wf / s32.unsqueeze(-2)
    r = torch.where(r >= 0, torch.floor(r + 0.5), torch.ceil(r - 0.5))
    r = torch.where(torch.isnan(r), torch.full_like(r, -8.0), r)
    q = torch.clamp(r, -8, 7).to(torch.int8)
    return pack_int4(q), s32.to(w_kn.dtype), s32, q


# rows of each group of 32 are read in this order by permute_B_rows_for_mixed_gemm (cutlass_preprocessors.cc:137-195, int4 case)
_REF_ROW_PERM4 = torch.tensor([8 * ((t % 8) // 2) + t % 2 + 2 * (t // 8) for t in range(32)])
# destination nibble d of every 32-bit word takes source nibble _REF_NIB4[d] (add_bias_and_interleave_int4s_inplace, :360-418)
_REF_NIB4 = torch.tensor([0, 2, 4, 6, 1, 3, 5, 7])


def ref_layout4(q_kn: torch.Tensor) -> torch.Tensor:
    """int4 values ``[K, N]`` (int8 in [-8, 7]) -> the reference's sm80 int4 bytes, shaped ``[K, N/2]`` int8.

    Closed form of preprocess_weights_for_mixed_gemm for PACKED_INT4_WEIGHT_ONLY (cutlass_preprocessors.cc:497-534): row
    permutation inside groups of 32 k, element transpose, ColumnMajorTileInterleave<64, 4>, +8 bias and the nibble interleave.
    The 32-bit words viewed as ``[N/4][K/64][4][8]`` = (n4, ktile, c, v) hold, in destination nibble d,
    ``(q[perm(64*ktile + 8*v + e), 4*n4 + c] + 8) & 15`` with ``e = _REF_NIB4[d]``."""
    K, N = q_kn.shape
    if K % 64 or N % 64:
        raise ValueError("reference int4 layout needs K % 64 == 0 and N % 64 == 0")
    u = ((q_kn.to(torch.int16) + 8) & 0xF).to(torch.uint8)                    # [K, N] biased nibbles
    kidx = torch.arange(K)
    perm = 32 * (kidx // 32) + _REF_ROW_PERM4[kidx % 32]
    a2 = u[perm].t().contiguous()                                            # [N, K]: A2[n, k'] = u[perm(k'), n]
    t = a2.view(N // 4, 4, K // 64, 8, 8).permute(0, 2, 1, 3, 4)             # [n4, ktile, c, v, e]
    t = t[..., _REF_NIB4].contiguous()                                       # [n4, ktile, c, v, d]
    lo, hi = t[..., 0::2], t[..., 1::2]                                      # nibble 2b -> low half of byte b
    return (lo | (hi << 4)).contiguous().view(torch.int8).view(K, N // 2)


def ref_layout4_inv(w_ref: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`ref_layout4`: reference int4 bytes ``[K, N/2]`` -> int8 values ``[K, N]`` in [-8, 7]."""
    K, N2 = w_ref.shape
    N = N2 * 2
    b = w_ref.contiguous().view(torch.uint8).view(N // 4, K // 64, 4, 8, 4)
    t = torch.stack([b & 0xF, b >> 4], dim=-1).reshape(N // 4, K // 64, 4, 8, 8)   # [n4, ktile, c, v, d]
    t = t[..., torch.argsort(_REF_NIB4)]                                          # [.., e]
    a2 = t.permute(0, 2, 1, 3, 4).contiguous().view(N, K)                          # A2[n, k']
    kidx = torch.arange(K)
    perm = 32 * (kidx // 32) + _REF_ROW_PERM4[kidx % 32]
    u = torch.empty(K, N, dtype=torch.uint8)
    u[perm] = a2.t()
    return (u.to(torch.int16) - 8).to(torch.int8).contiguous()


def b200_layout4(q_kn: torch.Tensor) -> torch.Tensor:
    """int4 values ``[K, N]`` -> the b200 int4 layout, shaped ``[K, N/2]`` int8 like the reference's
    processed tensor: output-feature-major rows of K/2 bytes; in every 32-bit word (8 consecutive k) nibble p < 4 holds
    ``q[8j + 2p] + 8`` and nibble 4 + p holds ``q[8j + 2p + 1] + 8`` -- ``(word >> 4p) & 0x000f000f`` is the adjacent-k pair."""
    K, N = q_kn.shape
    u = ((q_kn.t().contiguous().to(torch.int16) + 8) & 0xF).to(torch.uint8).view(N, K // 8, 8)   # [n, word, k%8]
    t = u[..., _REF_NIB4]                                                                         # nibble d <- k offset
    return (t[..., 0::2] | (t[..., 1::2] << 4)).contiguous().view(torch.int8).view(K, N // 2)


def b200_layout4_inv(w4: torch.Tensor) -> torch.Tensor:
    K, N2 = w4.shape
    N = N2 * 2
    b = w4.contiguous().view(torch.uint8).view(N, K // 8, 4)
    t = torch.stack([b & 0xF, b >> 4], dim=-1).reshape(N, K // 8, 8)
    u = t[..., torch.argsort(_REF_NIB4)].reshape(N, K)
    return (u.to(torch.int16) - 8).to(torch.int8).t().contiguous()


# ------------------------------------------------------------------------------------------------
# K1 GEMM arithmetic (the parity target) and the north-star CPU baseline
# ------------------------------------------------------------------------------------------------
def dequantize(q_kn: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """``Wd = fp16(q) * s`` with ONE fp16 rounding per weight (what both reference kernels feed the MAC).
    For bf16 scales (our extension; the reference has no bf16 path) the product is rounded to bf16."""
    return q_kn.to(scales.dtype) * scales


def gemm(x: torch.Tensor, q_kn: torch.Tensor, scales: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``y = fp16( x.float() @ fp16(fp16(q) * s).float() )`` -- fp32 accumulation, fp16 output (A3 in
    SURVEY.md appendix A).  ``bias`` (separate torch add in the reference, qlinear.py:61) is added in the
    output dtype after rounding, exactly like the reference module does."""
    wd = dequantize(q_kn, scales.to(x.dtype))
    y = (x.float() @ wd.float()).to(x.dtype)
    if bias is not None:
        y = y + bias
    return y


def gemm_f64(x: torch.Tensor, q_kn: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """Same sum in float64 (ground truth for tolerance studies; no output rounding)."""
    wd = dequantize(q_kn, scales.to(x.dtype))
    return x.double() @ wd.double()


def cpu_dequant_matmul(x: torch.Tensor, q_kn: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """The north-star CPU baseline: EETQ-style dequantise -> ``torch.matmul`` on the host cores
    (examples/layers/test_w8a16_gemm.py:44-47).  Identical arithmetic to :func:`gemm`; kept as a separate
    name because bench.py times *this* call (dequant included, every call, like a CPU EetqLinear would)."""
    wd = q_kn.to(torch.float16) * scales.to(torch.float16)
    return (x.float() @ wd.float()).to(x.dtype)


# ---------------------------------------------------------------------------------------------------------------
# glue ops the reference module also exports (csrc/eetpy.cpp:18-19).  CUDA-only in the reference (they cannot be built
# without a GPU), so these restatements follow the kernels' arithmetic line by line: parity for them is "unpinned" in the
# sense of the task text (no golden vectors, no executable reference) and rests on this restatement alone.
# ---------------------------------------------------------------------------------------------------------------
def rotary_embedding_neox(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor, head_size: int,
                          cos_sin_cache: torch.Tensor):
    """csrc/embedding_kernels/pos_encoding_kernels.cu:12-53.  query/key [num_tokens, num_heads, head_size] fp16;
    cos_sin_cache [max_position, rot_dim] = cos | sin; returns rotated COPIES.  Every product and the final
    difference / sum are separate fp16 operations (``q_x * cos - q_y * sin`` on ``__half`` operands)."""
    rot = cos_sin_cache.shape[1]
    e = rot // 2
    q, k = query.clone(), key.clone()
    cs = cos_sin_cache[positions.long()]                       # [T, rot]
    c, s = cs[:, None, :e], cs[:, None, e:]                    # broadcast over heads

    def rot_half(t):
        x, y = t[..., :e].clone(), t[..., e:rot].clone()
        t[..., :e] = (x * c).half() - (y * s).half()           # each torch op on fp16 tensors rounds once
        t[..., e:rot] = (y * c).half() + (x * s).half()
        return t

    return rot_half(q), rot_half(k)


def layernorm_forward(x: torch.Tensor, gamma: torch.Tensor, eps: float) -> torch.Tensor:
    """csrc/layernorm_kernels/layernorm.cu:25-51 (generalT5LayerNorm): fp32 variance of each row, ONE fp16 rounding of
    ``(x * rsqrt(var + eps)) * gamma`` clamped to the fp16 range (clamp_inf_for_half)."""
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return ((xf * r) * gamma.float()).clamp(-65504.0, 65504.0).half()


def norm_rel_err(y: torch.Tensor, y_ref: torch.Tensor) -> float:
    """The parity metric of BASELINE.md §5: ``max|y - y_ref| / max|y_ref|``."""
    d = (y.double() - y_ref.double()).abs().max().item()
    return d / max(y_ref.double().abs().max().item(), 1e-30)


# ------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def synth_weight(K: int, N: int, seed: int, dtype=torch.float16) -> torch.Tensor:
    """``[K, N]`` = ``nn.Linear.weight.t()`` with N(0, 0.02^2) entries (Llama init)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(K, N, generator=g) * 0.02).to(dtype)


def synth_act(M: int, K: int, seed: int = 7, dtype=torch.float16, uniform: bool = False) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(M, K, generator=g) if uniform else torch.randn(M, K, generator=g)
    return x.to(dtype)


# ------------------------------------------------------------------------------------------------
# loaders for the compiled checkers
# ------------------------------------------------------------------------------------------------
def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


_ref_lib = None
_port_lib = None


def ref_lib() -> Optional[ctypes.CDLL]:
    """``oracle/_ref/libref_oracle.so`` (reference quantiser/preprocessor, unmodified) or None."""
    global _ref_lib
    if _ref_lib is None:
        p = os.path.join(_HERE, "_ref", "libref_oracle.so")
        if not os.path.exists(p):
            return None
        _ref_lib = ctypes.CDLL(p)
        for f in ("ref_quant_fp16", "ref_quant_fp32", "ref_preprocess", "ref_quant4_fp16", "ref_quant4_fp32", "ref_preprocess4"):
            if hasattr(_ref_lib, f):
                getattr(_ref_lib, f).restype = ctypes.c_int
    return _ref_lib


def port_lib() -> Optional[ctypes.CDLL]:
    """``oracle/libw8a16_oracle.so`` (the scalar C restatement) or None."""
    global _port_lib
    if _port_lib is None:
        p = os.path.join(_HERE, "libw8a16_oracle.so")
        if not os.path.exists(p):
            return None
        _port_lib = ctypes.CDLL(p)
        _port_lib.oracle_ref_layout.restype = ctypes.c_int
        _port_lib.oracle_ref_layout_inv.restype = ctypes.c_int
    return _port_lib


def ref_quantize(w_kn: torch.Tensor):
    """Run the REFERENCE ``symmetric_quantize`` (+ its layout pass). Returns (unprocessed, processed, scales)."""
    lib = ref_lib()
    assert lib is not None, "oracle/_ref/libref_oracle.so not built (make -C oracle ref)"
    K, N = w_kn.shape
    w_kn = w_kn.contiguous()
    unp = torch.empty(K, N, dtype=torch.int8)
    pro = torch.empty(K, N, dtype=torch.int8)
    sc = torch.empty(N, dtype=w_kn.dtype)
    fn = lib.ref_quant_fp16 if w_kn.dtype == torch.float16 else lib.ref_quant_fp32
    rc = fn(_ptr(pro), _ptr(unp), _ptr(sc), _ptr(w_kn), ctypes.c_size_t(K), ctypes.c_size_t(N))
    assert rc == 0, "reference quantiser threw"
    return unp, pro, sc


def ref_preprocess(q_kn: torch.Tensor) -> torch.Tensor:
    """Run the REFERENCE ``preprocess_weights`` (arch 80) on row-major int8 ``[K, N]``."""
    lib = ref_lib()
    assert lib is not None, "oracle/_ref/libref_oracle.so not built (make -C oracle ref)"
    K, N = q_kn.shape
    q_kn = q_kn.contiguous()
    out = torch.empty(K, N, dtype=torch.int8)
    rc = lib.ref_preprocess(_ptr(out), _ptr(q_kn), ctypes.c_size_t(K), ctypes.c_size_t(N))
    assert rc == 0, "reference preprocessor threw"
    return out


def ref_quantize4(w_kn: torch.Tensor):
    """Run the REFERENCE ``symmetric_quantize`` with PACKED_INT4_WEIGHT_ONLY. Returns (unprocessed packed [K,N/2],
    processed [K,N/2], scales)."""
    lib = ref_lib()
    assert lib is not None and hasattr(lib, "ref_quant4_fp16"), "oracle/_ref/libref_oracle.so not built (make -C oracle ref)"
    K, N = w_kn.shape
    w_kn = w_kn.contiguous()
    unp = torch.empty(K, N // 2, dtype=torch.int8)
    pro = torch.empty(K, N // 2, dtype=torch.int8)
    sc = torch.empty(N, dtype=w_kn.dtype)
    fn = lib.ref_quant4_fp16 if w_kn.dtype == torch.float16 else lib.ref_quant4_fp32
    rc = fn(_ptr(pro), _ptr(unp), _ptr(sc), _ptr(w_kn), ctypes.c_size_t(K), ctypes.c_size_t(N))
    assert rc == 0, "reference quantiser threw"
    return unp, pro, sc


def ref_preprocess4(p_kn2: torch.Tensor) -> torch.Tensor:
    """Run the REFERENCE ``preprocess_weights(is_int4=True)`` (arch 80) on packed int4 ``[K, N/2]`` (K x N elements)."""
    lib = ref_lib()
    assert lib is not None and hasattr(lib, "ref_preprocess4"), "oracle/_ref/libref_oracle.so not built (make -C oracle ref)"
    K, N2 = p_kn2.shape
    p_kn2 = p_kn2.contiguous()
    out = torch.empty(K, N2, dtype=torch.int8)
    rc = lib.ref_preprocess4(_ptr(out), _ptr(p_kn2), ctypes.c_size_t(K), ctypes.c_size_t(2 * N2))
    assert rc == 0, "reference preprocessor threw"
    return out


def port_quantize(w_kn: torch.Tensor):
    """Scalar C restatement of the quantiser (w8a16_oracle.c). Returns (q, scales, s32)."""
    lib = port_lib()
    assert lib is not None, "oracle/libw8a16_oracle.so not built (make -C oracle port)"
    K, N = w_kn.shape
    w_kn = w_kn.contiguous()
    q = torch.empty(K, N, dtype=torch.int8)
    s32 = torch.empty(N, dtype=torch.float32)
    if w_kn.dtype == torch.float16:
        sc = torch.empty(N, dtype=torch.float16)
        lib.oracle_quantize_f16(_ptr(w_kn), ctypes.c_size_t(K), ctypes.c_size_t(N), _ptr(q), _ptr(sc), _ptr(s32))
    else:
        sc = s32
        lib.oracle_quantize_f32(_ptr(w_kn), ctypes.c_size_t(K), ctypes.c_size_t(N), _ptr(q), _ptr(s32))
    return q, sc, s32


def port_ref_layout(q_kn: torch.Tensor) -> torch.Tensor:
    lib = port_lib()
    assert lib is not None
    K, N = q_kn.shape
    out = torch.empty(K, N, dtype=torch.int8)
    rc = lib.oracle_ref_layout(_ptr(q_kn.contiguous()), ctypes.c_size_t(K), ctypes.c_size_t(N), _ptr(out))
    assert rc == 0
    return out


def port_gemm_f16(x: torch.Tensor, q_kn: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    lib = port_lib()
    assert lib is not None
    M, K = x.shape
    N = q_kn.shape[1]
    y = torch.empty(M, N, dtype=torch.float16)
    lib.oracle_gemm_f16(_ptr(x.contiguous()), _ptr(q_kn.contiguous()), _ptr(scales.contiguous()), _ptr(y),
                        ctypes.c_size_t(M), ctypes.c_size_t(N), ctypes.c_size_t(K))
    return y
