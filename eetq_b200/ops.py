"""The four hot-path callables of the reference's native module ``EETQ`` (csrc/eetpy.cpp:9-17), re-implemented on
top of libeetq_b200.so:

    quant_weights(origin_weight, quant_type, return_unprocessed_quantized_tensor=False)
    preprocess_weights(origin_weight, is_int4=False)
    w8_a16_gemm(input, weight, scale)
    w8_a16_gemm_(input, weight, scale, output, m, n, k)

Same names, arity, argument meaning and return order as the reference (fpA_intB_gemm_wrapper.cu:28-202).  The
"processed" weight tensor keeps the reference's nominal ``[K, N]`` int8 shape but its bytes are in the b200 layout
(output-feature-major), see DESIGN.md section 3.  Differences, all documented in INTEGRATION.md:

* the quantiser runs on the GPU (bit-exact with the reference's host loop); CPU inputs are accepted like the
  reference requires (results come back on the CPU), CUDA inputs are an extension (results stay on the device);
* ``w8_a16_gemm`` validates dtype / device / contiguity / shape agreement (the reference validates nothing) and
  accepts bf16 as well as fp16;
* packed int4 (``quant_weights(w, torch.quint4x2)`` / ``preprocess_weights(w, is_int4=True)``) is supported with the reference's
  shapes (``[K, N/2]`` int8, two values per byte) and bit-exact values; the reference has no Python-visible int4 GEMM
  (fpA_intB_gemm_wrapper.cu:154-159 hard-codes Int8b), ``w4_a16_gemm`` is this repository's addition.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import _cabi

__all__ = ["quant_weights", "preprocess_weights", "w8_a16_gemm", "w8_a16_gemm_", "w8_a16_gemm_bias", "w8_a16_gemm_residual",
           "convert_ref_checkpoint_weight", "to_ref_checkpoint_weight", "unpack_weights", "rotary_embedding_neox", "layernorm_forward",
           "w4_a16_gemm", "unpack_weights4", "convert_ref_checkpoint_weight4", "to_ref_checkpoint_weight4"]

_DTYPE_CODE = {torch.float16: _cabi.F16, torch.bfloat16: _cabi.BF16, torch.float32: _cabi.F32}


def _vp(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("eetq_b200 needs a CUDA (sm_100a) device; there is no CPU path")


def _default_device() -> torch.device:
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------------------------
# workspace for the split-K tcgen05 path: one zero-initialised buffer per (device, stream), grown on demand
# ---------------------------------------------------------------------------------------------------------------
_workspaces = {}
_retired_workspaces = []  # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def _workspace(device: torch.device, nbytes: int) -> Optional[torch.Tensor]:
    if nbytes == 0:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired_workspaces.append(ws)
        ws = torch.zeros(max(2 * nbytes, 1 << 24), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ---------------------------------------------------------------------------------------------------------------
# quant_weights  (symmetric_quantize_last_axis_of_tensor, fpA_intB_gemm_wrapper.cu:28-107)
# ---------------------------------------------------------------------------------------------------------------
def quant_weights(origin_weight: torch.Tensor, quant_type, return_unprocessed_quantized_tensor: bool = False) -> List[torch.Tensor]:
    w = origin_weight
    int4 = quant_type == getattr(torch, "quint4x2", None)
    if quant_type != torch.int8 and not int4:
        raise RuntimeError("Must be int4 or int8 quantization")  # same message as the reference (wrapper.cu:41)
    if w.numel() == 0:
        raise RuntimeError("weight should not be empty tensor")
    if w.dim() not in (2, 3):
        raise RuntimeError("Invalid dim. The dim of weight should be 2 or 3")
    if w.dtype not in (torch.float16, torch.float32, torch.bfloat16):
        raise RuntimeError("Invalid datatype. Weight must be FP16 or FP32")
    if not w.is_contiguous():
        raise RuntimeError("weight must be contiguous")
    on_cpu = not w.is_cuda
    dev = _default_device() if on_cpu else w.device
    K, N = w.shape[-2], w.shape[-1]
    if K % 64 or N % 64:
        raise RuntimeError(f"quant_weights: K ({K}) and N ({N}) must be multiples of 64")
    experts = 1 if w.dim() == 2 else w.shape[0]
    with torch.cuda.device(dev):
        wd = w.to(dev, non_blocking=False)
        # int4: two values per byte along the last axis, like the reference's [.., K, N/2] tensors (wrapper.cu:48-63)
        qshape = w.shape[:-1] + (N // 2,) if int4 else w.shape
        processed = torch.empty(qshape, dtype=torch.int8, device=dev)
        unprocessed = torch.empty(qshape, dtype=torch.int8, device=dev) if return_unprocessed_quantized_tensor else None
        scales = torch.empty(w.shape[:-2] + (N,), dtype=w.dtype, device=dev)
        s32 = torch.empty(experts, N, dtype=torch.float32, device=dev)
        L = _cabi.lib()
        for e in range(experts):  # 3-D input = one matrix per expert (cutlass_preprocessors.cc:614)
            we = wd if w.dim() == 2 else wd[e]
            pe = processed if w.dim() == 2 else processed[e]
            ue = None if unprocessed is None else (unprocessed if w.dim() == 2 else unprocessed[e])
            se = scales if w.dim() == 2 else scales[e]
            fn = L.eetq_b200_quantize4 if int4 else L.eetq_b200_quantize
            rc = fn(_vp(we), _DTYPE_CODE[w.dtype], K, N, _vp(pe), _vp(se), _vp(s32[e]), _vp(ue), _stream())
            _cabi.check(rc, "eetq_b200_quantize4" if int4 else "eetq_b200_quantize")
    if on_cpu:
        processed, scales = processed.cpu(), scales.cpu()
        unprocessed = None if unprocessed is None else unprocessed.cpu()
    if return_unprocessed_quantized_tensor:
        return [unprocessed, processed, scales]
    return [processed, scales]


# ---------------------------------------------------------------------------------------------------------------
# preprocess_weights  (preprocess_weights_cuda, fpA_intB_gemm_wrapper.cu:109-128)
# ---------------------------------------------------------------------------------------------------------------
def _layout_call(fn_name: str, t: torch.Tensor, int4: bool = False) -> torch.Tensor:
    if t.dtype not in (torch.int8, torch.uint8) or t.dim() != 2:
        raise RuntimeError(f"{fn_name}: expected a 2-D int8 tensor")
    on_cpu = not t.is_cuda
    dev = _default_device() if on_cpu else t.device
    K, N = t.shape
    if int4:
        N *= 2  # [K, N/2] bytes hold K x N values
    with torch.cuda.device(dev):
        src = t.contiguous().to(dev)
        dst = torch.empty_like(src)
        rc = getattr(_cabi.lib(), fn_name)(_vp(src), K, N, _vp(dst), _stream())
        _cabi.check(rc, fn_name)
    return dst.cpu() if on_cpu else dst


def preprocess_weights(origin_weight: torch.Tensor, is_int4: bool = False) -> torch.Tensor:
    """Row-major int8 ``[K, N]`` -> kernel layout (b200), returned with the same nominal shape like the reference."""
    if is_int4:
        # packed row-major [K, N/2] (low nibble = even column) -> b200 int4 layout.  The reference passes the BYTE column count
        # as the element count here (fpA_intB_gemm_wrapper.cu:121-126) and so rearranges only the first half of the buffer;
        # this converts the whole matrix
        return _layout_call("eetq_b200_pack4", origin_weight, int4=True)
    return _layout_call("eetq_b200_pack", origin_weight)


def unpack_weights(weight: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`preprocess_weights` (b200 layout -> row-major int8 ``[K, N]``)."""
    return _layout_call("eetq_b200_unpack", weight)


def unpack_weights4(weight: torch.Tensor) -> torch.Tensor:
    """b200 int4 layout ``[K, N/2]`` -> packed row-major ``[K, N/2]`` (two's-complement nibbles, low nibble = even column)."""
    return _layout_call("eetq_b200_unpack4", weight, int4=True)


def convert_ref_checkpoint_weight4(weight_ref: torch.Tensor) -> torch.Tensor:
    """The reference's processed int4 bytes (``quant_weights(w, torch.quint4x2)[0]`` of an EETQ build, sm80 layout,
    cutlass_preprocessors.cc:497-534) -> b200 int4 layout."""
    return _layout_call("eetq_b200_from_ref_layout4", weight_ref.view(torch.int8), int4=True)


def to_ref_checkpoint_weight4(weight: torch.Tensor) -> torch.Tensor:
    return _layout_call("eetq_b200_to_ref_layout4", weight, int4=True)


def convert_ref_checkpoint_weight(weight_ref: torch.Tensor) -> torch.Tensor:
    """Bytes saved by reference EETQ / HF ``EetqLinear.weight`` (sm80 interleaved layout,
    cutlass_preprocessors.cc:497-534) -> b200 layout."""
    out = _layout_call("eetq_b200_from_ref_layout", weight_ref.view(torch.int8))
    return out


def to_ref_checkpoint_weight(weight: torch.Tensor) -> torch.Tensor:
    """b200 layout -> the reference's interleaved bytes (for writing checkpoints other EETQ builds can load)."""
    return _layout_call("eetq_b200_to_ref_layout", weight)


# ---------------------------------------------------------------------------------------------------------------
# w8_a16_gemm / w8_a16_gemm_  (w8_a16_gemm_forward_cuda(_), fpA_intB_gemm_wrapper.cu:130-202)
# ---------------------------------------------------------------------------------------------------------------
def _check_gemm_args(x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor]):
    if not x.is_cuda:
        raise RuntimeError("w8_a16_gemm: input must be a CUDA tensor (eetq_b200 has no CPU path)")
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError(f"w8_a16_gemm: input dtype must be float16 or bfloat16, got {x.dtype}")
    if weight.dtype != torch.int8 or weight.dim() != 2:
        raise RuntimeError("w8_a16_gemm: weight must be a 2-D int8 tensor (output of quant_weights / preprocess_weights)")
    if weight.device != x.device or scale.device != x.device or (bias is not None and bias.device != x.device):
        raise RuntimeError("w8_a16_gemm: input, weight, scale (and bias) must be on the same device")
    if x.dim() < 1 or x.shape[-1] != weight.shape[0]:
        raise RuntimeError(f"w8_a16_gemm: input last dim {tuple(x.shape)} does not match weight K={weight.shape[0]}")
    if scale.numel() != weight.shape[1]:
        raise RuntimeError(f"w8_a16_gemm: scale has {scale.numel()} elements, expected N={weight.shape[1]}")
    if scale.dtype != x.dtype:
        raise RuntimeError(f"w8_a16_gemm: scale dtype {scale.dtype} must match input dtype {x.dtype}")
    if bias is not None and (bias.dtype != x.dtype or bias.numel() != weight.shape[1]):
        raise RuntimeError("w8_a16_gemm: bias must have N elements of the input dtype")
    if not weight.is_contiguous() or not scale.is_contiguous() or (bias is not None and not bias.is_contiguous()):
        raise RuntimeError("w8_a16_gemm: weight, scale and bias must be contiguous")


def _gemm_into(x2: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor], out2: torch.Tensor,
               M: int, N: int, K: int, flags: int = _cabi.FLAG_DEFAULT, residual: Optional[torch.Tensor] = None) -> None:
    L = _cabi.lib()
    ws_bytes = int(L.eetq_b200_workspace_bytes(M, N, K)) if M > 2 else 0  # 3 and 4 rows go to the tcgen05 kernel on the large shapes
    ws = _workspace(x2.device, ws_bytes)
    # a single row has no meaningful row stride (torch allows anything there): pass the contiguous value
    ldx = x2.stride(0) if M > 1 else K
    ldy = out2.stride(0) if M > 1 else N
    if residual is not None:
        rc = L.eetq_b200_w8a16_gemm_residual(_vp(x2), ldx, _vp(weight), _vp(scale), _vp(bias), _vp(residual),
                                             residual.stride(0) if M > 1 else N, _vp(out2), ldy, M, N, K, _DTYPE_CODE[x2.dtype], _vp(ws),
                                             0 if ws is None else ws.numel(), flags, _stream())
        _cabi.check(rc, "eetq_b200_w8a16_gemm_residual")
        return
    rc = L.eetq_b200_w8a16_gemm_ex(_vp(x2), ldx, _vp(weight), _vp(scale), _vp(bias), _vp(out2), ldy, M, N, K,
                                   _DTYPE_CODE[x2.dtype], _vp(ws), 0 if ws is None else ws.numel(), flags, _stream())
    _cabi.check(rc, "eetq_b200_w8a16_gemm")


def w8_a16_gemm_bias(input: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
                     flags: int = _cabi.FLAG_DEFAULT) -> torch.Tensor:
    """``w8_a16_gemm`` with the bias add fused into the kernel epilogue (the reference adds bias with a separate
    torch op, python/eetq/modules/qlinear.py:61)."""
    _check_gemm_args(input, weight, scale, bias)
    K, N = weight.shape
    x2 = input.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0) or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty(input.shape[:-1] + (N,), dtype=input.dtype, device=input.device)
    if M > 0:
        with torch.cuda.device(input.device):
            _gemm_into(x2, weight, scale, bias, out.view(-1, N), M, N, K, flags)
    return out


def w8_a16_gemm_residual(input: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, residual: torch.Tensor,
                         bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``residual + w8_a16_gemm(input, ...)`` with the add fused into the kernel epilogue (fp16 add of the rounded product,
    like ``hidden = residual + o_proj(x)``).  ``residual`` is ``[M, N]`` with unit inner stride (any 8-aligned row stride)."""
    _check_gemm_args(input, weight, scale, bias)
    K, N = weight.shape
    x2 = input.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0) or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    r2 = residual.reshape(-1, N)
    if r2.shape[0] != M or r2.dtype != input.dtype or r2.device != input.device:
        raise RuntimeError("w8_a16_gemm_residual: residual must be [M, N] of input's dtype on input's device")
    if r2.stride(-1) != 1 or (M > 1 and r2.stride(0) % 8 != 0) or r2.data_ptr() % 16 != 0:
        r2 = r2.contiguous()
    out = torch.empty(input.shape[:-1] + (N,), dtype=input.dtype, device=input.device)
    if M > 0:
        with torch.cuda.device(input.device):
            _gemm_into(x2, weight, scale, bias, out.view(-1, N), M, N, K, residual=r2)
    return out


def w4_a16_gemm(input: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
                flags: int = _cabi.FLAG_DEFAULT) -> torch.Tensor:
    """``y = input @ dequant(weight)`` for packed-int4 weights in the b200 int4 layout (``[K, N/2]`` int8 from
    ``quant_weights(w, torch.quint4x2)`` / ``preprocess_weights(w, is_int4=True)``).  Same arithmetic as ``w8_a16_gemm`` with q in
    [-8, 7]; what the reference's compiled-but-unselectable Int4b kernels compute (weightOnlyBatchedGemv/kernel.h:68-116)."""
    if weight.dtype != torch.int8 or weight.dim() != 2:
        raise RuntimeError("w4_a16_gemm: weight must be a 2-D int8 tensor [K, N/2]")
    K, N = weight.shape[0], weight.shape[1] * 2
    if not input.is_cuda:
        raise RuntimeError("w4_a16_gemm: input must be a CUDA tensor (eetq_b200 has no CPU path)")
    if input.dtype not in (torch.float16, torch.bfloat16) or scale.dtype != input.dtype or (bias is not None and bias.dtype != input.dtype):
        raise RuntimeError("w4_a16_gemm: input, scale (and bias) must share dtype float16 or bfloat16")
    if input.shape[-1] != K or scale.numel() != N or (bias is not None and bias.numel() != N):
        raise RuntimeError(f"w4_a16_gemm: shapes do not match K={K}, N={N}")
    if any(t.device != input.device for t in (weight, scale) + (() if bias is None else (bias,))):
        raise RuntimeError("w4_a16_gemm: all tensors must be on the same device")
    if not weight.is_contiguous() or not scale.is_contiguous() or (bias is not None and not bias.is_contiguous()):
        raise RuntimeError("w4_a16_gemm: weight, scale and bias must be contiguous")
    x2 = input.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0) or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty(input.shape[:-1] + (N,), dtype=input.dtype, device=input.device)
    if M > 0:
        with torch.cuda.device(input.device):
            L = _cabi.lib()
            ws = _workspace(input.device, int(L.eetq_b200_w4a16_workspace_bytes(M, N, K)))
            rc = L.eetq_b200_w4a16_gemm(_vp(x2), x2.stride(0) if M > 1 else K, _vp(weight), _vp(scale), _vp(bias), _vp(out), N, M, N, K,
                                        _DTYPE_CODE[input.dtype], _vp(ws), 0 if ws is None else ws.numel(), flags, _stream())
            _cabi.check(rc, "eetq_b200_w4a16_gemm")
    return out


def w8_a16_gemm(input: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """y = input @ dequant(weight) ; callee allocates the output on input's device with input's dtype
    (fpA_intB_gemm_wrapper.cu:139-140).  Any leading shape is accepted (the reference handles 2-D and 3-D)."""
    return w8_a16_gemm_bias(input, weight, scale, None)


def w8_a16_gemm_(input: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, output: torch.Tensor, m: int, n: int,
                 k: int) -> torch.Tensor:
    """In-place variant with a caller-owned output (fpA_intB_gemm_wrapper.cu:176-202)."""
    _check_gemm_args(input, weight, scale, None)
    if weight.shape[0] != k or weight.shape[1] != n or input.numel() != m * k or output.numel() != m * n:
        raise RuntimeError("w8_a16_gemm_: m, n, k do not match the tensor sizes")
    if output.dtype != input.dtype or output.device != input.device or not output.is_contiguous() or not input.is_contiguous():
        raise RuntimeError("w8_a16_gemm_: output must be a contiguous tensor of input's dtype on input's device")
    if m > 0:
        with torch.cuda.device(input.device):
            _gemm_into(input.view(m, k), weight, scale, None, output.view(m, n), m, n, k)
    return output


# ---------------------------------------------------------------------------------------------------------------
# the two glue ops the reference module also exports (csrc/eetpy.cpp:18-19)
# ---------------------------------------------------------------------------------------------------------------
def rotary_embedding_neox(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor, head_size: int,
                          cos_sin_cache: torch.Tensor) -> None:
    """In-place GPT-NeoX rotary embedding of ``query`` and ``key`` (``[..., num_heads, head_size]`` fp16, contiguous),
    ``positions`` int64 ``[num_tokens]``, ``cos_sin_cache`` ``[max_position, rot_dim]`` = cos | sin halves -- same
    arguments, in-place semantics and fp16 arithmetic as the reference (pos_encoding_kernels.cu:12-87)."""
    if query.dtype != torch.float16 or key.dtype != torch.float16 or cos_sin_cache.dtype != torch.float16:
        raise RuntimeError("rotary_embedding_neox: fp16 tensors expected")
    if not (query.is_cuda and key.is_cuda and positions.is_cuda and cos_sin_cache.is_cuda):
        raise RuntimeError("rotary_embedding_neox: CUDA tensors expected (eetq_b200 has no CPU path)")
    if not (query.is_contiguous() and key.is_contiguous() and cos_sin_cache.is_contiguous()) or positions.dtype != torch.int64:
        raise RuntimeError("rotary_embedding_neox: contiguous tensors and int64 positions expected")
    num_heads = query.shape[-2]
    num_tokens = query.numel() // (num_heads * head_size)
    if key.shape != query.shape or query.shape[-1] != head_size or positions.numel() != num_tokens:
        raise RuntimeError("rotary_embedding_neox: shape mismatch")
    with torch.cuda.device(query.device):
        rc = _cabi.lib().eetq_b200_rotary_embedding_neox(_vp(positions.contiguous()), _vp(query), _vp(key), num_tokens, num_heads, head_size,
                                                          _vp(cos_sin_cache), cos_sin_cache.shape[1], _stream())
        _cabi.check(rc, "eetq_b200_rotary_embedding_neox")


def layernorm_forward(input: torch.Tensor, gamma: torch.Tensor, out: torch.Tensor, eps: float) -> None:
    """T5-style RMS norm of the rows of ``input`` ``[b, n, c]`` fp16 into ``out`` (layernorm.cu:88-110)."""
    if input.dtype != torch.float16 or gamma.dtype != torch.float16 or out.dtype != torch.float16:
        raise RuntimeError("layernorm_forward: fp16 tensors expected")
    if not (input.is_cuda and gamma.is_cuda and out.is_cuda):
        raise RuntimeError("layernorm_forward: CUDA tensors expected (eetq_b200 has no CPU path)")
    if not (input.is_contiguous() and out.is_contiguous() and gamma.is_contiguous()) or out.shape != input.shape:
        raise RuntimeError("layernorm_forward: contiguous tensors of equal shape expected")
    n = input.shape[-1]
    m = input.numel() // n
    if gamma.numel() != n:
        raise RuntimeError("layernorm_forward: gamma must have input.shape[-1] elements")
    with torch.cuda.device(input.device):
        rc = _cabi.lib().eetq_b200_layernorm_forward(_vp(input), _vp(gamma), _vp(out), m, n, float(eps), _stream())
        _cabi.check(rc, "eetq_b200_layernorm_forward")
