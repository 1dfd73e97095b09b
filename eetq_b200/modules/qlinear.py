"""Module surface of the hot path: ``W8A16Linear``, ``EetqLinear``, ``EetqLinearMMFunction`` and
``quantize_and_preprocess_weights`` -- same names, buffers, state-dict keys and call semantics as
/root/reference/python/eetq/modules/qlinear.py:14-124, on top of :mod:`eetq_b200.ops`.

Buffers (identical names/shapes/dtypes to the reference, so state dicts line up key for key):
  W8A16Linear : qweight int8 [in, out] (kernel-layout bytes), weight_scales fp16 [out], bias fp16 [out] | None
  EetqLinear  : weight  int8 [in, out], weight_scales (registered late via register_scale), bias

Differences: quantisation happens on the layer's own device (no ``.cpu()`` round trip, qlinear.py:16); bias is
fused into the kernel epilogue instead of a second torch kernel (qlinear.py:61,77); bf16 modules are accepted.
``W8A16LoraLinear`` (qlinear.py:127-186) is broken in the reference (no ``super().__init__``) and is not mirrored.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from ..ops import preprocess_weights, quant_weights, w8_a16_gemm, w8_a16_gemm_bias

__all__ = ["quantize_and_preprocess_weights", "W8A16Linear", "EetqLinearMMFunction", "EetqLinear"]


def quantize_and_preprocess_weights(weight: torch.Tensor, scales: torch.Tensor = None):
    """``nn.Linear.weight`` ([out, in]) -> (kernel-layout int8 [in, out], scales [out]).  qlinear.py:14-24.

    fp16/bf16/fp32 weights are quantised per output channel; int8 weights (bitsandbytes ingest,
    utils/quantizer.py:46-48) are only re-laid-out and need ``scales``."""
    w_kn = torch.t(weight).contiguous()
    if w_kn.dtype == torch.int8:
        assert scales is not None  # need scales for real quantization
        return preprocess_weights(w_kn), scales
    if w_kn.dtype in (torch.float16, torch.bfloat16, torch.float32):
        qweight, scales = quant_weights(w_kn, torch.int8, False)
        return qweight, scales
    raise ValueError("Unsupported data type: {}".format(weight.dtype))


class W8A16Linear(nn.Module):
    def __init__(self, in_features, out_features, bias=True, dev="cuda:0", dtype=torch.float16):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.register_buffer("qweight", torch.zeros((in_features, out_features), dtype=torch.int8, device=dev))
        self.register_buffer("weight_scales", torch.zeros((out_features), dtype=dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features), dtype=dtype, device=dev))
        else:
            self.bias = None

    @classmethod
    def from_torch(cls, linear, scales=None, init_only=False):
        wdtype = linear.weight.dtype
        act_dtype = wdtype if wdtype in (torch.float16, torch.bfloat16) else torch.float16
        q = cls(linear.in_features, linear.out_features, bias=linear.bias is not None, dev=linear.weight.device,
                dtype=act_dtype)
        if init_only:  # just prepare for loading weights
            return q
        if linear.bias is not None:
            q.bias = linear.bias.detach().clone().to(act_dtype)
        weight = linear.weight.detach()
        if not weight.is_cuda and torch.cuda.is_available() and weight.dtype != torch.int8:
            weight = weight.cuda()  # quantise on the GPU; results are moved back to the layer's device below
        int8_weight, scales = quantize_and_preprocess_weights(weight, scales)
        q.qweight = int8_weight.to(linear.weight.device)
        q.weight_scales = scales.to(act_dtype).to(linear.weight.device)
        return q

    @torch.no_grad()
    def forward(self, input):
        return w8_a16_gemm_bias(input, self.qweight, self.weight_scales, self.bias)

    def extra_repr(self):
        return "in_features={}, out_features={}, bias={}".format(self.in_features, self.out_features, self.bias is not None)


class EetqLinearMMFunction(Function):
    """Autograd wrapper (qlinear.py:64-94).  Backward dequantises the weight with the identity-GEMM trick the
    reference uses (``w8_a16_gemm(eye(K), W, s)``) and returns ``grad_out @ W_dq^T``."""

    @staticmethod
    def forward(ctx, x, weight, scales, bias=None):
        ctx.save_for_backward(x, weight, scales, bias)
        return w8_a16_gemm_bias(x, weight, scales, bias)

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, scales, bias = ctx.saved_tensors
        grad_input = None
        if ctx.needs_input_grad[0]:
            identity = torch.eye(weight.shape[0], device=weight.device, dtype=input.dtype)
            w_dq = w8_a16_gemm(identity, weight, scales)  # [K, N]
            grad_input = grad_output.matmul(w_dq.transpose(0, 1))
        return grad_input, None, None, None


class EetqLinear(nn.Module):
    """The HF-transformers-facing module (buffer ``weight`` + late-registered ``weight_scales``), qlinear.py:96-124."""

    def __init__(self, in_features, out_features, bias=True, device="cuda:0", dtype=torch.float16):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.register_buffer("weight", torch.zeros((in_features, out_features), dtype=torch.int8, device=device))
        if bias:
            self.register_buffer("bias", torch.zeros((out_features), dtype=dtype, device=device))
        else:
            self.bias = None
        self._act_dtype = dtype

    def register(self, buffer_name, tensor):
        self.register_buffer(buffer_name, tensor)

    def register_scale(self, device):
        out_features = self.weight.shape[-1]
        self.register_buffer("weight_scales", torch.zeros((out_features), dtype=self._act_dtype, device=device))

    def forward(self, input):
        if self.training:
            return EetqLinearMMFunction.apply(input, self.weight, self.weight_scales, self.bias)
        with torch.no_grad():
            return EetqLinearMMFunction.apply(input, self.weight, self.weight_scales, self.bias)
