"""Module surface of the w8a16 hot path on top of :mod:`eetq_b200.ops`.

Mirrors — by name, buffer layout, state-dict keys and call semantics — the reference module file
/root/reference/python/eetq/modules/qlinear.py:
  quantize_and_preprocess_weights (:14-24), W8A16Linear (:27-62), EetqLinearMMFunction (:64-94), EetqLinear (:96-124).

State-dict compatibility (so checkpoints and HF integrations line up key for key):
  W8A16Linear : ``qweight`` int8 [in, out] (kernel-layout bytes), ``weight_scales`` [out], ``bias`` [out] or absent
  EetqLinear  : ``weight``  int8 [in, out], ``weight_scales`` registered late through ``register_scale``, ``bias``

Checkpoint layout marker.  The int8 bytes of ``qweight`` / ``weight`` are in the b200 layout here but in the sm80 interleaved
layout in checkpoints written by reference EETQ / Hugging Face (python/eetq/models/base.py:108-146) -- same key, shape and
dtype, different meaning.  Every module therefore carries a persistent ``weight_layout`` buffer (value ``B200_LAYOUT``) that is
saved with it; ``load_state_dict`` converts the weight on the GPU (closed form of cutlass_preprocessors.cc:497-534) when the
marker is absent, i.e. when the state dict comes from a reference build, and refuses unknown marker values.  Reference
builds in turn reject b200 checkpoints (unexpected key) instead of silently mis-reading them; ``export_reference_state_dict``
writes the reference's bytes for them.

What differs from the reference: weights are quantised on the layer's own device by the GPU quantiser (the reference
round-trips through ``.cpu()``, :16), the bias add is fused into the kernel epilogue (the reference issues a second torch
kernel, :61 and :77), bf16 modules are accepted.  ``W8A16LoraLinear`` (:127-186) never calls ``nn.Module.__init__`` in the
reference and cannot be instantiated there; it is not mirrored.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn
from torch.autograd import Function

from ..ops import (convert_ref_checkpoint_weight, convert_ref_checkpoint_weight4, preprocess_weights, quant_weights, to_ref_checkpoint_weight,
                   to_ref_checkpoint_weight4, w4_a16_gemm, w8_a16_gemm, w8_a16_gemm_bias)

__all__ = ["quantize_and_preprocess_weights", "W8A16Linear", "W4A16Linear", "EetqLinearMMFunction", "EetqLinear", "B200_LAYOUT",
           "B200_LAYOUT_INT4", "export_reference_state_dict"]

B200_LAYOUT = 200        # value of the ``weight_layout`` marker: bytes are output-feature-major, biased
B200_LAYOUT_INT4 = 204   # same for packed int4: rows of K/2 bytes, nibbles interleaved inside every 32-bit word
_LAYOUT_KEY = "weight_layout"


class _LayoutAwareMixin:
    """``load_state_dict`` support shared by W8A16Linear / EetqLinear: convert reference-layout weights, check the marker."""

    _weight_key = "qweight"
    _layout_value = B200_LAYOUT

    # reference-layout bytes <-> this module's layout (GPU kernels; looked up at call time so that tests can stub them)
    def _from_ref(self, w: torch.Tensor) -> torch.Tensor:
        return convert_ref_checkpoint_weight(w)

    def _to_ref(self, w: torch.Tensor) -> torch.Tensor:
        return to_ref_checkpoint_weight(w)

    def _register_layout_marker(self, device) -> None:
        self.register_buffer(_LAYOUT_KEY, torch.tensor([self._layout_value], dtype=torch.int32, device=device))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        wkey, lkey = prefix + self._weight_key, prefix + _LAYOUT_KEY
        if wkey in state_dict:
            if lkey not in state_dict:
                # written by a reference build: sm80 interleaved bytes -> b200 layout (a GPU kernel; results return to the
                # tensor's own device)
                w = state_dict[wkey]
                if w.dtype != torch.int8 or w.dim() != 2:
                    error_msgs.append(f"{wkey}: expected a 2-D int8 tensor, got {tuple(w.shape)} {w.dtype}")
                else:
                    if not torch.cuda.is_available():
                        raise RuntimeError(f"{wkey} is in the reference layout; converting it needs a CUDA device "
                                           "(eetq_b200 has no CPU path)")
                    src = w if w.is_cuda else w.cuda()
                    state_dict = dict(state_dict)  # do not touch the caller's dict
                    state_dict[wkey] = self._from_ref(src.contiguous()).to(w.device)
                    state_dict[lkey] = torch.tensor([self._layout_value], dtype=torch.int32)
            else:
                marker = int(state_dict[lkey].reshape(-1)[0])
                if marker != self._layout_value:
                    error_msgs.append(f"{lkey} = {marker}: weight layout does not match this module (expected {self._layout_value}; "
                                      f"{B200_LAYOUT} = int8, {B200_LAYOUT_INT4} = packed int4)")
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def export_reference_state_dict(model: nn.Module) -> dict:
    """``model.state_dict()`` with every quantised weight converted back to the reference's interleaved bytes and the layout
    markers dropped: loadable by reference EETQ / Hugging Face ``EetqLinear`` (python/eetq/models/base.py:108-146)."""
    sd = dict(model.state_dict())
    for name, mod in model.named_modules():
        if isinstance(mod, _LayoutAwareMixin):
            prefix = name + "." if name else ""
            wkey = prefix + mod._weight_key
            w = sd[wkey]
            src = w if w.is_cuda else w.cuda()
            sd[wkey] = mod._to_ref(src.contiguous()).to(w.device)
            sd.pop(prefix + _LAYOUT_KEY, None)
    return sd

_FLOAT_WEIGHT_DTYPES = (torch.float16, torch.bfloat16, torch.float32)


def quantize_and_preprocess_weights(weight: torch.Tensor, scales: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``nn.Linear.weight`` (``[out, in]``) -> ``(int8 [in, out] in kernel layout, scales [out])``.

    Floating-point weights go through the per-output-channel quantiser; weights that are already int8 (the
    bitsandbytes ingest path of ``eet_quantize``) are only re-laid-out and must come with their ``scales``.
    """
    kn = weight.t().contiguous()  # the native side works on [K = in, N = out]
    if kn.dtype in _FLOAT_WEIGHT_DTYPES:
        packed, computed_scales = quant_weights(kn, torch.int8, False)
        return packed, computed_scales
    if kn.dtype == torch.int8:
        if scales is None:
            raise AssertionError("int8 weights need their scales")  # the reference asserts the same (:19)
        return preprocess_weights(kn), scales
    raise ValueError("Unsupported data type: {}".format(weight.dtype))


def _activation_dtype_for(weight_dtype: torch.dtype) -> torch.dtype:
    return weight_dtype if weight_dtype in (torch.float16, torch.bfloat16) else torch.float16


class W8A16Linear(_LayoutAwareMixin, nn.Module):
    """Weight-only int8 linear layer; drop-in for ``nn.Linear`` at inference time."""

    _weight_key = "qweight"

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dev="cuda:0", dtype: torch.dtype = torch.float16):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("qweight", torch.zeros(in_features, out_features, dtype=torch.int8, device=dev))
        self._register_layout_marker(dev)
        self.register_buffer("weight_scales", torch.zeros(out_features, dtype=dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=dev))
        else:
            self.bias = None

    @classmethod
    def from_torch(cls, linear: nn.Module, scales: Optional[torch.Tensor] = None, init_only: bool = False) -> "W8A16Linear":
        """Build from an ``nn.Linear`` (or a bitsandbytes int8 linear plus ``scales``).  ``init_only`` only allocates the
        buffers so that a quantised checkpoint can be loaded into the skeleton (quantizer.py:40-45 in the reference)."""
        target = linear.weight.device
        act = _activation_dtype_for(linear.weight.dtype)
        layer = cls(linear.in_features, linear.out_features, bias=linear.bias is not None, dev=target, dtype=act)
        if init_only:
            return layer
        source = linear.weight.detach()
        if source.dtype != torch.int8 and not source.is_cuda and torch.cuda.is_available():
            source = source.cuda()  # the quantiser is a GPU kernel; results go back to `target` below
        packed, used_scales = quantize_and_preprocess_weights(source, scales)
        layer.qweight = packed.to(target)
        layer.weight_scales = used_scales.to(act).to(target)
        if linear.bias is not None:
            layer.bias = linear.bias.detach().to(act).clone()
        return layer

    @torch.no_grad()
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return w8_a16_gemm_bias(input, self.qweight, self.weight_scales, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class W4A16Linear(_LayoutAwareMixin, nn.Module):
    """Weight-only packed-int4 linear layer.  The reference has no int4 module: its Python reaches int4 only through
    ``quant_weights(w, torch.quint4x2)`` / ``preprocess_weights(w, is_int4=True)`` (csrc/eetpy.cpp:11-17) and its wrapper never selects
    the Int4b kernels it compiles (fpA_intB_gemm_wrapper.cu:154-159).  This module puts those two calls and ``w4_a16_gemm`` behind the
    ``W8A16Linear`` surface: ``qweight`` int8 ``[in, out/2]`` (two values per byte, b200 int4 layout), ``weight_scales`` ``[out]``,
    optional ``bias``; a state dict without the layout marker is taken to hold the reference's processed int4 bytes and converted."""

    _weight_key = "qweight"
    _layout_value = B200_LAYOUT_INT4

    def _from_ref(self, w: torch.Tensor) -> torch.Tensor:
        return convert_ref_checkpoint_weight4(w)

    def _to_ref(self, w: torch.Tensor) -> torch.Tensor:
        return to_ref_checkpoint_weight4(w)

    def __init__(self, in_features: int, out_features: int, bias: bool = True, dev="cuda:0", dtype: torch.dtype = torch.float16):
        super().__init__()
        if in_features % 64 or out_features % 64:
            raise ValueError("W4A16Linear needs in_features and out_features to be multiples of 64")
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("qweight", torch.zeros(in_features, out_features // 2, dtype=torch.int8, device=dev))
        self._register_layout_marker(dev)
        self.register_buffer("weight_scales", torch.zeros(out_features, dtype=dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=dev))
        else:
            self.bias = None

    @classmethod
    def from_torch(cls, linear: nn.Module, init_only: bool = False) -> "W4A16Linear":
        target = linear.weight.device
        act = _activation_dtype_for(linear.weight.dtype)
        layer = cls(linear.in_features, linear.out_features, bias=linear.bias is not None, dev=target, dtype=act)
        if init_only:
            return layer
        source = linear.weight.detach()
        if source.dtype not in _FLOAT_WEIGHT_DTYPES:
            raise ValueError("Unsupported data type: {}".format(source.dtype))
        if not source.is_cuda and torch.cuda.is_available():
            source = source.cuda()  # the quantiser is a GPU kernel; results go back to `target` below
        packed, scales = quant_weights(source.t().contiguous(), torch.quint4x2, False)
        layer.qweight = packed.to(target)
        layer.weight_scales = scales.to(act).to(target)
        if linear.bias is not None:
            layer.bias = linear.bias.detach().to(act).clone()
        return layer

    @torch.no_grad()
    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return w4_a16_gemm(input, self.qweight, self.weight_scales, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, bits=4"


class EetqLinearMMFunction(Function):
    """Autograd wrapper: forward = w8a16 GEMM (+bias); backward = ``grad_out @ dequant(W)^T`` where the dequantised
    weight is obtained with the identity-GEMM trick the reference uses, ``w8_a16_gemm(eye(K), W, s)``."""

    @staticmethod
    def forward(ctx, x, weight, scales, bias=None):
        ctx.save_for_backward(x, weight, scales, bias)
        return w8_a16_gemm_bias(x, weight, scales, bias)

    @staticmethod
    def backward(ctx, grad_output):
        x, weight, scales, _bias = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        k = weight.shape[0]
        dequantised = w8_a16_gemm(torch.eye(k, device=weight.device, dtype=x.dtype), weight, scales)  # [K, N]
        return grad_output.matmul(dequantised.t()), None, None, None


class EetqLinear(_LayoutAwareMixin, nn.Module):
    """The layer Hugging Face transformers instantiates for ``quant_method="eetq"``: int8 ``weight`` now, scales later."""

    _weight_key = "weight"

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device="cuda:0", dtype: torch.dtype = torch.float16):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self._act_dtype = dtype
        self.register_buffer("weight", torch.zeros(in_features, out_features, dtype=torch.int8, device=device))
        self._register_layout_marker(device)
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=device))
        else:
            self.bias = None

    def register(self, buffer_name: str, tensor: torch.Tensor) -> None:
        self.register_buffer(buffer_name, tensor)

    def register_scale(self, device) -> None:
        n = self.weight.shape[-1]
        self.register_buffer("weight_scales", torch.zeros(n, dtype=self._act_dtype, device=device))

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        args = (input, self.weight, self.weight_scales, self.bias)
        if self.training:
            return EetqLinearMMFunction.apply(*args)
        with torch.no_grad():
            return EetqLinearMMFunction.apply(*args)
