"""``eet_quantize`` — swap every ``nn.Linear`` of a model (``lm_head`` excepted) for a :class:`W8A16Linear`.

Same call signature and behaviour as /root/reference/python/eetq/utils/quantizer.py:40-61; the per-layer quantisation
itself runs on the GPU (the reference spends ~90 s of single-threaded host time on Llama-2-7B, SURVEY.md section 3A;
here all 224 linears take ~0.05 s on one B200).
"""
from __future__ import annotations

import torch
from torch import nn

from ..modules.qlinear import W4A16Linear, W8A16Linear
from .base import find_layers, find_submodule, get_named_linears, set_op_by_name

__all__ = ["eet_quantize", "replace_with_eet_qlinear", "structure_mapping"]

# name of the decoder-layer container per model family (python/eetq/utils/mapping.py:1-18)
_DECODER_CONTAINER = {"llama": "layers", "baichuan": "layers"}


def structure_mapping(model: nn.Module, target_model: str = "llama") -> dict:
    """``{"decoder": <attribute name of the ModuleList of decoder layers>}`` for a supported family."""
    try:
        return {"decoder": _DECODER_CONTAINER[target_model]}
    except KeyError:
        raise NotImplementedError(f"structure_mapping: unsupported model family {target_model!r}") from None


def _to_w8a16(linear: nn.Module, init_only: bool, bits: int = 8) -> nn.Module:
    wdtype = linear.weight.dtype
    if bits == 4:  # our extension: packed int4 (the reference's eet_quantize only knows int8)
        if wdtype not in (torch.float16, torch.bfloat16, torch.float32):
            raise ValueError("Unsupported data type: {}".format(wdtype))
        return W4A16Linear.from_torch(linear, init_only=init_only)
    if bits != 8:
        raise ValueError(f"eet_quantize: bits must be 8 or 4 (got {bits})")
    if wdtype == torch.int8:
        # bitsandbytes.nn.Linear8bitLt keeps per-output-row abs-max in SCB; its int8 code is w / SCB * 127
        scales = linear.state_dict()["SCB"] / 127.0
        return W8A16Linear.from_torch(linear, scales=scales, init_only=init_only)
    if wdtype in (torch.float16, torch.bfloat16, torch.float32):
        return W8A16Linear.from_torch(linear, scales=None, init_only=init_only)
    raise ValueError("Unsupported data type: {}".format(wdtype))


def eet_quantize(model: nn.Module, init_only: bool = False, include=(nn.Linear,), exclude=("lm_head",), device="cuda:0",
                 verbose: bool = False, bits: int = 8) -> nn.Module:
    """Quantise ``model`` in place and return it.  ``init_only=True`` builds the quantised skeleton without touching the
    weights (for loading an already-quantised checkpoint).  ``device`` is accepted for signature compatibility; each layer
    stays on the device its weight lives on.  ``bits=4`` (extension) swaps in :class:`W4A16Linear` instead."""
    targets = find_layers(model, include=include, exclude=exclude)
    for dotted_name, linear in targets.items():
        set_op_by_name(model, dotted_name, _to_w8a16(linear, init_only, bits))
        if verbose:
            print("[EET][INFO] quantized {}".format(dotted_name))
    return model


def replace_with_eet_qlinear(model: nn.Module, init_only: bool = False, target_model: str = "llama", device="cuda:0") -> None:
    """Per-decoder-layer variant of :func:`eet_quantize` (python/eetq/utils/quantizer.py:13-38): every ``nn.Linear`` inside the
    decoder layers of ``model`` (found through :func:`structure_mapping`) becomes a :class:`W8A16Linear`; modules outside
    the decoder stack (embeddings, ``lm_head``) are left alone."""
    layers = find_submodule(model, structure_mapping(model, target_model)["decoder"])
    for layer in layers:
        for dotted_name, linear in get_named_linears(layer).items():
            set_op_by_name(layer, dotted_name, _to_w8a16(linear, init_only))
