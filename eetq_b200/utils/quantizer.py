"""``eet_quantize``: walk a model, swap every ``nn.Linear`` (except ``lm_head``) for a ``W8A16Linear``
(/root/reference/python/eetq/utils/quantizer.py:40-61).  Quantisation runs on the GPU (the reference spends ~90 s
of single-threaded host time on Llama-2-7B)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..modules.qlinear import W8A16Linear
from .base import find_layers, set_op_by_name

__all__ = ["eet_quantize"]


def eet_quantize(model, init_only=False, include=(nn.Linear,), exclude=("lm_head",), device="cuda:0", verbose=False):
    named_linears = find_layers(model, include=include, exclude=exclude)
    for name, linear in named_linears.items():
        if linear.weight.dtype in (torch.float16, torch.bfloat16, torch.float32):  # nn.Linear
            q_linear = W8A16Linear.from_torch(linear, scales=None, init_only=init_only)
        elif linear.weight.dtype == torch.int8:  # bitsandbytes.nn.Linear8bitLt: per-row absmax in SCB
            scales = torch.div(linear.state_dict()["SCB"], 127.0)
            q_linear = W8A16Linear.from_torch(linear, scales=scales, init_only=init_only)
        else:
            raise ValueError("Unsupported data type: {}".format(linear.weight.dtype))
        set_op_by_name(model, name, q_linear)
        if verbose:
            print("[EET][INFO] quantized {}".format(name))
        del linear
    return model
