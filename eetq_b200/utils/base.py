"""Module find/replace helpers on the quantisation path (/root/reference/python/eetq/utils/base.py:25-38, 270-285).
The reference's offline TP re-pack helpers (base.py:132-250) exist only because its interleaved layout cannot be
sliced; the b200 layout's column shards are contiguous, so they are not needed."""
from __future__ import annotations

import torch.nn as nn

__all__ = ["find_submodule", "get_op_by_name", "set_op_by_name", "get_named_linears", "get_named_layers", "find_layers"]


def find_submodule(module, sub_name):
    if hasattr(module, sub_name):
        return getattr(module, sub_name)
    for _, child in module.named_children():
        try:
            return find_submodule(child, sub_name)
        except ValueError:
            continue
    raise ValueError(f"Cannot find submodule {sub_name} in module {type(module).__name__}")


def get_op_by_name(module, op_name):
    for name, m in module.named_modules():
        if name == op_name:
            return m
    raise ValueError(f"Cannot find op {op_name} in module {type(module).__name__}")


def set_op_by_name(layer, name, new_module):
    """Replace the sub-module at dotted path ``name`` (numeric components index into containers)."""
    parent = layer
    *path, leaf = name.split(".")
    for part in path:
        parent = parent[int(part)] if part.isdigit() else getattr(parent, part)
    if leaf.isdigit():
        parent[int(leaf)] = new_module
    else:
        setattr(parent, leaf, new_module)


def get_named_linears(module):
    return {name: m for name, m in module.named_modules() if isinstance(m, nn.Linear) and "lm_head" not in name}


def get_named_layers(module, layers=(nn.Linear,)):
    return {name: m for name, m in module.named_modules() if type(m) in layers}


def find_layers(module, include=(nn.Linear,), exclude=("lm_head",)):
    """Exact-type match on ``include`` and substring match on ``exclude`` (base.py:280-285)."""
    return {name: m for name, m in module.named_modules()
            if type(m) in tuple(include) and not any(e in name for e in exclude)}
