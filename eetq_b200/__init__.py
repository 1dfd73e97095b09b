"""eetq_b200 -- a B200-native (sm_100a) implementation of EETQ's w8a16 weight-only GEMM hot path.

Python surface mirrors the reference package ``eetq`` for that path
(/root/reference/python/eetq/__init__.py:1-3): modules (W8A16Linear, EetqLinear, EetqLinearMMFunction),
utils (eet_quantize, find_layers, set_op_by_name) and the native callables (quant_weights, preprocess_weights,
w8_a16_gemm, w8_a16_gemm_) that the reference keeps in its C++ extension ``EETQ``.
"""
from .modules import *  # noqa: F401,F403
from .ops import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403

__version__ = "0.1.0"
