"""Drop-in stand-in for the reference's native extension module ``EETQ`` (setup.py:100, csrc/eetpy.cpp:7-19):

    from EETQ import quant_weights, preprocess_weights, w8_a16_gemm, w8_a16_gemm_, rotary_embedding_neox, layernorm_forward

resolves to the B200 implementations in :mod:`eetq_b200.ops` -- all six symbols the reference module exports.
"""
from eetq_b200.ops import (layernorm_forward, preprocess_weights, quant_weights, rotary_embedding_neox, w8_a16_gemm,  # noqa: F401
                           w8_a16_gemm_)

__all__ = ["quant_weights", "preprocess_weights", "w8_a16_gemm", "w8_a16_gemm_", "rotary_embedding_neox", "layernorm_forward"]
