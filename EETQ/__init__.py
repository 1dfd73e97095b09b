"""Drop-in stand-in for the reference's native extension module ``EETQ`` (setup.py:100, csrc/eetpy.cpp:7-19):

    from EETQ import quant_weights, preprocess_weights, w8_a16_gemm

resolves to the B200 implementations in :mod:`eetq_b200.ops`.  ``rotary_embedding_neox`` and ``layernorm_forward``
(eetpy.cpp:18-19) are outside the w8a16 hot path and are not provided.
"""
from eetq_b200.ops import preprocess_weights, quant_weights, w8_a16_gemm, w8_a16_gemm_  # noqa: F401

__all__ = ["quant_weights", "preprocess_weights", "w8_a16_gemm", "w8_a16_gemm_"]
