"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference quantiser/preprocessor
(oracle/_ref/libref_oracle.so, built by `make -C oracle ref` from /root/reference) on seeded inputs.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The fixtures travel to the GPU box with the repository; nothing on that box reads /root/reference.

Each .npz holds: w (the input, fp16 or fp32 bits), q (reference `unprocessed` int8 [K,N]), w_ref (reference
`processed` bytes, sm80 interleaved layout), scales (reference scales, input dtype), plus x / y for the GEMM
arithmetic cases (y from the oracle restatement on the REFERENCE's q and scales -- the reference GEMM itself
cannot execute without an sm70..sm89 GPU).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import w8a16_oracle as o  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.float16:
        return t.view(torch.int16).numpy()
    return t.numpy()


def main():
    assert o.ref_lib() is not None, "build oracle/_ref first: make -C oracle ref"
    cases = [
        ("q_f16_128x64_s0", 128, 64, 0, torch.float16, "randn"),
        ("q_f16_256x128_s1", 256, 128, 1, torch.float16, "randn"),
        ("q_f16_192x320_s2", 192, 320, 2, torch.float16, "randn"),
        ("q_f32_128x64_s3", 128, 64, 3, torch.float32, "randn"),
        ("q_f16_128x128_rand_s1", 128, 128, 1, torch.float16, "rand"),     # U[0,1) like examples/layers/test_w8a16_gemm.py:22
        ("q_f16_128x64_edge", 128, 64, 4, torch.float16, "edge"),           # zero column, +-max, ties
    ]
    for name, K, N, seed, dtype, kind in cases:
        g = torch.Generator().manual_seed(seed)
        if kind == "randn":
            w = (torch.randn(K, N, generator=g) * 0.02).to(dtype)
        elif kind == "rand":
            w = torch.rand(K, N, generator=g).to(dtype)
        else:
            w = (torch.randn(K, N, generator=g) * 0.02).to(dtype)
            w[:, 5] = 0                      # all-zero column -> NaN path (q = 127, scale 0)
            w[:, 6] = 0.5; w[3, 6] = -0.5    # every entry at +-amax -> 127 / -128 clamp
            w[:, 7] = 0.0; w[0, 7] = 1.0; w[1, 7] = 0.5 * (1.0 / 128.0) * 3  # exact .5 ties: 1.5 -> 2 (half away)
            w[2, 7] = -0.5 * (1.0 / 128.0) * 5                                # -2.5 -> -3
        unp, pro, sc = o.ref_quantize(w)
        x = o.synth_act(3, K, seed=7 + seed)
        if dtype == torch.float16:
            y = o.gemm(x, unp, sc)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), w=bits(w), q=unp.numpy(), w_ref=pro.numpy(),
                                scales=bits(sc), x=bits(x), y=bits(y), dtype="float16")
        else:
            np.savez_compressed(os.path.join(OUT, name + ".npz"), w=bits(w), q=unp.numpy(), w_ref=pro.numpy(),
                                scales=bits(sc), dtype="float32")
        print(name, "ok")


if __name__ == "__main__":
    main()
