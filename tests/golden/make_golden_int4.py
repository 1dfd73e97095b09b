"""Generate the packed-int4 golden fixtures under tests/golden/int4/ by running the UNMODIFIED reference quantiser /
preprocessor (oracle/_ref/libref_oracle.so, `make -C oracle ref`) with QuantType::PACKED_INT4_WEIGHT_ONLY on seeded inputs.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden_int4.py

Each .npz holds: w (input bits), q4 (reference `unprocessed`: packed row-major [K, N/2]), w4_ref (reference `processed`
bytes, sm80 int4 layout), scales (reference scales), and for fp16 cases x / y (y from the oracle GEMM restatement on the
REFERENCE's q and scales: the reference's int4 kernels are not selectable from its Python and its CUTLASS GEMM cannot
execute without an sm70..sm89 GPU).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import w8a16_oracle as o  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "int4")


def bits(t: torch.Tensor) -> np.ndarray:
    return t.view(torch.int16).numpy() if t.dtype == torch.float16 else t.numpy()


def main():
    assert o.ref_lib() is not None and hasattr(o.ref_lib(), "ref_quant4_fp16"), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    cases = [
        ("q4_f16_128x64_s0", 128, 64, 0, torch.float16, "randn"),
        ("q4_f16_256x128_s1", 256, 128, 1, torch.float16, "randn"),
        ("q4_f16_192x320_s2", 192, 320, 2, torch.float16, "randn"),
        ("q4_f32_128x64_s3", 128, 64, 3, torch.float32, "randn"),
        ("q4_f16_128x64_edge", 128, 64, 4, torch.float16, "edge"),
    ]
    for name, K, N, seed, dtype, kind in cases:
        g = torch.Generator().manual_seed(seed)
        w = (torch.randn(K, N, generator=g) * 0.02).to(dtype)
        if kind == "edge":
            w[:, 5] = 0                       # all-zero column: 0/0 -> NaN -> int(NaN) = INT_MIN -> -8, scale 0
            w[:, 6] = 0.5; w[3, 6] = -0.5     # every entry at +-amax: +8 clamps to 7, -8 stays
            w[:, 7] = 0.0; w[0, 7] = 1.0; w[1, 7] = 1.5 / 8.0; w[2, 7] = -2.5 / 8.0   # exact ties: 1.5 -> 2, -2.5 -> -3
            w[9, 11] = float("nan")           # NaN weight: ignored by the abs-max, quantises to -8
        unp, pro, sc = o.ref_quantize4(w)
        rec = dict(w=bits(w), q4=unp.numpy(), w4_ref=pro.numpy(), scales=bits(sc), dtype=str(dtype).replace("torch.", ""))
        if dtype == torch.float16 and kind != "edge":
            x = o.synth_act(3, K, seed=7 + seed)
            rec.update(x=bits(x), y=bits(o.gemm(x, o.unpack_int4(unp), sc)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "ok")


if __name__ == "__main__":
    main()
