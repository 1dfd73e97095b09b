"""A/B micro-benchmark of the decode-row kernels (development tool): SIMT / token-major mma.sync / weights-in-A mma.sync (v2) for
int8 and int4 weights, plus the reference kernels rebuilt for sm_100a; timed like kbench.py."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import eetq_b200  # noqa: E402
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402
from kbench import L2_BYTES, time_graph  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gemv.so")
    ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    cur = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = []

    def rec(**r):
        print(json.dumps(r), flush=True)
        out.append(r)

    only_simt4 = os.environ.get("KBENCH_ONLY_INT4_SIMT") == "1"
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        for bits in ((4,) if only_simt4 else (8, 4)):
            wbytes = K * N * bits // 8
            pool = max(2, (2 * L2_BYTES) // wbytes + 1)
            ws = [torch.randint(-128, 128, (K, N * bits // 8), dtype=torch.int8, device=dev) for _ in range(pool)]
            sc = (torch.rand(N, device=dev) * 0.01).half()
            for M in ((1, 2) if only_simt4 else (1, 2, 3, 4, 8)):
                x = torch.randn(M, K, device=dev).half()
                algo = wbytes + 2 * N + 2 * M * K + 2 * M * N
                for pdl in (False, True):
                    p = _cabi.FLAG_PDL if pdl else 0
                    variants = {}
                    if bits == 8:
                        variants["default"] = lambda w: w8_a16_gemm_bias(x, w, sc, None, flags=p)
                        # FORCE_GEMV = the streaming kernels: SIMT for 1-2 rows, the mma.sync kernel from 3 rows on (same as "mma2" there)
                        variants["simt"] = lambda w: w8_a16_gemm_bias(x, w, sc, None, flags=p | _cabi.FLAG_FORCE_GEMV)
                        variants["mma2"] = lambda w: w8_a16_gemm_bias(x, w, sc, None, flags=p | _cabi.FLAG_FORCE_MMA)
                    else:
                        if M <= 4:
                            variants["simt"] = lambda w: eetq_b200.w4_a16_gemm(x, w, sc, flags=p | _cabi.FLAG_FORCE_GEMV)
                        if not only_simt4:
                            variants["mma2"] = lambda w: eetq_b200.w4_a16_gemm(x, w, sc, flags=p | _cabi.FLAG_FORCE_MMA)
                    for name, fn in variants.items():
                        med, best = time_graph([(lambda w=w, fn=fn: fn(w)) for w in ws])
                        rec(bits=bits, K=K, N=N, M=M, pdl=pdl, kernel=name, us=round(med, 2), us_best=round(best, 2), gbs=round(algo / med / 1e3, 1))
                if ref is not None and M <= 4 and not only_simt4:
                    y = torch.empty(M, N, device=dev, dtype=torch.float16)
                    f = ref.ref_w8a16_gemv if bits == 8 else getattr(ref, "ref_w4a16_gemv", None)
                    if f is not None:
                        med, best = time_graph([(lambda w=w: f(vp(x), vp(w), vp(sc), vp(y), M, N, K, cur())) for w in ws])
                        rec(bits=bits, K=K, N=N, M=M, pdl=False, kernel="reference_sm100a", us=round(med, 2), us_best=round(best, 2),
                            gbs=round(algo / med / 1e3, 1))
            del ws
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kbench_mma2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
