"""w4a16 kernel micro-benchmark (development tool): our int4 decode kernel, the widen + tcgen05 path, and the reference's Int4b
decode kernel rebuilt for sm_100a, timed like kbench.py (CUDA-graph replays cycling > 2x L2 of distinct weights)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import eetq_b200  # noqa: E402
from kbench import L2_BYTES, time_graph  # noqa: E402


def algo_bytes4(M, N, K):
    return K * N // 2 + 2 * N + 2 * M * K + 2 * M * N


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gemv.so")
    ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
    if ref is not None and not hasattr(ref, "ref_w4a16_gemv"):
        ref = None
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    out = []
    for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
        pool = max(2, (2 * L2_BYTES) // (K * N // 2) + 1)
        ws = [torch.randint(-128, 128, (K, N // 2), dtype=torch.int8, device=dev) for _ in range(pool)]
        sc = (torch.rand(N, device=dev) * 0.01).half()
        for M in (1, 2, 4, 16, 256):
            x = torch.randn(M, K, device=dev).half()
            med, best = time_graph([(lambda w=w: eetq_b200.w4_a16_gemm(x, w, sc)) for w in ws])
            r = dict(kernel="w4a16", K=K, N=N, M=M, us=round(med, 2), us_best=round(best, 2), gbs=round(algo_bytes4(M, N, K) / med / 1e3, 1))
            print(json.dumps(r), flush=True)
            out.append(r)
            if ref is not None and M <= 4:
                y = torch.empty(M, N, device=dev, dtype=torch.float16)
                call = lambda w: ref.ref_w4a16_gemv(vp(x), vp(w), vp(sc), vp(y), M, N, K, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                med, best = time_graph([(lambda w=w: call(w)) for w in ws])
                r = dict(kernel="reference_int4_gemv_sm100a", K=K, N=N, M=M, us=round(med, 2), us_best=round(best, 2),
                         gbs=round(algo_bytes4(M, N, K) / med / 1e3, 1))
                print(json.dumps(r), flush=True)
                out.append(r)
        del ws
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kbench_int4.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
