"""Kernel micro-benchmark (development tool; bench.py is the contract benchmark).

Times each kernel with CUDA events around CUDA-graph replays of launches that cycle through a pool of distinct
weight buffers larger than 2x L2 (so "GB/s" is HBM, not L2), and reports algorithmic GB/s / TFLOP/s.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import eetq_b200  # noqa: E402
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402

L2_BYTES = 128 << 20


def algo_bytes(M, N, K):
    return K * N + 2 * N + 2 * M * K + 2 * M * N


def time_graph(fn_list, reps=20, inner=1):
    """fn_list: callables each enqueueing one launch; captured back-to-back into one graph."""
    for f in fn_list[:2]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            for f in fn_list:
                f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / (len(fn_list) * inner))  # us per launch
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kbench.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--tc-only", action="store_true")
    ap.add_argument("--gemv-only", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _cabi.lib()
    results = []
    shapes = [(4096, 4096), (4096, 11008), (11008, 4096)] if args.tc_only else [(4096, 4096), (4096, 11008), (11008, 4096), (4096, 12288), (4096, 22016)]

    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gemv.so")
    ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    for (K, N) in shapes:
        pool = max(2, (2 * L2_BYTES) // (K * N) + 1)
        ws = [torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev) for _ in range(pool)]
        sc = (torch.rand(N, device=dev) * 0.01).half()
        for M in ([] if args.tc_only else [1, 2, 3, 4] if not args.quick else [1]):
            x = torch.randn(M, K, device=dev).half()
            for mode, pdl in ((1, False), (1, True)):
                flags = _cabi.FLAG_FORCE_GEMV | (_cabi.FLAG_PDL if pdl else 0)
                fns = [(lambda w=w: w8_a16_gemm_bias(x, w, sc, None, flags=flags)) for w in ws]
                med, best = time_graph(fns)
                r = dict(kernel="gemv", mma=os.environ.get("EETQ_B200_GEMV_MMA", "1") != "0" and M >= 2, K=K, N=N, M=M, pdl=pdl, us=med, us_best=best,
                         gbs=algo_bytes(M, N, K) / med / 1e3)
                print(json.dumps(r), flush=True)
                results.append(r)
            if ref is not None:
                y = torch.empty(M, N, device=dev, dtype=torch.float16)
                st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                def ref_call(w):
                    ref.ref_w8a16_gemv(vp(x), vp(w), vp(sc), vp(y), M, N, K, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                fns = [(lambda w=w: ref_call(w)) for w in ws]
                med, best = time_graph(fns)
                r = dict(kernel="reference_gemv_sm100a", K=K, N=N, M=M, us=med, us_best=best, gbs=algo_bytes(M, N, K) / med / 1e3)
                print(json.dumps(r), flush=True)
                results.append(r)
        # torch fp16 GEMV/GEMM on dequantised weights for context (2 bytes/weight)
        if not args.tc_only:
            wf = [torch.randn(K, N, device=dev).half() for _ in range(max(2, pool // 2))]
            x = torch.randn(1, K, device=dev).half()
            fns = [(lambda w=w: torch.matmul(x, w)) for w in wf]
            med, best = time_graph(fns)
            r = dict(kernel="torch_fp16_matmul", K=K, N=N, M=1, us=med, gbs_fp16=(2 * K * N) / med / 1e3)
            print(json.dumps(r), flush=True)
            results.append(r)
            del wf
        for M in ([] if args.gemv_only else [8, 16, 64, 256, 1024] if not args.quick else [16, 256, 1024]):
            x = torch.randn(M, K, device=dev).half()
            tc_pdl = os.environ.get("KBENCH_TC_PDL", "0") == "1"
            fns = [(lambda w=w: w8_a16_gemm_bias(x, w, sc, None, flags=_cabi.FLAG_FORCE_TC | (_cabi.FLAG_PDL if tc_pdl else 0))) for w in ws]
            med, best = time_graph(fns)
            r = dict(kernel="gemm_tc", impl=os.environ.get("EETQ_B200_TC_IMPL", "v2"), dqw=os.environ.get("EETQ_B200_TC_DQW", "8"),
                     l2promo=os.environ.get("EETQ_B200_TC_L2PROMO", "256"), pdl=bool(tc_pdl), K=K, N=N, M=M, us=med, us_best=best,
                     gbs=algo_bytes(M, N, K) / med / 1e3, tflops=2.0 * M * N * K / med / 1e6)
            print(json.dumps(r), flush=True)
            results.append(r)
        del ws
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
