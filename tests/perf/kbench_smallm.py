"""Where should the dispatch cut between the SIMT streaming kernel and the tcgen05 kernel sit?  Times M = 4..16."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi
from eetq_b200.ops import w8_a16_gemm_bias
from kbench import time_graph, algo_bytes, L2_BYTES
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
for (K, N) in [(4096, 4096), (4096, 11008), (11008, 4096)]:
    pool = max(2, (2 * L2_BYTES) // (K * N) + 1)
    ws = [torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev) for _ in range(pool)]
    sc = (torch.rand(N, device=dev) * 0.01).half()
    for M in (4, 5, 6, 8, 12, 16):
        x = torch.randn(M, K, device=dev).half()
        row = dict(K=K, N=N, M=M)
        for name, flag in (("gemv", _cabi.FLAG_FORCE_GEMV), ("tc", _cabi.FLAG_FORCE_TC)):
            if name == "gemv" and M > 8:
                continue
            fns = [(lambda w=w: w8_a16_gemm_bias(x, w, sc, None, flags=flag | _cabi.FLAG_PDL)) for w in ws]
            med, best = time_graph(fns)
            row[name + "_us"] = round(med, 2)
            row[name + "_gbs"] = round(algo_bytes(M, N, K) / med / 1e3, 1)
        print(json.dumps(row), flush=True)
    del ws
