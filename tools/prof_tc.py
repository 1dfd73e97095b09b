"""ncu target: the tcgen05 kernel on Llama-7B shapes at M = 16, 64, 256, 1024 (cudaProfilerStart/Stop range)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi
from eetq_b200.ops import w8_a16_gemm_bias
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
cases = []
for (K, N, M) in [(4096, 4096, 16), (4096, 11008, 64), (4096, 4096, 256), (4096, 4096, 1024), (11008, 4096, 1024)]:
    w = torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev)
    sc = (torch.rand(N, device=dev) * 0.01).half()
    x = torch.randn(M, K, device=dev).half()
    cases.append((x, w, sc))
def run():
    for x, w, sc in cases:
        w8_a16_gemm_bias(x, w, sc, None, flags=_cabi.FLAG_FORCE_TC)
    torch.cuda.synchronize()
run()
torch.cuda.profiler.start(); run(); torch.cuda.profiler.stop()
print("done")
