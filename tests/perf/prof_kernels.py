"""Launch a fixed set of kernels inside a cudaProfilerStart/Stop range for `ncu --profile-from-start off`."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gemv.so")
ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None

cases = []
for (K, N) in [(4096, 4096), (4096, 12288), (4096, 22016), (11008, 4096)]:
    w = torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev)
    sc = (torch.rand(N, device=dev) * 0.01).half()
    cases.append((K, N, w, sc))

def run_all():
    for (K, N, w, sc) in cases:
        x = torch.randn(1, K, device=dev).half()
        w8_a16_gemm_bias(x, w, sc, None, flags=_cabi.FLAG_FORCE_GEMV)
        if ref is not None:
            y = torch.empty(1, N, device=dev, dtype=torch.float16)
            ref.ref_w8a16_gemv(vp(x), vp(w), vp(sc), vp(y), 1, N, K, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    K, N, w, sc = cases[0]
    for M in (16, 64, 256, 1024):
        x = torch.randn(M, K, device=dev).half()
        w8_a16_gemm_bias(x, w, sc, None, flags=_cabi.FLAG_FORCE_TC)
    torch.cuda.synchronize()

run_all()            # warm-up (attributes, workspace)
torch.cuda.profiler.start()
run_all()
torch.cuda.profiler.stop()
print("done")
