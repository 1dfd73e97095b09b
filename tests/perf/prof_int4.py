"""Decode-row kernels on 4096 x 11008 inside a cudaProfilerStart/Stop range for `ncu --profile-from-start off`:
int8 SIMT M=1, int4 SIMT M=1, reference Int4b M=1, int4 mma M=2, int8 mma M=4 (development tool)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import eetq_b200  # noqa: E402
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gemv.so")
ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
K, N = 4096, 11008
w8 = torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev)
w4 = torch.randint(-128, 128, (K, N // 2), dtype=torch.int8, device=dev)
sc = (torch.rand(N, device=dev) * 0.01).half()
xs = {m: torch.randn(m, K, device=dev).half() for m in (1, 2, 4)}
y = torch.empty(4, N, device=dev, dtype=torch.float16)


def run_all():
    w8_a16_gemm_bias(xs[1], w8, sc, None, flags=_cabi.FLAG_FORCE_GEMV)
    eetq_b200.w4_a16_gemm(xs[1], w4, sc, flags=_cabi.FLAG_FORCE_GEMV)
    if ref is not None and hasattr(ref, "ref_w4a16_gemv"):
        ref.ref_w4a16_gemv(vp(xs[1]), vp(w4), vp(sc), vp(y), 1, N, K, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    eetq_b200.w4_a16_gemm(xs[2], w4, sc, flags=_cabi.FLAG_FORCE_MMA)
    w8_a16_gemm_bias(xs[4], w8, sc, None, flags=_cabi.FLAG_FORCE_MMA)
    torch.cuda.synchronize()


run_all()
torch.cuda.profiler.start()
run_all()
torch.cuda.profiler.stop()
print("done")
