"""Time the 128 fused GEMV launches of one Llama-2-7B decode token as one CUDA graph (development tool).
Kernel-selection knobs are environment variables read by the library (EETQ_B200_GEMV_IMPL / _CTAS / _LOWREG)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
L = _cabi.lib()
H, I, LAYERS = 4096, 11008, 32
pdl = int(os.environ.get("CHAIN_PDL", "1"))
vp = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())


def mk(K, N):
    return (torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev), (torch.rand(N, device=dev) * 0.01).half())


layers = [dict(qkv=mk(H, 3 * H), o=mk(H, H), gu=mk(H, 2 * I), down=mk(I, H), ln=torch.ones(H, device=dev).half()) for _ in range(LAYERS)]
x = torch.randn(H, device=dev).half() * 0.1
x2 = torch.zeros(H, device=dev).half()
qkv = torch.zeros(3 * H, device=dev).half()
attn = torch.randn(H, device=dev).half() * 0.1
gu = torch.zeros(2 * I, device=dev).half()


def gemv(xin, ldx, wt, y, K, N, norm=None, xmode=0, res=None):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.eetq_b200_w8a16_gemv_fused(vp(xin), ldx, vp(wt[0]), vp(wt[1]), None, vp(norm), 1e-5, xmode, vp(res), N, vp(y), N, 1, N, K,
                                      _cabi.F16, pdl, st)
    _cabi.check(rc, "gemv")


def chain():
    for w in layers:
        gemv(x, H, w["qkv"], qkv, H, 3 * H, norm=w["ln"], xmode=1)
        gemv(attn, H, w["o"], x2, H, H, res=x)
        gemv(x2, H, w["gu"], gu, H, 2 * I, norm=w["ln"], xmode=1)
        gemv(gu, 2 * I, w["down"], x, I, H, xmode=2, res=x2)


side = torch.cuda.Stream()
with torch.cuda.stream(side):
    chain()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 4)
ts.sort()
ms = ts[len(ts) // 2]
nbytes = LAYERS * sum(K * N + 2 * N + 2 * K + 2 * N for K, N in [(H, 3 * H), (H, H), (H, 2 * I), (I, H)])
print(json.dumps({"knobs": {k: v for k, v in os.environ.items() if k.startswith("EETQ_B200") or k == "CHAIN_PDL"},
                  "ms_per_token_gemv": ms, "us_per_launch": ms * 1e3 / (4 * LAYERS), "gbs": nbytes / ms / 1e6}), flush=True)
