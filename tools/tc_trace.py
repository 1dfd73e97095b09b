"""Timeline of the tcgen05 kernel's pipeline roles (instrumented build, fp16): where does a CTA's time go?

Per shape: medians over CTAs, in SM cycles relative to the CTA's first instruction, of
  setup done | first weight TMA issued | dequant warp 0: A stage i written | MMA thread: stage i issued | MMA done |
  epilogue segment begin/end | CTA end;  plus the globaltimer span of the whole grid.
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi  # noqa: E402


def med(v):
    v = sorted(v)
    return v[len(v) // 2] if v else None


def run(M, K, N, reps=3):
    dev = torch.device("cuda", 0)
    L = _cabi.lib()
    w = torch.randint(-128, 128, (K, N), dtype=torch.int8, device=dev)
    sc = (torch.rand(N, device=dev) * 0.01).half()
    x = torch.randn(M, K, device=dev).half()
    y = torch.empty(M, N, device=dev, dtype=torch.float16)
    ws_bytes = int(L.eetq_b200_workspace_bytes(M, N, K))
    ws = torch.zeros(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    grid, slots = ctypes.c_int(0), ctypes.c_int(0)
    _cabi.check(L.eetq_b200_w8a16_gemm_trace_info(M, N, K, ctypes.byref(grid), ctypes.byref(slots)), "trace_info")
    G, S = grid.value, slots.value
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = None
    for _ in range(reps):
        tr = torch.zeros(G * S, dtype=torch.int64, device=dev)
        rc = L.eetq_b200_w8a16_gemm_trace(vp(x), K, vp(w), vp(sc), vp(y), N, M, N, K, vp(ws), ws.numel(), vp(tr), tr.numel() * 8, st)
        _cabi.check(rc, "gemm_trace")
        torch.cuda.synchronize()
        out = tr.view(G, S).cpu()
    t0 = out[:, 49]
    rel = lambda col: [int(out[g, col] - t0[g]) for g in range(G) if out[g, col] != 0]
    res = {
        "M": M, "K": K, "N": N, "grid": G, "units_per_cta": med([int(v) for v in out[:, 54]]),
        "grid_span_ns": int(out[:, 53].max() - out[:, 48].min()),
        "cta_start_skew_ns": int(out[:, 48].max() - out[:, 48].min()),
        "setup_done": med(rel(50)), "first_w_tma": med(rel(51)),
        "dequant_stage_done": [med(rel(i)) for i in range(12)],
        "mma_stage_issued": [med(rel(16 + i)) for i in range(12)],
        "mma_done": med(rel(30)),
        "epi_seg": [[med(rel(32 + 2 * i)), med(rel(33 + 2 * i))] for i in range(4)],
        "joint_fixup": [med(rel(40)), med(rel(41))],
        "cta_end": med(rel(52)), "cta_end_max": max(rel(52)),
    }
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    shapes = [(16, 4096, 4096), (64, 4096, 4096), (256, 4096, 4096), (1024, 4096, 4096), (64, 4096, 11008), (256, 11008, 4096)]
    for (M, K, N) in shapes:
        run(M, K, N)
