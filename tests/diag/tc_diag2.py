"""Which 256-k stage of which feature row goes wrong in the tcgen05 kernel?  Token t carries ones in stage (t % spt) only, so
y[t, n] = s_n * sum_{k in stage} q[k, n]; a wrong value is matched against the sums of all stages of that row."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402
from oracle import w8a16_oracle as o  # noqa: E402


def main():
    M, K, N = [int(v) for v in sys.argv[1:4]]
    dev = torch.device("cuda", 0)
    w = o.synth_weight(K, N, seed=1)
    q, s, _ = o.quantize(w)
    wq = o.b200_layout(q).to(dev)
    spt = K // 256
    x = torch.zeros(M, K, dtype=torch.float16)
    for t in range(M):
        st = t % spt
        x[t, st * 256:(st + 1) * 256] = 1
    wd = (q.half() * s.half()).float()                      # [K, N]
    S = wd.view(spt, 256, N).sum(1)                          # [spt, N] per-stage sums
    yr = S[[t % spt for t in range(M)]]                      # [M, N]
    sd = s.to(dev)
    xd = x.to(dev)
    for rep in range(3):
        y = w8_a16_gemm_bias(xd, wq, sd, None, flags=_cabi.FLAG_FORCE_TC).float().cpu()
        torch.cuda.synchronize()
        d = (y - yr).abs()
        bad = (d > 2e-2 * yr.abs().max()).nonzero()
        print(f"rep {rep}: bad elements {bad.shape[0]} of {M * N}", flush=True)
        seen = 0
        rows_bad = {}
        for t, n in bad.tolist():
            rows_bad.setdefault((t // 256, n), []).append(t)
        for (tt, n), ts in list(rows_bad.items())[:12]:
            stages = sorted({t % spt for t in ts})
            t = ts[0]
            st = t % spt
            cand = (S[:, n] - y[t, n]).abs()
            best = int(cand.argmin())
            print(f"   token-tile {tt} feature {n} (tile {n // 128}, lane {n % 128}): {len(ts)} bad tokens, stages {stages}; token {t} stage {st}: "
                  f"got {y[t, n]:.4f} want {yr[t, n]:.4f}; closest stage sum: stage {best} ({S[best, n]:.4f}); zero? {abs(y[t, n]) < 1e-6}")


if __name__ == "__main__":
    main()
