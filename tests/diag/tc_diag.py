"""Per-tile error map of the tcgen05 kernel for one shape (diagnostics): which (token tile, feature tile) blocks are wrong,
by how much, and whether repeated launches agree."""
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
    bt = int(os.environ.get("EETQ_B200_TC_BT", "0")) or (16 if M <= 16 else 32 if M <= 32 else 64 if M <= 64 else 128 if M <= 128 else 256)
    dev = torch.device("cuda", 0)
    w = o.synth_weight(K, N, seed=1)
    q, s, _ = o.quantize(w)
    wq = o.b200_layout(q).to(dev)
    x = o.synth_act(M, K)
    yr = o.gemm(x, q, s).float()
    scale = yr.abs().max()
    outs = []
    for rep in range(3):
        y = w8_a16_gemm_bias(x.to(dev), wq, s.to(dev), None, flags=_cabi.FLAG_FORCE_TC)
        torch.cuda.synchronize()
        outs.append(y.cpu().float())
    same = all(torch.equal(outs[0], t) for t in outs[1:])
    for rep, y in enumerate(outs):
        d = (y - yr).abs() / scale
        bad = []
        for tt in range((M + bt - 1) // bt):
            for nt in range((N + 127) // 128):
                e = d[tt * bt:(tt + 1) * bt, nt * 128:(nt + 1) * 128].max().item()
                if e > 1e-3:
                    bad.append((tt, nt, round(e, 4)))
        print(f"M={M} K={K} N={N} bt={bt} dqw={os.environ.get('EETQ_B200_TC_DQW', '8')} rep={rep}: err={d.max().item():.3e} "
              f"bad_tiles={len(bad)} first={bad[:8]} repeatable={same}", flush=True)
        if bad and rep == 0:
            tt, nt, _ = bad[0]
            blk = d[tt * bt:(tt + 1) * bt, nt * 128:(nt + 1) * 128]
            rows = (blk.max(dim=1).values > 1e-3).nonzero().flatten().tolist()
            cols = (blk.max(dim=0).values > 1e-3).nonzero().flatten().tolist()
            print(f"   tile ({tt},{nt}): bad token rows {len(rows)} [{rows[:6]}..{rows[-3:]}], bad features {len(cols)} [{cols[:6]}..{cols[-3:]}]", flush=True)


if __name__ == "__main__":
    main()
