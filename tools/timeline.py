"""In-situ timeline of one decode step (development tool).  Needs a library built with EETQ_B200_BUILD_TRACE=1.

Replays the decode graph, collects the per-CTA {kernel, block, event, globaltimer} records and prints, for every kernel launch of the
step, when its first CTA started, when its CTAs passed the dependency wait, finished their main loop and exited -- all relative to the
start of the step.  Shows where a token's 1.7 ms actually go (launch gaps, dependency stalls, tails)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eetq_b200 import _cabi, eet_quantize  # noqa: E402
from eetq_b200.decode import LLAMA2_7B, LlamaSkeleton, W8A16LlamaDecoder  # noqa: E402
import dataclasses  # noqa: E402

NAMES = {1: "gemv", 2: "attn", 3: "lm_head", 4: "embed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--ctx", type=int, default=1040)
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    s = dataclasses.replace(LLAMA2_7B, layers=args.layers)
    model = LlamaSkeleton(s, device=dev, seed=1000)
    eet_quantize(model)
    dec = W8A16LlamaDecoder.from_model(model, max_ctx=args.ctx + 64)
    dec.kcache.normal_()
    dec.vcache.normal_()
    dec.set_context(args.ctx, 1)
    dec.capture()
    for _ in range(3):
        dec.step()
    torch.cuda.synchronize()
    cap = 400000
    buf = torch.zeros(2 + 2 * cap, dtype=torch.int64, device=dev)
    on = _cabi.lib().eetq_b200_set_timeline(buf.data_ptr(), cap)
    assert on == 1, "library was not built with EETQ_B200_BUILD_TRACE=1"
    dec.step()
    torch.cuda.synchronize()
    _cabi.lib().eetq_b200_set_timeline(None, 0)
    h = buf.cpu()
    n = int(h[0])
    rec = h[2:2 + 2 * n].view(n, 2)
    key, t = rec[:, 0], rec[:, 1]
    tag = (key >> 56) & 0xff
    blk = (key >> 16) & 0xffffffffff
    ev = key & 0xffff
    t0 = int(t.min())
    # group records into launches: sort CTA start events by time; a new launch begins when the tag changes or block 0 reappears
    launches = []
    order = torch.argsort(t)
    cur = None
    # per (tag, launch) we need all events; launches of the same tag do not overlap in their START events order-wise
    starts = [(int(t[i]), int(tag[i]), int(blk[i])) for i in order.tolist() if int(ev[i]) == 0]
    seen = set()
    bounds = []  # (tag, first start time)
    for ts, tg, b in starts:
        if cur is None or tg != cur or (tg, b) in seen:
            cur = tg
            seen = set()
            bounds.append([tg, ts])
        seen.add((tg, b))
    # assign every record to the latest launch of its tag that started at or before it
    import bisect
    by_tag = {}
    for i, (tg, ts) in enumerate(bounds):
        by_tag.setdefault(tg, []).append((ts, i))
    ev_t = {}
    for i in range(n):
        tg, ti, e = int(tag[i]), int(t[i]), int(ev[i])
        lst = by_tag[tg]
        j = bisect.bisect_right([x[0] for x in lst], ti) - 1
        li = lst[max(j, 0)][1]
        ev_t.setdefault(li, {}).setdefault(e, []).append(ti - t0)
    rows = []
    for li, (tg, ts) in enumerate(bounds):
        d = ev_t.get(li, {})
        row = {"i": li, "kernel": NAMES.get(tg, str(tg)), "ctas": len(d.get(0, []))}
        for e, v in sorted(d.items()):
            v = sorted(v)
            row[f"e{e}"] = [round(v[0] / 1e3, 2), round(v[len(v) // 2] / 1e3, 2), round(v[-1] / 1e3, 2)]  # us: first / median / last
        rows.append(row)
    prev_end = 0.0
    print("times in us from the start of the step; per event: first / median / last CTA")
    for r in rows:
        last = max(v[2] for k, v in r.items() if k.startswith("e"))
        print(f"{r['i']:3d} {r['kernel']:8s} ctas={r['ctas']:4d} " + " ".join(f"{k}={v[0]:.1f}/{v[1]:.1f}/{v[2]:.1f}" for k, v in r.items() if k.startswith("e"))
              + f"  | ends {last:.1f} (+{last - prev_end:.1f})")
        prev_end = last
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"))


if __name__ == "__main__":
    main()
