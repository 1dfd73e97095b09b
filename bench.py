#!/usr/bin/env python
"""bench.py -- the contract benchmark: decode tokens/sec of Llama-2-7B with w8a16 linears on N B200s.

Metric (BASELINE.json): "decode tokens/sec Llama-2-7B w8a16 @1/2/4/8 B200; GEMV HBM GB/s vs roofline".
Workload (config.workload): random-init Llama-2-7B (hidden 4096, inter 11008, 32 layers, 32 heads, vocab 32000), every
nn.Linear except lm_head quantised by eet_quantize, batch 1, a 1024-token synthetic prompt prefetched into the KV
cache by a real prefill, then greedy decode.  One "step" = one decoded token (224 quantised linears as 128 fused
streaming GEMVs + 32 attention launches + one final-norm/lm_head/arg-max launch).  Each step streams ~6.5 GB of int8 weights, far more than the 126 MB L2,
so no explicit L2 flush is needed between steps (config.l2 says so).

  value      device-resident decode: K CUDA-graph replays back to back, token fed back on the device
  e2e        the same K steps through W8A16LlamaDecoder.step_host(): token id copied from pinned host memory every
             step, next token id copied back and synchronised every step
  roofline   the streaming GEMV (w8a16_gemv_kernel), timed live with CUDA events over all 128 GEMV launches of one token
  cpu_baseline  EETQ-style dequantise -> torch.matmul on the host cores (oracle.cpu_dequant_matmul), bounded sample

N > 1 (torchrun): every linear column-sharded over the ranks, attention sharded by head, lm_head by vocabulary; the activation
vectors are exchanged by the producing kernels' epilogues over NVLink (LL words, see eetq_b200/decode.py) -- strong scaling.
--impl reference: the reference's CPU dequant->matmul path on the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-2-7b", choices=["llama-2-7b", "llama-2-13b"])
    ap.add_argument("--prompt", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=0, help="debug: override the layer count (marks the run as reduced)")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def shape_of(args):
    from eetq_b200.decode import LLAMA2_7B, LLAMA2_13B, LlamaShape
    import dataclasses

    s = LLAMA2_7B if args.model == "llama-2-7b" else LLAMA2_13B
    if args.layers:
        s = dataclasses.replace(s, layers=args.layers, name=s.name + f"-{args.layers}layers-DEBUG")
    return s


def linear_shapes(s):
    """(K, N, count per layer) of the quantised nn.Linear modules eet_quantize converts."""
    return [(s.hidden, s.hidden, 4), (s.hidden, s.inter, 2), (s.inter, s.hidden, 1)]


# ----------------------------------------------------------------------------------------------------------------
def cpu_layer_time(s, reps: int, threads: int):
    """Median seconds for the 7 quantised linears of ONE decoder layer at M=1 on the host cores, EETQ-style:
    dequantise (q.half() * scales) then torch.matmul, every call (oracle.cpu_dequant_matmul)."""
    from oracle import w8a16_oracle as o

    torch.set_num_threads(threads)
    mats = []
    for (K, N, cnt) in linear_shapes(s):
        q = torch.randint(-128, 128, (K, N), dtype=torch.int8)
        sc = (torch.rand(N) * 0.01).half()
        x = torch.randn(1, K).half()
        mats.append((x, q, sc, cnt))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for x, q, sc, cnt in mats:
            for _ in range(cnt):
                o.cpu_dequant_matmul(x, q, sc)
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2]


def best_cpu_threads(s):
    """The CPU arm should be as fast as the host allows: torch's intra-op pool does not scale monotonically on this
    workload (M=1 matmul + element-wise dequant), so try a few pool sizes once and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[-1], float("inf")
    for c in cands:
        t = cpu_layer_time(s, 1, c)
        if t < best_t:
            best, best_t = c, t
    return best


def cpu_lm_head_time(s, threads: int):
    """Seconds for the fp16 lm_head matmul (weights pre-existing fp16, computed in fp32 like the rest of the CPU arm)."""
    torch.set_num_threads(threads)
    w = torch.randn(s.vocab, s.hidden)
    x = torch.randn(1, s.hidden)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        torch.matmul(x, w.t())
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_tokens_per_s(s, t_layer, t_lm_head):
    return 1.0 / (s.layers * t_layer + t_lm_head)


def bench_config(args, s, world, exchange):
    """`config` of the JSON line -- identical for our arm and the reference arm (same metric, same workload)."""
    par = "single-gpu" if world == 1 else (f"column-sharded linears + head-sharded attention + vocab-sharded lm_head x{world}; activations exchanged "
                                           + ("by LL words pushed over NVLink from the producing kernels' epilogues (fused all-gather)"
                                              if exchange == "ll" else "by ncclAllGather"))
    return {"workload": f"{s.name} eet_quantize w8a16, batch 1, prompt {args.prompt} (real prefill), greedy decode",
            "model": s.name, "global_batch": 1, "seq_len": args.prompt, "ctx_at_end": args.prompt + args.warmup + args.steps,
            "parallelism": par, "pdl": not args.no_pdl, "cuda_graph": True,
            "l2": "each step streams the rank's int8 weights (>= 0.8 GB per GPU, >> 126 MB L2): inputs larger than L2, no flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    s = shape_of(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = best_cpu_threads(s)
    # one step = a bounded sample of one token: the 7 quantised linears of ONE decoder layer (1/layers of a token), so
    # that K steps + W warm-up steps end within minutes on the host cores; ms_per_step is the REAL time of that step
    steps, warm = args.steps, args.warmup
    if warm:
        cpu_layer_time(s, warm, threads)
    t0 = time.perf_counter()
    t_layer = cpu_layer_time(s, steps, threads)
    wall = time.perf_counter() - t0
    t_lm = cpu_lm_head_time(s, threads)
    tps = cpu_tokens_per_s(s, t_layer, t_lm)
    exchange = os.environ.get("EETQ_B200_EXCHANGE", "ll") if world > 1 else "none"
    sample = (f"each step = the 7 quantised linears of 1 of {s.layers} decoder layers (EETQ-style dequant q.half()*s then torch.matmul, every "
              f"call; thread pool = fastest of 8/16/32/64/all); median step {t_layer * 1e3:.1f} ms; tokens/s = 1 / ({s.layers} x step + "
              f"lm_head {t_lm * 1e3:.1f} ms); attention excluded (negligible on the CPU next to the dequantisation)")
    line = {
        "impl": "reference", "metric": "decode_tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": wall * 1e3 / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": bench_config(args, s, world, exchange),
        "step_is": f"1/{s.layers} of a token (one decoder layer's linears); value is scaled to whole tokens",
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference has no CPU forward and its GPU kernels refuse sm>=90; this is the north-star CPU path "
                "(EETQ-style dequantise -> torch.matmul, examples/layers/test_w8a16_gemm.py:44-47) via oracle/w8a16_oracle.py",
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from eetq_b200 import _cabi, eet_quantize
    from eetq_b200.decode import LlamaSkeleton, W8A16LlamaDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    s = shape_of(args)
    ctx_max = args.prompt + args.steps + args.warmup + 64

    t0 = time.perf_counter()
    model = LlamaSkeleton(s, device=dev, dtype=torch.float16, seed=1000)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    eet_quantize(model)
    torch.cuda.synchronize()
    t_quant = time.perf_counter() - t0
    dec = W8A16LlamaDecoder.from_model(model, max_ctx=ctx_max, pdl=not args.no_pdl, rank=rank, world_size=world)

    # synthetic prompt, real prefill through the batched tcgen05 kernels (cold = first call: lazy module loads, tensor-map
    # encodes, workspace allocation; warm = the same prompt again)
    g = torch.Generator(device=dev).manual_seed(11)
    prompt = torch.randint(0, s.vocab, (args.prompt,), generator=g, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    e0.record()
    dec.prefill(prompt)
    e1.record()
    torch.cuda.synchronize()
    prefill_cold_ms = e0.elapsed_time(e1)
    barrier()
    e0.record()
    first_token = dec.prefill(prompt)
    e1.record()
    torch.cuda.synchronize()
    prefill_ms = e0.elapsed_time(e1)

    # N > 1: the sharded decoder must generate exactly the single-GPU tokens (rank 0 also runs an unsharded decoder)
    tokens_match = None
    n_check = min(16, args.steps)
    if world > 1:
        ref_tokens = None
        if rank == 0:
            single = W8A16LlamaDecoder.from_model(model, max_ctx=ctx_max, pdl=not args.no_pdl, rank=0, world_size=1)
            ref_tokens = single.generate(prompt, n_check + 1)
            del single
        got = [int(first_token.item())]
        dec.capture()
        for _ in range(n_check):
            dec.step()
            got.append(int(dec.token.item()))
        if rank == 0:
            tokens_match = bool(got == ref_tokens)
        barrier()
        dec.prefill(prompt)   # rewind to the end of the prompt
    del model
    torch.cuda.empty_cache()
    if dec.graph is None:
        dec.capture()

    # ------------------------------------------------------------------ value: device-resident decode
    for _ in range(args.warmup):
        dec.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    profile = os.environ.get("BENCH_PROFILE_RANGE") == "1"   # ncu --profile-from-start off captures only the timed steps
    if profile:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        dec.step()
    e1.record()
    barrier()
    if profile:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tps = args.steps / (ms / 1e3)

    # ------------------------------------------------------------------ e2e: host token in / host token out every step
    tok_h = torch.zeros(1, dtype=torch.int64).pin_memory()
    out_h = torch.zeros(1, dtype=torch.int64).pin_memory()
    dec.set_context(args.prompt, 1)
    tok_h[0] = 1
    for _ in range(min(3, args.warmup)):
        dec.step_host(tok_h, out_h)
        tok_h.copy_(out_h)
    dec.set_context(args.prompt, int(tok_h[0]))
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        dec.step_host(tok_h, out_h)
        tok_h.copy_(out_h)
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_tps = args.steps / e2e_s

    # ------------------------------------------------------------------ roofline of the dominant kernel (streaming GEMV)
    # all GEMV launches of one token on THIS rank's shards (4 fused GEMVs x layers, each on its own weights => working set
    # >> L2), captured as their own graph over plain scratch vectors (no exchange) and timed with CUDA events
    H, I = s.hidden, s.inter
    sx, sx2 = torch.randn(H, device=dev).half(), torch.randn(H, device=dev).half()
    sattn, sact = torch.randn(H, device=dev).half(), torch.randn(I, device=dev).half()
    sqkv = torch.zeros(3 * dec.Hl, dtype=torch.float16, device=dev)
    n0, i0 = dec.plan["hidden"][0], dec.plan["inter"][0]

    def gemv_only():
        for w in dec.layers:
            sl = slice(n0, n0 + w["o"].N)
            dec._gemv(sx, H, w["qkv"], sqkv, norm_w=w["ln1"], xmode=1)
            dec._gemv(sattn, H, w["o"], sx2[sl], residual=sx[sl])
            dec._gemv(sx2, H, w["gu"], sact[i0:i0 + dec.Il], norm_w=w["ln2"], xmode=1, epi=1)
            dec._gemv(sact, I, w["down"], sx[sl], residual=sx2[sl])

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gemv_only()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg):
        gemv_only()
    for _ in range(3):
        gg.replay()
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        gg.replay()
    e1.record()
    torch.cuda.synchronize()
    n_launch = 4 * s.layers
    us_per_launch = e0.elapsed_time(e1) * 1e3 / (reps * n_launch)
    w0_ = dec.layers[0]
    # algorithmic bytes: K*N weights + 2N scales + 2K activations + 2 x outputs
    per_layer = [(w0_["qkv"].K, w0_["qkv"].N, w0_["qkv"].N), (w0_["o"].K, w0_["o"].N, w0_["o"].N), (w0_["gu"].K, w0_["gu"].N, w0_["gu"].N // 2),
                 (w0_["down"].K, w0_["down"].N, w0_["down"].N)]
    bytes_per_launch = sum(K * N + 2 * N + 2 * K + 2 * outs for K, N, outs in per_layer) / 4.0
    achieved = bytes_per_launch / us_per_launch / 1e3  # GB/s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic per launch of this kernel from the committed `ncu --set full` capture (profiles/gemv_traffic.json, written by
    # tools/gemv_traffic.py: dram__bytes_read.sum + dram__bytes_write.sum averaged over the four decode GEMV shapes), or null
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemv_traffic.json")))
        ent = tj.get(f"{s.name}/world{world}")
        if ent:
            traffic, traffic_src = float(ent["bytes_per_launch"]), f"profiles/gemv_traffic.json ({ent.get('source', 'ncu --set full')}, commit {ent.get('commit', '?')})"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": "w8a16_gemv_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "us_per_launch": us_per_launch, "bytes_per_launch": bytes_per_launch, "rank": rank,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "how": f"CUDA events around {reps} replays of a graph holding the {n_launch} GEMV launches of one token on this rank's shards "
                   f"(each on its own layer's weights, {dec.weight_bytes_per_token() / 1e9:.2f} GB working set)"}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1, bounded sample)
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        threads = best_cpu_threads(s)
        t_layer = cpu_layer_time(s, 5, threads)
        t_lm = cpu_lm_head_time(s, threads)
        cpu = {"value": cpu_tokens_per_s(s, t_layer, t_lm), "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"7 quantised linears of 1 of {s.layers} decoder layers (EETQ-style dequant q.half()*s -> torch.matmul "
                         f"per call, oracle.cpu_dequant_matmul), median of 5; tokens/s = 1/({s.layers} x {t_layer * 1e3:.0f} ms + lm_head "
                         f"{t_lm * 1e3:.1f} ms); attention excluded"}

    if rank == 0:
        wbytes = dec.weight_bytes_per_token()
        line = {
            "metric": "decode_tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": bench_config(args, s, world, dec.exchange),
            "clocks": clocks,
            "e2e": {"value": e2e_tps, "unit": "tokens/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                    "api": "W8A16LlamaDecoder.step_host(pinned token) -> pinned next token, synchronised every step"},
            "gpu_launches": int(dec.launches_per_step) * args.steps,
            "launches_per_step": int(dec.launches_per_step),
            "roofline": roof,
            "cpu_baseline": cpu,
            "weights_only_roofline_tokens_per_s": float(roof["peak"]) * 1e9 / wbytes,
            "weight_bytes_per_token_per_gpu": wbytes,
            "tokens_match_single_gpu": tokens_match,
            "prefill_ms": prefill_ms, "prefill_cold_ms": prefill_cold_ms, "build_s": t_build, "quantize_s": t_quant,
        }
        emit(line)
    if world > 1:
        # tearing down NCCL while CUDA graphs that captured collectives are alive can hang: flush and leave
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so file
    descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
