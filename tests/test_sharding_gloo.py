"""CPU, world_size-2 gloo test of the N>1 host logic of the decoder: the per-rank shards W8A16LlamaDecoder builds
(heads for q|k|v, contiguous feature blocks for o / down, interleaved (gate, up) rows for the MLP, vocabulary rows for the
lm_head) are contiguous byte ranges of the b200 layout, quantise-then-shard == shard-then-quantise bit for bit
(SURVEY.md section 8e), and gathering every rank's outputs in rank order reconstructs the unsharded layer."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _quantised_tiny(o, shape):
    """LlamaSkeleton on the CPU whose linears are W8A16Linear modules filled by the ORACLE quantiser (no GPU here)."""
    from eetq_b200.decode import LlamaSkeleton
    from eetq_b200.modules.qlinear import W8A16Linear
    from eetq_b200.utils.base import find_layers, set_op_by_name

    m = LlamaSkeleton(shape, device="cpu", dtype=torch.float16, seed=5, std=0.05)
    raw = {}
    for name, lin in find_layers(m).items():
        w_kn = lin.weight.detach().t().contiguous()
        q, s, _ = o.quantize(w_kn)
        ql = W8A16Linear(lin.in_features, lin.out_features, bias=False, dev="cpu")
        ql.qweight = o.b200_layout(q)
        ql.weight_scales = s
        set_op_by_name(m, name, ql)
        raw[name] = (w_kn, q, s)
    return m, raw


def _worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eetq_b200.decode import LlamaShape, W8A16LlamaDecoder
    from oracle import w8a16_oracle as o

    shape = LlamaShape(hidden=256, inter=512, layers=1, heads=2, vocab=128, name="tiny-gloo")
    model, raw = _quantised_tiny(o, shape)
    dec = W8A16LlamaDecoder(model, shape, max_ctx=16, rank=rank, world_size=world, exchange="nccl")
    plan, w = dec.plan, dec.layers[0]
    D = shape.head_dim
    h0, h1 = plan["heads"]
    n0, n1 = plan["hidden"]
    i0, i1 = plan["inter"]
    pre = "model.layers.0."
    ok = True

    def rows_of(shard):   # [N_local, K] int8 rows (unbiased) of a shard
        return o.b200_layout_inv(shard.w).t().contiguous()

    # 1. q|k|v shard = the rows of this rank's heads, q then k then v; quantise-then-shard == shard-then-quantise
    want_q = torch.cat([raw[pre + f"self_attn.{n}_proj"][1][:, h0 * D:h1 * D] for n in "qkv"], 1)
    ok &= torch.equal(o.b200_layout_inv(w["qkv"].w), want_q)
    w_kn = raw[pre + "self_attn.k_proj"][0][:, h0 * D:h1 * D].contiguous()
    q_loc, s_loc, _ = o.quantize(w_kn)
    ok &= torch.equal(q_loc, raw[pre + "self_attn.k_proj"][1][:, h0 * D:h1 * D])
    ok &= torch.equal(s_loc, w["qkv"].scales[(h1 - h0) * D:2 * (h1 - h0) * D])
    # 2. o / down shards = contiguous feature blocks
    ok &= torch.equal(o.b200_layout_inv(w["o"].w), raw[pre + "self_attn.o_proj"][1][:, n0:n1])
    ok &= torch.equal(o.b200_layout_inv(w["down"].w), raw[pre + "mlp.down_proj"][1][:, n0:n1])
    # 3. gate|up shard = interleaved (g_i, u_i) rows of this rank's block
    gu = rows_of(w["gu"])
    ok &= torch.equal(gu[0::2], raw[pre + "mlp.gate_proj"][1][:, i0:i1].t())
    ok &= torch.equal(gu[1::2], raw[pre + "mlp.up_proj"][1][:, i0:i1].t())
    # 4. one MLP through the shards + all-gathers in rank order == the unsharded MLP (oracle arithmetic)
    x = o.synth_act(1, shape.hidden, seed=3)
    gq, gs = o.b200_layout_inv(w["gu"].w), w["gu"].scales
    y = o.gemm(x, gq, gs)[0]
    act_loc = (torch.nn.functional.silu(y[0::2].float()).half() * y[1::2])
    parts = [torch.empty_like(act_loc) for _ in range(world)]
    dist.all_gather(parts, act_loc)
    act = torch.cat(parts)
    g_full = o.gemm(x, raw[pre + "mlp.gate_proj"][1], raw[pre + "mlp.gate_proj"][2])[0]
    u_full = o.gemm(x, raw[pre + "mlp.up_proj"][1], raw[pre + "mlp.up_proj"][2])[0]
    ok &= torch.equal(act, torch.nn.functional.silu(g_full.float()).half() * u_full)
    down_loc = o.gemm(act[None], o.b200_layout_inv(w["down"].w), w["down"].scales)[0]
    parts = [torch.empty_like(down_loc) for _ in range(world)]
    dist.all_gather(parts, down_loc)
    ok &= torch.equal(torch.cat(parts), o.gemm(act[None], raw[pre + "mlp.down_proj"][1], raw[pre + "mlp.down_proj"][2])[0])
    # 5. vocabulary-sharded lm_head: exchanging (value, index) candidates gives the unsharded arg-max (first maximum)
    v0, v1 = plan["vocab"]
    ok &= torch.equal(dec.lm_head_w, model.lm_head.weight[v0:v1])
    h = torch.randn(shape.hidden, generator=torch.Generator().manual_seed(1)).half()
    logits_loc = (h.float() @ dec.lm_head_w.float().t()).half()
    cand = torch.tensor([float(logits_loc.max()), float(v0 + int(logits_loc.argmax()))])
    cands = [torch.empty_like(cand) for _ in range(world)]
    dist.all_gather(cands, cand)
    best = max(cands, key=lambda c: (float(c[0]), -float(c[1])))
    full = (h.float() @ model.lm_head.weight.float().t()).half()
    ok &= int(best[1]) == int(full.argmax())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        q_out.put(int(flag.item()))
    dist.destroy_process_group()


def test_decoder_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1
