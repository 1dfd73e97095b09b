"""CPU, world_size-2 gloo test of the N>1 host logic: column (output-feature) sharding of the b200 layout is a
contiguous byte range per rank, quantise-then-shard == shard-then-quantise bit for bit, and
the all-gather of per-rank outputs in rank order reconstructs the unsharded result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eetq_b200.decode import _ShardedLinear
    from eetq_b200.modules.qlinear import W8A16Linear
    from oracle import w8a16_oracle as o

    K, N1, N2 = 256, 128, 384
    lins, qs, ss = [], [], []
    for i, N in enumerate((N1, N2)):                       # two linears fused row-wise, like q|k|v or gate|up
        w = o.synth_weight(K, N, seed=20 + i)
        q, s, _ = o.quantize(w)
        lin = W8A16Linear(K, N, bias=False, dev="cpu")
        lin.qweight = o.b200_layout(q)
        lin.weight_scales = s
        lins.append(lin); qs.append(q); ss.append(s)
    sh = _ShardedLinear(lins, rank, world)
    q_full, s_full = torch.cat(qs, 1), torch.cat(ss, 0)
    n0, n1 = sh.n_begin, sh.n_begin + sh.n_local
    # 1. this rank's bytes are exactly the b200 layout of its column slice (contiguous rows of the fused matrix)
    ok = torch.equal(sh.w, o.b200_layout(q_full[:, n0:n1].contiguous()))
    # 2. quantise-then-shard == shard-then-quantise
    w_full = torch.cat([o.synth_weight(K, N, seed=20 + i) for i, N in enumerate((N1, N2))], 1)
    q_loc, s_loc, _ = o.quantize(w_full[:, n0:n1].contiguous())
    ok = ok and torch.equal(q_loc, q_full[:, n0:n1]) and torch.equal(s_loc, s_full[n0:n1]) and torch.equal(sh.scales, s_loc)
    # 3. all-gather of the per-rank outputs (rank order) == unsharded output
    x = o.synth_act(2, K, seed=3)
    y_loc = o.gemm(x, o.b200_layout_inv(sh.w), sh.scales)
    parts = [torch.empty_like(y_loc) for _ in range(world)]
    dist.all_gather(parts, y_loc)
    y = torch.cat(parts, 1)
    ok = ok and torch.equal(y, o.gemm(x, q_full, s_full))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        q_out.put(int(flag.item()))
    dist.destroy_process_group()


def test_column_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1
