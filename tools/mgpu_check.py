"""torchrun check of the column-sharded decode path (N > 1): every rank must produce the same greedy tokens as a
single-GPU decoder built from the same seeded model.  Run: torchrun --nproc-per-node N tools/mgpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eetq_b200  # noqa: E402
from eetq_b200.decode import LlamaShape, LlamaSkeleton, W8A16LlamaDecoder  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = LlamaShape(hidden=1024, inter=2816, layers=3, heads=8, vocab=2048, name="tiny-mgpu")   # 2816 = 44*64; shards stay %64
model = LlamaSkeleton(shape, device=dev, seed=7, std=0.05)
eetq_b200.eet_quantize(model)
prompt = torch.randint(0, shape.vocab, (24,), generator=torch.Generator(device=dev).manual_seed(1), device=dev)
ref = W8A16LlamaDecoder.from_model(model, max_ctx=128).generate(prompt, 16)
allok = 1
for mode in ("nccl", "p2p"):
    dec = W8A16LlamaDecoder.from_model(model, max_ctx=128, rank=rank, world_size=world, allgather=mode)
    out = dec.generate(prompt, 16)
    ok = torch.tensor([1 if out == ref else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    allok &= int(ok.item())
    if rank == 0:
        print("MGPU_CHECK", "PASS" if int(ok.item()) == 1 else "FAIL", "world", world, "requested", mode, "used", dec.allgather, out[:8],
              ref[:8], flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if allok else 1)   # no NCCL teardown: destroying the group with captured graphs alive can hang
