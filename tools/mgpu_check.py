"""Multi-GPU correctness check (run under torchrun): the column/head/vocab-sharded decoder must generate exactly the tokens of
the single-GPU decoder, for the fused LL exchange and for the NCCL all-gather baseline."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eetq_b200  # noqa: E402
from eetq_b200.decode import LlamaShape, LlamaSkeleton, W8A16LlamaDecoder  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    layers = int(os.environ.get("MGPU_LAYERS", "4"))
    shape = LlamaShape(hidden=4096, inter=11008, layers=layers, heads=32, vocab=32000, name=f"7b-{layers}layer")
    model = LlamaSkeleton(shape, device=dev, seed=1000)
    eetq_b200.eet_quantize(model)
    prompt = torch.randint(0, shape.vocab, (200,), generator=torch.Generator(device=dev).manual_seed(11), device=dev)
    n_new = 24
    ref = W8A16LlamaDecoder.from_model(model, max_ctx=256).generate(prompt, n_new)
    ok_all = True
    for exch in ("ll", "nccl"):
        dec = W8A16LlamaDecoder.from_model(model, max_ctx=256, rank=rank, world_size=world, exchange=exch)
        got = dec.generate(prompt, n_new)
        ok = got == ref
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"MGPU_CHECK world={world} exchange={dec.exchange} tokens_match_single_gpu={bool(flag.item())} "
                  f"launches_per_step={dec.launches_per_step}", flush=True)
            if not ok:
                print("  ref:", ref[:12], "\n  got:", got[:12], flush=True)
        ok_all &= bool(flag.item())
        del dec
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok_all else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
