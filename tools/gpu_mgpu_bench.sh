#!/bin/bash
# run with: gpurun --gpus N -- bash tools/gpu_mgpu_bench.sh N    (contract bench only, fused exchange)
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 128 --warmup 8 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
cut -c1-330 gpurun_out/bench_n$N.json; grep -E 'rror|rc=' gpurun_out/bench_n$N.err | head -5
