#!/bin/bash
# run with: gpurun --gpus 8 -- bash tools/gpu_mgpu8.sh   (token check, Llama-2-7B and Llama-2-13B (BASELINE C5) sharded 8 ways)
N=${1:-8}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 64 --warmup 4 --model llama-2-13b > gpurun_out/bench_13b_n$N.json 2> gpurun_out/bench_13b_n$N.err; echo "rc=$?" >> gpurun_out/bench_13b_n$N.err
grep -E 'MGPU|rror|rc=' gpurun_out/mgpu_check_$N.log | head
cut -c1-400 gpurun_out/bench_n$N.json; grep -E 'rror|rc=' gpurun_out/bench_n$N.err | head -5
cut -c1-400 gpurun_out/bench_13b_n$N.json; grep -E 'rror|rc=' gpurun_out/bench_13b_n$N.err | head -5
