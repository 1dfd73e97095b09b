#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 32 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
cut -c1-420 gpurun_out/bench_n$N.json; grep -E 'rror|rc=' gpurun_out/bench_n$N.err | head -5
