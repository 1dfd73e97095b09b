#!/bin/bash
# run with: gpurun --gpus N -- bash tools/gpu_mgpu.sh N   (token equality vs single GPU, then the bench with both exchanges)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
EETQ_B200_EXCHANGE=nccl timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?" >> gpurun_out/bench_n${N}_nccl.err
grep -E 'MGPU|rror|rc=' gpurun_out/mgpu_check_$N.log | head; tail -n 5 gpurun_out/mgpu_check_$N.log | cut -c1-300
cut -c1-1500 gpurun_out/bench_n$N.json; grep -E 'rror|rc=' gpurun_out/bench_n$N.err | head -5
cut -c1-600 gpurun_out/bench_n${N}_nccl.json; grep -E 'rror|rc=' gpurun_out/bench_n${N}_nccl.err | head -5
