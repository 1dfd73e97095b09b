#!/bin/bash
# run with: gpurun --gpus N -- bash tools/gpu_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
EETQ_B200_ALLGATHER=p2p timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n${N}_p2p.json 2> gpurun_out/bench_n${N}_p2p.err
grep -E 'MGPU|rror|p2p' gpurun_out/mgpu_check_$N.log | head; cat gpurun_out/bench_n$N.json; grep -E 'rror|p2p|rc=' gpurun_out/bench_n$N.err | head -5; cat gpurun_out/bench_n${N}_p2p.json
