#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_check_$N.log
EETQ_B200_ALLGATHER=p2p timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --warmup 4 > gpurun_out/bench_n${N}_p2p.json 2> gpurun_out/bench_n${N}_p2p.err; echo "rc=$?" >> gpurun_out/bench_n${N}_p2p.err
grep -E 'MGPU|rror|rc=' gpurun_out/mgpu_check_$N.log | head; cat gpurun_out/bench_n${N}_p2p.json | cut -c1-700; grep -E 'rror|p2p|rc=' gpurun_out/bench_n${N}_p2p.err | head -5
