#!/bin/bash
# round 2 debug call: tcgen05 failing shapes (tile error map), decode tests after the attention/rope fixes, bench, ncu of attention
mkdir -p gpurun_out
O=gpurun_out/dbg1.log
{
for sh in "64 4096 11008" "256 4096 11008" "1024 4096 4096" "64 4096 4096"; do
  timeout 100 python tools/tc_diag.py $sh
  EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py $sh
done
EETQ_B200_TC_BT=128 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=64 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=16 timeout 100 python tools/tc_diag.py 64 4096 11008
EETQ_B200_TC_DQW=16 timeout 100 python tools/tc_diag.py 64 4096 11008
} > $O 2>&1
timeout 900 python -m pytest tests/test_decode_gpu.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_decode.log 2>&1; echo "rc=$?" >> gpurun_out/t_decode.log
timeout 600 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
BENCH_PROFILE_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_decode -c 2 \
    -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --layers 4 > gpurun_out/ncu_attn.log 2>&1
BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
cut -c1-400 $O
tail -n 15 gpurun_out/t_decode.log
cut -c1-1200 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
