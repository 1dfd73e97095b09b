#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 python tools/kbench.py --tc-only > gpurun_out/kbench_tc.log 2>&1
BENCH_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fused -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
tail -n 4 gpurun_out/t_all.log
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
cat gpurun_out/kbench_tc.log | tail -n 40
