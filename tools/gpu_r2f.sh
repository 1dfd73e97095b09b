#!/bin/bash
mkdir -p gpurun_out
{
timeout 100 python tools/tc_diag2.py 1024 4096 4096
EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag2.py 1024 4096 4096
} > gpurun_out/dbg3.log 2>&1
cut -c1-400 gpurun_out/dbg3.log
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_gemm_gpu.py -q -x --timeout 300 -k "not tc_gemm_baseline and not long_context" > gpurun_out/t_part.log 2>&1; tail -n 5 gpurun_out/t_part.log
timeout 600 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
