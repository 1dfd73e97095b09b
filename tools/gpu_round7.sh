#!/bin/bash
mkdir -p gpurun_out
run() { echo "### $*" >> gpurun_out/chain2.log; env "$@" timeout 300 python tools/gemv_chain.py >> gpurun_out/chain2.log 2>&1; }
: > gpurun_out/chain2.log
run A=base
run EETQ_B200_GEMV_PREFETCH=1
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_decode_gpu.py -m gpu -q --timeout 300 --timeout-method=thread -x -k "gemv or decode or fused" > gpurun_out/t_gemv.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemv.log
EETQ_B200_GEMV_PREFETCH=1 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_pre1.json 2> gpurun_out/bench_pre1.err
cat gpurun_out/chain2.log | grep -v "^$"; tail -n 3 gpurun_out/t_gemv.log; cat gpurun_out/bench_pre1.json | cut -c1-400
