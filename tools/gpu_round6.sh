#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> gpurun_out/tc_debug.log
if grep -q "TC_DEBUG PASS" gpurun_out/tc_debug.log; then
  timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 200 --timeout-method=thread -x -k "tc or bf16 or identity or ragged or bias or modules or repeat or graph" > gpurun_out/t_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_tc.log
  timeout 600 python tools/kbench.py --tc-only > gpurun_out/kbench_tc.log 2>&1
fi
tail -n 6 gpurun_out/tc_debug.log; tail -n 4 gpurun_out/t_tc.log; tail -n 16 gpurun_out/kbench_tc.log
