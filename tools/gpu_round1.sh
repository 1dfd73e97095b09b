#!/bin/bash
# first GPU pass: parity tests, tcgen05 debug, kernel microbench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> gpurun_out/tc_debug.log
timeout 900 python -m pytest tests/test_quantize_gpu.py -m gpu -q --timeout 120 --timeout-method=thread > gpurun_out/t_quant.log 2>&1
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -k "gemv or reference or uniform or identity or extreme" > gpurun_out/t_gemv.log 2>&1
timeout 1200 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -k "not (gemv or reference or uniform or identity or extreme)" > gpurun_out/t_tc.log 2>&1
timeout 900 python tools/kbench.py > gpurun_out/kbench.log 2>&1
tail -5 gpurun_out/tc_debug.log gpurun_out/t_quant.log gpurun_out/t_gemv.log gpurun_out/t_tc.log
tail -40 gpurun_out/kbench.log
