#!/bin/bash
# A/B of the decode-row kernels on one B200: parity of the v2 mma.sync kernel (int8 + int4), then the micro-benchmark
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_int4_gpu.py -m gpu -q -k "mma" --timeout 120 --timeout-method=thread > $O/t_mma2.log 2>&1; echo "pytest rc=$?" >> $O/t_mma2.log
tail -n 30 $O/t_mma2.log
timeout 400 python tools/kbench_mma2.py > $O/kbench_mma2.log 2>&1; echo "rc=$?" >> $O/kbench_mma2.log
cat $O/kbench_mma2.log
