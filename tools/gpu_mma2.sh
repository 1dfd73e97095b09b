#!/bin/bash
# decode-row kernels on one B200: parity of the mma.sync kernel (int8 + int4) and int4 suite, micro-benchmark, ncu of five kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_int4_gpu.py -m gpu -q --timeout 120 --timeout-method=thread > $O/t_mma2.log 2>&1; echo "pytest rc=$?" >> $O/t_mma2.log
tail -n 12 $O/t_mma2.log
timeout 300 python tools/kbench_mma2.py > $O/kbench_mma2.log 2>&1; echo "rc=$?" >> $O/kbench_mma2.log
tail -n 3 $O/kbench_mma2.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/prof_int4 -f python tools/prof_int4.py > $O/ncu_int4.log 2>&1; echo "rc=$?" >> $O/ncu_int4.log
tail -n 4 $O/ncu_int4.log
