#!/bin/bash
# decode-row kernels on one B200: parity of the streaming kernels (int8 + int4), then the A/B micro-benchmark (twice: int4 rows-per-group knob)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_int4_gpu.py -m gpu -q --timeout 120 --timeout-method=thread > $O/t_mma2.log 2>&1; echo "pytest rc=$?" >> $O/t_mma2.log
tail -n 6 $O/t_mma2.log
timeout 300 python tests/perf/kbench_mma2.py > $O/kbench_mma2.log 2>&1; echo "rc=$?" >> $O/kbench_mma2.log
tail -n 2 $O/kbench_mma2.log
cp $O/kbench_mma2.json $O/kbench_mma2_r2.json
EETQ_B200_GEMV4_R=1 KBENCH_ONLY_INT4_SIMT=1 timeout 200 python tests/perf/kbench_mma2.py > $O/kbench_mma2_r1.log 2>&1; echo "rc=$?" >> $O/kbench_mma2_r1.log
cp $O/kbench_mma2.json $O/kbench_mma2_r1.json
grep simt $O/kbench_mma2_r1.log
