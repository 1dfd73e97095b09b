#!/bin/bash
# round 2: after the warp re-convergence fix in the tcgen05 kernel -- racy shapes, full parity suite, tcgen05 kernel bench
mkdir -p gpurun_out
O=gpurun_out/dbg2.log
{
for sh in "256 4096 11008" "1024 4096 4096" "1024 11008 4096" "1024 4096 11008"; do
  timeout 100 python tools/tc_diag.py $sh
done
EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=64 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_DQW=16 timeout 100 python tools/tc_diag.py 64 4096 11008
EETQ_B200_TC_DQW=16 timeout 100 python tools/tc_diag.py 1024 4096 4096
} > $O 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc.json > gpurun_out/kb_tc.log 2>&1
EETQ_B200_TC_DQW=16 timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc_dqw16.json > gpurun_out/kb_tc_dqw16.log 2>&1
timeout 400 python tools/kbench.py --gemv-only --out gpurun_out/kb_gemv.json > gpurun_out/kb_gemv.log 2>&1
cut -c1-300 $O
tail -n 5 gpurun_out/t_all.log
grep gemm_tc gpurun_out/kb_tc.log | cut -c1-260
