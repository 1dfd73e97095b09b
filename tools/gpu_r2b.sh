#!/bin/bash
# round 2, call B: BT=256 diagnosis + v2 performance picture
mkdir -p gpurun_out
O=gpurun_out
{
for dq in 8 16; do
  EETQ_B200_TC_DQW=$dq timeout 100 python tools/tc_diag.py 1024 4096 4096
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_BT=128 timeout 100 python tools/tc_diag.py 1024 4096 4096
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_BT=256 timeout 100 python tools/tc_diag.py 256 1024 256
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_BT=256 timeout 100 python tools/tc_diag.py 512 4096 4096
  EETQ_B200_TC_DQW=$dq timeout 100 python tools/tc_diag.py 1024 4096 11008
done
} > $O/b_diag.log 2>&1
timeout 300 python tools/tc_trace.py > $O/b_trace_dqw8.jsonl 2> $O/b_trace_dqw8.err
EETQ_B200_TC_DQW=16 timeout 300 python tools/tc_trace.py > $O/b_trace_dqw16.jsonl 2> $O/b_trace_dqw16.err
timeout 400 python tools/kbench.py --tc-only --out $O/b_kb_v2_dqw8.json > $O/b_kb_v2_dqw8.log 2>&1
EETQ_B200_TC_DQW=16 timeout 400 python tools/kbench.py --tc-only --out $O/b_kb_v2_dqw16.json > $O/b_kb_v2_dqw16.log 2>&1
EETQ_B200_TC_L2PROMO=128 timeout 400 python tools/kbench.py --tc-only --quick --out $O/b_kb_v2_promo128.json > $O/b_kb_v2_promo128.log 2>&1
KBENCH_TC_PDL=1 timeout 400 python tools/kbench.py --tc-only --quick --out $O/b_kb_v2_pdl.json > $O/b_kb_v2_pdl.log 2>&1
# new MMA streaming kernel for M = 2..8 vs the SIMT one vs the reference kernel
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "gemv or bf16 or bias or extreme or identity" --timeout 200 --timeout-method=thread > $O/b_t_gemv.log 2>&1; echo "rc=$?" >> $O/b_t_gemv.log
timeout 500 python tools/kbench.py --gemv-only --out $O/b_kb_gemv_mma.json > $O/b_kb_gemv_mma.log 2>&1
cat $O/b_diag.log
grep gemm_tc $O/b_kb_v2_dqw8.log | cut -c1-230
tail -n 3 $O/b_t_gemv.log
