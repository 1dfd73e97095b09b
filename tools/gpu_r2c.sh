#!/bin/bash
# round 2, call C: tcgen05 kernel with uniform-datapath issue (elect.sync), joint fix-up, vector slots
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/tc_debug.py > $O/c_tc_debug_dqw8.log 2>&1; echo "rc=$?" >> $O/c_tc_debug_dqw8.log
EETQ_B200_TC_DQW=16 timeout 200 python tools/tc_debug.py > $O/c_tc_debug_dqw16.log 2>&1; echo "rc=$?" >> $O/c_tc_debug_dqw16.log
{
for dq in 8 16; do
  EETQ_B200_TC_DQW=$dq timeout 100 python tools/tc_diag.py 1024 4096 4096
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py 1024 4096 4096
  EETQ_B200_TC_DQW=$dq timeout 100 python tools/tc_diag.py 1024 4096 11008
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py 1024 4096 11008
  EETQ_B200_TC_DQW=$dq EETQ_B200_TC_BT=256 timeout 100 python tools/tc_diag.py 512 4096 4096
done
} > $O/c_diag.log 2>&1
timeout 300 python tools/tc_trace.py > $O/c_trace_dqw8.jsonl 2> $O/c_trace_dqw8.err
EETQ_B200_TC_DQW=16 timeout 300 python tools/tc_trace.py > $O/c_trace_dqw16.jsonl 2> $O/c_trace_dqw16.err
timeout 400 python tools/kbench.py --tc-only --out $O/c_kb_v2_dqw8.json > $O/c_kb_v2_dqw8.log 2>&1
EETQ_B200_TC_DQW=16 timeout 400 python tools/kbench.py --tc-only --out $O/c_kb_v2_dqw16.json > $O/c_kb_v2_dqw16.log 2>&1
EETQ_B200_TC_L2PROMO=128 timeout 400 python tools/kbench.py --tc-only --quick --out $O/c_kb_v2_promo128.json > $O/c_kb_v2_promo128.log 2>&1
KBENCH_TC_PDL=1 timeout 400 python tools/kbench.py --tc-only --quick --out $O/c_kb_v2_pdl.json > $O/c_kb_v2_pdl.log 2>&1
EETQ_B200_TC_NOSPLIT=1 timeout 400 python tools/kbench.py --tc-only --quick --out $O/c_kb_v2_nosplit.json > $O/c_kb_v2_nosplit.log 2>&1
tail -n 4 $O/c_tc_debug_dqw8.log; tail -n 3 $O/c_tc_debug_dqw16.log
cat $O/c_diag.log | cut -c1-260
grep gemm_tc $O/c_kb_v2_dqw8.log | cut -c1-230
cat $O/c_trace_dqw8.jsonl
