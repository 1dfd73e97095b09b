#!/bin/bash
# round 2, call D: tcgen05 fixes (call C content) + the rewritten decode path on one GPU
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
timeout 900 python -m pytest tests/test_decode_gpu.py -q --timeout 300 --timeout-method=thread > $O/d_t_decode.log 2>&1; echo "rc=$?" >> $O/d_t_decode.log
timeout 600 python bench.py --skip-cpu-baseline --steps 64 > $O/d_bench_n1.json 2> $O/d_bench_n1.err; echo "rc=$?" >> $O/d_bench_n1.err
tail -n 4 $O/c_tc_debug_dqw8.log; tail -n 3 $O/c_tc_debug_dqw16.log
cat $O/c_diag.log | cut -c1-260
grep gemm_tc $O/c_kb_v2_dqw8.log | cut -c1-230
cat $O/c_trace_dqw8.jsonl
tail -n 30 $O/d_t_decode.log
cat $O/d_bench_n1.json | cut -c1-1500; tail -n 5 $O/d_bench_n1.err
