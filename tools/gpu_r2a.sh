#!/bin/bash
# round 2, call A: new tcgen05 kernel (TMEM A operand, persistent stream-K) -- correctness, A/B against the round-1 kernel, timeline
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_a.txt 2>&1
timeout 200 python tools/tc_debug.py > $O/a_tc_debug_dqw8.log 2>&1; echo "rc=$?" >> $O/a_tc_debug_dqw8.log
EETQ_B200_TC_DQW=16 timeout 200 python tools/tc_debug.py > $O/a_tc_debug_dqw16.log 2>&1; echo "rc=$?" >> $O/a_tc_debug_dqw16.log
if grep -q "TC_DEBUG PASS" $O/a_tc_debug_dqw8.log; then
  timeout 900 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 300 --timeout-method=thread > $O/a_t_gemm.log 2>&1; echo "rc=$?" >> $O/a_t_gemm.log
  timeout 300 python tools/tc_trace.py > $O/a_trace_dqw8.jsonl 2> $O/a_trace_dqw8.err
  EETQ_B200_TC_DQW=16 timeout 300 python tools/tc_trace.py > $O/a_trace_dqw16.jsonl 2> $O/a_trace_dqw16.err
  timeout 400 python tools/kbench.py --tc-only --out $O/a_kb_v2_dqw8.json > $O/a_kb_v2_dqw8.log 2>&1
  EETQ_B200_TC_DQW=16 timeout 400 python tools/kbench.py --tc-only --out $O/a_kb_v2_dqw16.json > $O/a_kb_v2_dqw16.log 2>&1
  EETQ_B200_TC_L2PROMO=128 timeout 400 python tools/kbench.py --tc-only --quick --out $O/a_kb_v2_promo128.json > $O/a_kb_v2_promo128.log 2>&1
  KBENCH_TC_PDL=1 timeout 400 python tools/kbench.py --tc-only --quick --out $O/a_kb_v2_pdl.json > $O/a_kb_v2_pdl.log 2>&1
fi
EETQ_B200_TC_IMPL=v1 timeout 400 python tools/kbench.py --tc-only --quick --out $O/a_kb_v1.json > $O/a_kb_v1.log 2>&1
timeout 180 compute-sanitizer --tool racecheck python tools/tc_debug.py --tiny > $O/a_racecheck.log 2>&1; echo "rc=$?" >> $O/a_racecheck.log
tail -n 20 $O/a_tc_debug_dqw8.log; tail -n 4 $O/a_tc_debug_dqw16.log; tail -n 5 $O/a_t_gemm.log
grep gemm_tc $O/a_kb_v2_dqw8.log | cut -c1-220
grep gemm_tc $O/a_kb_v1.log | cut -c1-220
