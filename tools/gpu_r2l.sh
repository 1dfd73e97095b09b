#!/bin/bash
# tcgen05 kernel: bench (with and without PDL), ncu --set full of 5 representative launches; decode timeline; M<=16 cut
mkdir -p gpurun_out
timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc.json > gpurun_out/kb_tc.log 2>&1
KBENCH_TC_PDL=1 timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc_pdl.json > gpurun_out/kb_tc_pdl.log 2>&1
EETQ_B200_TC_DQW=16 timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc_dqw16.json > gpurun_out/kb_tc_dqw16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -o gpurun_out/prof_tc -f python tools/prof_tc.py > gpurun_out/ncu_tc.log 2>&1
EETQ_B200_LIB=$PWD/eetq_b200/libeetq_b200_trace.so timeout 300 python tools/timeline.py --layers 4 > gpurun_out/timeline.log 2>&1
timeout 300 python tools/kbench_smallm.py > gpurun_out/kbench_smallm.log 2>&1
grep gemm_tc gpurun_out/kb_tc.log | cut -c1-50,95-260
tail -n 3 gpurun_out/ncu_tc.log
cut -c1-300 gpurun_out/timeline.log
