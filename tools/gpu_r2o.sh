#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py > gpurun_out/tc_trace.jsonl 2> gpurun_out/tc_trace.err
EETQ_B200_TC_BT=128 timeout 400 python tools/kbench.py --tc-only --quick --out gpurun_out/kb_tc_bt128.json > gpurun_out/kb_tc_bt128.log 2>&1
timeout 400 python tools/kbench.py --tc-only --quick --out gpurun_out/kb_tc_q.json > gpurun_out/kb_tc_q.log 2>&1
EETQ_B200_TC_NOSPLIT=1 timeout 400 python tools/kbench.py --tc-only --quick --out gpurun_out/kb_tc_nosplit.json > gpurun_out/kb_tc_nosplit.log 2>&1
cat gpurun_out/tc_trace.jsonl
for f in kb_tc_q kb_tc_bt128 kb_tc_nosplit; do echo $f; grep gemm_tc gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(' K=%5d N=%5d M=%4d us=%7.2f TF=%7.1f GB/s=%7.1f'%(r['K'],r['N'],r['M'],r['us'],r['tflops'],r['gbs']))
"; done
