#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 120 > gpurun_out/t_gemm.log 2>&1; tail -n 3 gpurun_out/t_gemm.log
timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc.json > gpurun_out/kb_tc.log 2>&1
KBENCH_TC_PDL=1 timeout 400 python tools/kbench.py --tc-only --out gpurun_out/kb_tc_pdl.json > gpurun_out/kb_tc_pdl.log 2>&1
timeout 300 python tools/tc_trace.py > gpurun_out/tc_trace.jsonl 2> gpurun_out/tc_trace.err
for f in kb_tc kb_tc_pdl; do echo $f; grep gemm_tc gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(' K=%5d N=%5d M=%4d us=%7.2f TF=%7.1f GB/s=%7.1f'%(r['K'],r['N'],r['M'],r['us'],r['tflops'],r['gbs']))
"; done
python - <<'PY'
import json
for l in open('gpurun_out/tc_trace.jsonl'):
    r=json.loads(l); print(r['M'],r['K'],r['N'],'units',r['units_per_cta'],'first_tma',r['first_w_tma'],'dq0',r['dequant_stage_done'][0],'mma_done',r['mma_done'],'epi',r['epi_seg'][0],'fixup',r['joint_fixup'],'end',r['cta_end'],r['cta_end_max'])
PY
timeout 300 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('decode', round(d['value'],1), 'prefill_ms', round(d['prefill_ms'],2), 'cold', round(d['prefill_cold_ms'],1))"
