#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_quantize_gpu.py -q -x --timeout 120 > gpurun_out/t_gemm.log 2>&1; tail -n 3 gpurun_out/t_gemm.log
timeout 600 python tools/kbench.py --gemv-only --out gpurun_out/kb_gemv.json > gpurun_out/kb_gemv.log 2>&1
python - <<'PY'
import json
rows=json.load(open('gpurun_out/kb_gemv.json'))
ref={(r['K'],r['N'],r['M']):r['us'] for r in rows if r['kernel']=='reference_gemv_sm100a'}
for r in rows:
    if r['kernel']=='gemv':
        k=(r['K'],r['N'],r['M']); print('K=%5d N=%5d M=%d pdl=%-5s ours=%6.2f ref=%6.2f %s'%(r['K'],r['N'],r['M'],r['pdl'],r['us'],ref.get(k,0),'WIN' if r['us']<=ref.get(k,0) else 'lose'))
PY
