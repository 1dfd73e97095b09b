#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -q --timeout 300 > gpurun_out/t_decode.log 2>&1; tail -n 3 gpurun_out/t_decode.log
for cfg in "none 1" "ll 1" "ll 0" "ll 1"; do
  set -- $cfg
  EETQ_B200_LOCAL_EXCHANGE=$1 EETQ_B200_LL_NOWAIT=$2 timeout 300 python bench.py --skip-cpu-baseline --steps 128 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_x.json"))
    print("exchange=$1 nowait=$2", round(d["value"],1), "tok/s  e2e", round(d["e2e"]["value"],1), " gemv us/launch", round(d["roofline"]["us_per_launch"],2))
except Exception as e:
    print("exchange=$1 nowait=$2 FAILED", e); print(open("gpurun_out/bench_x.err").read()[-1500:])
PY
done
cp gpurun_out/bench_x.json gpurun_out/bench_n1.json
EETQ_B200_LIB=$PWD/eetq_b200/libeetq_b200_trace.so timeout 300 python tools/timeline.py --layers 4 > gpurun_out/timeline.log 2>&1
cut -c1-250 gpurun_out/timeline.log | sed -n 6,14p
