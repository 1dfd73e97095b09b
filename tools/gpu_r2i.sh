#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -q --timeout 300 > gpurun_out/t_decode.log 2>&1; tail -n 3 gpurun_out/t_decode.log
for cfg in "0 24" "1 24" "3 8" "3 24" "3 48" "2 24"; do
  set -- $cfg
  EETQ_B200_GEMV_L2PREFETCH=$1 EETQ_B200_L2_NEXT_MB=$2 timeout 300 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_l2_$1_$2.json 2> gpurun_out/bench_l2_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_l2_$1_$2.json"))
print("knob=$1 next_mb=$2", round(d["value"],1), "tok/s  gemv us/launch", round(d["roofline"]["us_per_launch"],2), "frac", round(d["roofline"]["frac"],3))
PY
done
