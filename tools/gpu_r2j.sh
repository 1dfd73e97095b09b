#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 12" "2 32" "2 48" "2 64" "2 96"; do
  set -- $cfg
  EETQ_B200_GEMV_L2PREFETCH=$1 EETQ_B200_L2_NEXT_MB=$2 timeout 300 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_l2_$1_$2.json 2> gpurun_out/bench_l2_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_l2_$1_$2.json"))
print("knob=$1 next_mb=$2", round(d["value"],1), "tok/s  gemv us/launch", round(d["roofline"]["us_per_launch"],2), "frac", round(d["roofline"]["frac"],3))
PY
done
