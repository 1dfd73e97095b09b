#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 16 0" "1 12 2" "1 24 2" "0 16 0" "1 12 2" "1 6 2" "1 4 0" "1 2 0"; do
  set -- $cfg
  EETQ_B200_L2_NEXT=$1 EETQ_B200_L2_NEXT_MB=$2 EETQ_B200_L2_NEXT_WHEN=$3 timeout 300 python bench.py --skip-cpu-baseline --steps 128 > gpurun_out/bench_l2n.json 2> gpurun_out/bench_l2n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_l2n.json"))
print("l2_next=$1 mb=$2 when=$3", round(d["value"],1), "tok/s  e2e", round(d["e2e"]["value"],1), " gemv us/launch", round(d["roofline"]["us_per_launch"],2))
PY
done
