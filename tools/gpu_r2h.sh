#!/bin/bash
mkdir -p gpurun_out
EETQ_B200_LIB=$PWD/eetq_b200/libeetq_b200_trace.so timeout 300 python tools/timeline.py --layers 4 > gpurun_out/timeline.log 2>&1
cat gpurun_out/timeline.log | cut -c1-330
timeout 600 python bench.py --skip-cpu-baseline --steps 64 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
