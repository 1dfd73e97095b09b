#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -q --timeout 300 > gpurun_out/t_decode.log 2>&1; tail -n 3 gpurun_out/t_decode.log
timeout 300 python bench.py --skip-cpu-baseline --steps 128 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-200 gpurun_out/bench_n1.json; tail -n 2 gpurun_out/bench_n1.err
EETQ_B200_LIB=$PWD/eetq_b200/libeetq_b200_trace.so timeout 300 python tools/timeline.py --layers 4 > gpurun_out/timeline.log 2>&1
grep -E "attn|lm_head" gpurun_out/timeline.log | cut -c1-300
