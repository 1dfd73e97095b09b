#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kbench.py --gemv-only --out gpurun_out/kbench_gemv.json > gpurun_out/kbench_gemv.log 2>&1
grep gemv gpurun_out/kbench_gemv.log | cut -c1-200
