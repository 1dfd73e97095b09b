#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_tc -f python tools/prof_tc.py > gpurun_out/ncu_tc.log 2>&1
tail -n 3 gpurun_out/ncu_tc.log
