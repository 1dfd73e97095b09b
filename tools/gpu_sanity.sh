#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
tail -n 3 gpurun_out/t_all.log gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_n1.json; tail -n 2 gpurun_out/bench_n1.err
