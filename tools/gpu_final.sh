#!/bin/bash
# full validation of the committed state: parity suite, smoke, both bench arms, launch list, small-M GEMV vs tcgen05
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 16 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
timeout 300 python tools/kbench_smallm.py > gpurun_out/kbench_smallm.log 2>&1
BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/t_all.log gpurun_out/smoke.log
cat gpurun_out/bench_ref.json; tail -n 2 gpurun_out/bench_ref.err
cat gpurun_out/bench_n1.json; tail -n 2 gpurun_out/bench_n1.err
cat gpurun_out/kbench_smallm.log
