#!/bin/bash
# second GPU pass: full parity suite, decode tests, bench, kernel microbench, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 python bench.py --no-pdl --skip-cpu-baseline > gpurun_out/bench_n1_nopdl.json 2> gpurun_out/bench_n1_nopdl.err
EETQ_B200_GEMV_IMPL=ldg timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1_ldg.json 2> gpurun_out/bench_n1_ldg.err
timeout 600 python tools/kbench.py --quick > gpurun_out/kbench.log 2>&1
# ncu: launch list of one decode step (timed range only), then full-set capture of the GEMVs and the standalone kernel set
BENCH_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels -f \
    python tools/prof_kernels.py > gpurun_out/ncu_prof.log 2>&1
for f in t_all smoke; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done
echo "== bench"; cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1_nopdl.json; cat gpurun_out/bench_n1_ldg.json
echo "== kbench"; cat gpurun_out/kbench.log | tail -n 30
