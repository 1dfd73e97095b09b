#!/bin/bash
# full validation of the committed state on one B200 (gpurun -- bash tools/gpu_final.sh)
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > $O/t_all.log 2>&1; echo "pytest rc=$?" >> $O/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --impl reference --steps 16 --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
# launch list of one decode step (cold-cache, serialised: compare shares)
BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > $O/ncu_launches.log 2>&1
# DRAM traffic of the decode GEMVs: two whole layers (8 launches) of the timed step
BENCH_PROFILE_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:w8a16_gemv_kernel -c 8 \
    -o $O/prof_gemv -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --layers 2 > $O/ncu_gemv.log 2>&1
BENCH_PROFILE_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_decode -c 2 \
    -o $O/prof_attn -f python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --layers 2 > $O/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -o $O/prof_tc -f python tools/prof_tc.py > $O/ncu_tc.log 2>&1
timeout 400 python tests/perf/kbench.py --tc-only --out $O/kb_tc.json > $O/kb_tc.log 2>&1
timeout 300 python tests/perf/kbench_smallm.py > $O/kbench_smallm.log 2>&1
EETQ_B200_LIB=$PWD/eetq_b200/libeetq_b200_trace.so timeout 300 python tools/timeline.py --layers 4 > $O/timeline.log 2>&1
timeout 200 compute-sanitizer --tool racecheck python tests/diag/tc_debug.py --tiny > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
timeout 200 compute-sanitizer --tool synccheck python tests/diag/tc_debug.py --tiny > $O/sanitizer_synccheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_synccheck.log
tail -n 3 $O/t_all.log $O/smoke.log
cut -c1-700 $O/bench_ref.json; tail -n 2 $O/bench_ref.err
cat $O/bench_n1.json; tail -n 2 $O/bench_n1.err
cat $O/kbench_smallm.log
tail -n 4 $O/sanitizer_racecheck.log $O/sanitizer_synccheck.log
