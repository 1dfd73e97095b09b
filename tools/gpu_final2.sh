#!/bin/bash
# end-of-round validation of the committed state on one B200 (bounded: this round's remaining GPU budget is a few minutes)
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread > $O/t_all.log 2>&1; echo "pytest rc=$?" >> $O/t_all.log
tail -n 5 $O/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -n 2 $O/smoke.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
cat $O/bench_n1.json; tail -n 2 $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
cut -c1-600 $O/bench_ref.json; tail -n 2 $O/bench_ref.err
timeout 200 python tests/perf/kbench_mma2.py > $O/kbench_mma2.log 2>&1; echo "rc=$?" >> $O/kbench_mma2.log
tail -n 2 $O/kbench_mma2.log
