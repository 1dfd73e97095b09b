#!/bin/bash
# int4 bring-up on one B200: parity tests of the packed-int4 path, its micro-benchmark, then the rest of the GPU suite
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_int4_gpu.py -m gpu -q --timeout 120 --timeout-method=thread -x > $O/t_int4.log 2>&1; echo "pytest rc=$?" >> $O/t_int4.log
tail -n 25 $O/t_int4.log
timeout 200 python tools/kbench_int4.py > $O/kbench_int4.log 2>&1; echo "rc=$?" >> $O/kbench_int4.log
cat $O/kbench_int4.log
timeout 700 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --deselect tests/test_int4_gpu.py --durations=8 > $O/t_rest.log 2>&1; echo "pytest rc=$?" >> $O/t_rest.log
tail -n 16 $O/t_rest.log
