#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/t_dec.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_dec.log
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_chain.json 2> gpurun_out/bench_chain.err; echo "rc=$?" >> gpurun_out/bench_chain.err
EETQ_B200_CHAIN=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_nochain.json 2> gpurun_out/bench_nochain.err; echo "rc=$?" >> gpurun_out/bench_nochain.err
tail -n 5 gpurun_out/t_dec.log; cat gpurun_out/bench_chain.json; tail -n 3 gpurun_out/bench_chain.err; cat gpurun_out/bench_nochain.json; tail -n 2 gpurun_out/bench_nochain.err
