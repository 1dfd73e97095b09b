#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/t_dec.log 2>&1; echo "rc=$?" >> gpurun_out/t_dec.log
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_kvpf.json 2> gpurun_out/bench_kvpf.err; echo "rc=$?" >> gpurun_out/bench_kvpf.err
EETQ_B200_KV_PREFETCH=0 timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench_nokvpf.json 2> gpurun_out/bench_nokvpf.err
tail -n 3 gpurun_out/t_dec.log; cut -c1-260 gpurun_out/bench_kvpf.json; tail -n 2 gpurun_out/bench_kvpf.err; cut -c1-260 gpurun_out/bench_nokvpf.json
