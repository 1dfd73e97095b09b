#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 300 --timeout-method=thread -x -k "decode or gemv or fused or bf16" > gpurun_out/t_dec.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_dec.log
EETQ_B200_GEMV_LOWREG=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 300 --timeout-method=thread -x -k "test_gemv_matches_oracle" > gpurun_out/t_low.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_low.log
run() { echo "### $*" >> gpurun_out/chain.log; env "$@" timeout 300 python tools/gemv_chain.py >> gpurun_out/chain.log 2>&1; }
: > gpurun_out/chain.log
run A=base
run CHAIN_PDL=0
run EETQ_B200_GEMV_CTAS=3
run EETQ_B200_GEMV_CTAS=4
run EETQ_B200_GEMV_LOWREG=1
run EETQ_B200_GEMV_LOWREG=1 EETQ_B200_GEMV_CTAS=3
run EETQ_B200_GEMV_LOWREG=1 EETQ_B200_GEMV_CTAS=4
run EETQ_B200_GEMV_IMPL=tma
run EETQ_B200_GEMV_IMPL=tma EETQ_B200_GEMV_CTAS=2
run EETQ_B200_GEMV_IMPL=tma EETQ_B200_GEMV_CTAS=2 CHAIN_PDL=0
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
BENCH_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -n 4 gpurun_out/t_dec.log gpurun_out/t_low.log
cat gpurun_out/chain.log | grep -v "^$"
cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
