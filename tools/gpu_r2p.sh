#!/bin/bash
mkdir -p gpurun_out
EETQ_B200_TC_NOSPLIT=1 timeout 300 python tools/tc_trace.py > gpurun_out/tc_trace_nosplit.jsonl 2> gpurun_out/tc_trace_nosplit.err
cat gpurun_out/tc_trace_nosplit.jsonl
