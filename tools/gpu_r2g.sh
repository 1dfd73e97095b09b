#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/dbg4.log
{
for sh in "256 4096 11008" "1024 4096 4096" "1024 11008 4096" "1024 4096 11008"; do
  timeout 100 python tools/tc_diag.py $sh
done
EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=64 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_DQW=16 timeout 100 python tools/tc_diag.py 1024 4096 4096
} > $O 2>&1
cut -c1-200 $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -n 6 gpurun_out/t_all.log
