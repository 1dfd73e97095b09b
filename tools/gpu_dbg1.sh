#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/dbg1.log
{
for sh in "64 4096 11008" "256 4096 11008" "1024 4096 4096" "1024 11008 4096" "64 4096 4096"; do
  timeout 100 python tools/tc_diag.py $sh
  EETQ_B200_TC_NOSPLIT=1 timeout 100 python tools/tc_diag.py $sh
done
EETQ_B200_TC_BT=128 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=64 timeout 100 python tools/tc_diag.py 1024 4096 4096
EETQ_B200_TC_BT=16 timeout 100 python tools/tc_diag.py 64 4096 11008
EETQ_B200_TC_BT=32 timeout 100 python tools/tc_diag.py 64 4096 11008
EETQ_B200_TC_DQW=16 timeout 100 python tools/tc_diag.py 64 4096 11008
} > $O 2>&1
cut -c1-400 $O
