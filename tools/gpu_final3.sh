#!/bin/bash
# last call of the round: full parity suite on the final build, decode-row A/B, ncu of the final decode-row kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread > $O/t_all.log 2>&1; echo "pytest rc=$?" >> $O/t_all.log
tail -n 4 $O/t_all.log
timeout 200 python tools/kbench_mma2.py > $O/kbench_mma2.log 2>&1; echo "rc=$?" >> $O/kbench_mma2.log
grep '"mma2"' $O/kbench_mma2.log | grep -v '"M": 3' | cut -c1-140
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/prof_int4_final -f python tools/prof_int4.py > $O/ncu_int4_final.log 2>&1; echo "rc=$?" >> $O/ncu_int4_final.log
tail -n 2 $O/ncu_int4_final.log
