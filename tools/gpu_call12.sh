#!/bin/bash
mkdir -p gpurun_out
SPN_MBAR_HINT_NS=0 python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "hint 0 build rc=$?"; tail -2 gpurun_out/build.log
timeout 300 python tests/_tc_time.py 2>&1 | grep -E "tc |aborts| ms x" > gpurun_out/r02_tc_nohint_v12.txt 2>&1
cat gpurun_out/r02_tc_nohint_v12.txt
