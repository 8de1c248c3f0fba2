#!/bin/bash
mkdir -p gpurun_out
for h in 100000 2000 100; do
  SPN_MBAR_HINT_NS=$h python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "hint $h build rc=$?"
  timeout 300 python tests/_tc_time.py 2>&1 | grep -E "tc |aborts"
done > gpurun_out/r02_tc_hint_v10.txt 2>&1
cat gpurun_out/r02_tc_hint_v10.txt
