#!/bin/bash
mkdir -p gpurun_out
SPN_MBAR_HINT_NS=0 python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tsum_kernel -c 2 -o gpurun_out/r02_tsum_v13 python tools/tc_one.py > gpurun_out/r02_ncu_v13.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r02_ncu_v13.log
ls -la gpurun_out/*.ncu-rep
