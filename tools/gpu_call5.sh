#!/bin/bash
mkdir -p gpurun_out
SPN_TC_PROFILE=1 python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "profile build rc=$?"; tail -3 gpurun_out/build.log
SPN_TC_PROFILE=1 timeout 300 python tests/_tc_time.py > gpurun_out/r02_tc_prof_v5.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_tc_prof_v5.txt
