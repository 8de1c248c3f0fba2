#!/bin/bash
mkdir -p gpurun_out
SPN_TC_PROFILE=1 python conditional-ratspn-for-rl_b200/build.py --force > /dev/null 2>&1; echo "profile build rc=$?"
SPN_TC_PROFILE=1 timeout 300 python tests/_tc_time.py > gpurun_out/r02_tc_prof_v3.txt 2>&1; echo "rc=$?"; cat gpurun_out/r02_tc_prof_v3.txt
