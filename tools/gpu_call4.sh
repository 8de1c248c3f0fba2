#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_general.py -q -x > gpurun_out/r02_tc_tests_v4.log 2>&1; echo "tc pytest rc=$?"
tail -5 gpurun_out/r02_tc_tests_v4.log
timeout 300 python tests/_tc_time.py > gpurun_out/r02_tc_time_v4.txt 2>&1; echo "tc_time rc=$?"; cat gpurun_out/r02_tc_time_v4.txt
SPN_TC_PROFILE=1 python conditional-ratspn-for-rl_b200/build.py --force > /dev/null 2>&1; echo "profile build rc=$?"
SPN_TC_PROFILE=1 timeout 300 python tests/_tc_time.py > gpurun_out/r02_tc_prof_v4.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_tc_prof_v4.txt
