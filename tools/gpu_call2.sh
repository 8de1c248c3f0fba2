#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_general.py -q -x -s > gpurun_out/r02_tc_tests_v2.log 2>&1; echo "tc pytest rc=$?"
tail -30 gpurun_out/r02_tc_tests_v2.log
timeout 300 python tests/_tc_time.py > gpurun_out/r02_tc_time_v2.txt 2>&1; echo "tc_time rc=$?"; cat gpurun_out/r02_tc_time_v2.txt
