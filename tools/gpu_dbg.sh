#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_general.py tests/test_gpu_ps.py -q -x > gpurun_out/r02_v50_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_v50_tests.log
timeout 600 python tools/prof_cfg5_wide.py 2>&1 | grep -v Warn | tail -34
