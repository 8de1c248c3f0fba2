#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 120 python tools/tc_dbg2.py 24 2 3 257 24 2>&1 | grep "^iter" | cut -c1-60;  timeout 120 python tools/tc_dbg2.py 24 4 40 384 24 2>&1 | grep "^iter" | cut -c1-60; done > gpurun_out/r02_tc_dbg_v35.txt 2>&1
for args in "40 3 40 640" "16 3 5 700" "64 3 2 600" "8 4 3 500" "56 2 3 300" "20 2 3 257 24"; do timeout 120 python tools/tc_dbg2.py $args 2>&1 | grep "^iter" | cut -c1-60; done >> gpurun_out/r02_tc_dbg_v35.txt 2>&1
echo "pass: $(grep -c 'bad 0' gpurun_out/r02_tc_dbg_v35.txt) fail: $(grep -vc 'bad 0' gpurun_out/r02_tc_dbg_v35.txt)"
bash tools/gpu_tc_cycle.sh r02_tc_v35
