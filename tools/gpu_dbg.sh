#!/bin/bash
mkdir -p gpurun_out
for args in "24 2 3 257 24" "24 4 40 384 24" "40 3 40 640" "16 3 5 700" "64 3 2 600" "8 4 3 500" "56 2 3 300" "20 2 3 257 24" "40 2 3 100 40"; do timeout 120 python tools/tc_dbg2.py $args 2>&1 | grep "iter" | cut -c1-100; done > gpurun_out/r02_tc_dbg_v37.txt 2>&1
echo "pass: $(grep -c 'bad 0' gpurun_out/r02_tc_dbg_v37.txt) fail: $(grep -vc 'bad 0' gpurun_out/r02_tc_dbg_v37.txt)"; grep -v "bad 0" gpurun_out/r02_tc_dbg_v37.txt | head -5
bash tools/gpu_tc_cycle.sh r02_tc_v37
