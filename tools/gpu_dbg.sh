#!/bin/bash
SPN_TC_PROFILE=1 SPN_LIB_PATH=$PWD/build/variants/libspn_prof.so timeout 300 python tests/_tc_time.py bwdx 2>&1 | tail -1
timeout 300 python tests/_tc_time.py bwdx 2>&1 | grep "^tc"
SPN_LIB_PATH=$PWD/build/variants/libspn_a_old.so timeout 300 python tests/_tc_time.py bwdx 2>&1 | grep "^tc"
timeout 300 python tests/_tc_time.py bwdx 2>&1 | grep "^tc"
for args in "24 2 3 257 24" "40 3 40 640" "64 3 2 600"; do timeout 120 python tools/tc_dbg2.py $args 2>&1 | grep "^iter" | cut -c1-100; done > gpurun_out/r02_tc_dbg_v44.txt 2>&1
echo "pass: $(grep -c 'bad 0' gpurun_out/r02_tc_dbg_v44.txt) of $(wc -l < gpurun_out/r02_tc_dbg_v44.txt)"
