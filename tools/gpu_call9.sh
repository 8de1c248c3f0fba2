#!/bin/bash
mkdir -p gpurun_out
SPN_TC_DEBUGSW=1 python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "debugsw build rc=$?"; tail -3 gpurun_out/build.log
for d in 0 1 2 3 4; do echo "== SPN_TC_DEBUG=$d"; SPN_TC_DEBUG=$d timeout 300 python tests/_tc_time.py bwdx 2>&1 | grep -E "tc |aborts"; SPN_TC_DEBUG=$d timeout 300 python tests/_tc_time.py fwd 2>&1 | grep -E "tc |aborts"; done > gpurun_out/r02_tc_dbg_v9.txt 2>&1
cat gpurun_out/r02_tc_dbg_v9.txt
