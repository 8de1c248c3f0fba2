#!/bin/bash
mkdir -p gpurun_out
SPN_TC_DEBUGSW=1 SPN_TC_PROFILE=1 python conditional-ratspn-for-rl_b200/build.py --force > gpurun_out/build.log 2>&1; echo "build rc=$?"
for d in 0 2 4; do echo "== SPN_TC_DEBUG=$d"; SPN_TC_PROFILE=1 SPN_TC_DEBUG=$d timeout 300 python tests/_tc_time.py bwdx 2>&1 | grep -E "tc |clk per tile"; done > gpurun_out/r02_tc_prof_v11.txt 2>&1
cat gpurun_out/r02_tc_prof_v11.txt
