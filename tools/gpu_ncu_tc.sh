#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02_tsum_v38}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsum_ -c 2 -o gpurun_out/$TAG python tools/tc_one.py > gpurun_out/${TAG}.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out/$TAG.ncu-rep
