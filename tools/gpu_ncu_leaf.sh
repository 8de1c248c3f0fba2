#!/bin/bash
# ncu --set full of the leaf kernels (tensor-core and exact) at the cfg 2 shape + per-phase clocks of the experiment build
mkdir -p gpurun_out
TAG=${1:-r02_leaf}
timeout 300 python tools/leaf_tc_one.py > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'leaf_' -c 8 -o gpurun_out/${TAG} python tools/leaf_tc_one.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}.ncu-rep
python tools/ncu_summary.py gpurun_out/${TAG}.ncu-rep gpurun_out/${TAG}_summary.csv | cut -c1-220
SPN_LIB_PATH=build/variants/libspn_prof.so timeout 120 python tools/leaf_tc_prof.py 0 31 > gpurun_out/${TAG}_prof.log 2>&1; cat gpurun_out/${TAG}_prof.log | tail -18
