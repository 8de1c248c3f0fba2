#!/bin/bash
# ncu --set full of the tensor-core kernels (forward and backward, all four Sum layers) of one training step
mkdir -p gpurun_out
TAG=${1:-r02_step}
timeout 900 ncu --set full --clock-control none -k regex:'tsum_|sum_tc_bwd|tc_gprep|tc_dw_finish|tc_wimage|tc_logz' -c 48 -o gpurun_out/${TAG} python tools/one_step.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}.ncu-rep
