#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'sum_ps|leaf_ps|sample_' -c 30 -o gpurun_out/r02_families python tools/ncu_families.py 128 > gpurun_out/r02_ncu_families.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_ncu_families.log
timeout 120 python tools/ncu_summary.py gpurun_out/r02_families.ncu-rep gpurun_out/r02_ncu_families_summary.csv | cut -c1-170 | head -34
