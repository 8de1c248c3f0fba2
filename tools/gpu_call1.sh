#!/bin/bash
# round-2 GPU call 1: micro-benchmarks, full GPU test-suite, bench (incl. secondary configs), ncu of the CSPN / sampling kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
nproc >> gpurun_out/r02_gpu.txt; free -g >> gpurun_out/r02_gpu.txt
timeout 120 ./tools/bench/pipe_bench > gpurun_out/r02_pipe_bench.txt 2>&1; echo "pipe_bench rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu_tests_v1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_gpu_tests_v1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; echo "bench rc=$?"
head -c 1500 gpurun_out/r02_bench_v1.json; echo
timeout 600 ncu --set full --clock-control none -k regex:'sum_ps|leaf_ps|sample_' -c 60 -o gpurun_out/r02_families python tools/ncu_families.py 128 > gpurun_out/r02_ncu_families.log 2>&1; echo "ncu rc=$?"
timeout 120 python tools/ncu_summary.py gpurun_out/r02_families.ncu-rep gpurun_out/r02_ncu_families_summary.csv > /dev/null 2>&1; echo "summary rc=$?"
ls -la gpurun_out
