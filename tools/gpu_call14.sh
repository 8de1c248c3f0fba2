#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests_v14.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_gpu_tests_v14.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v14.json 2> gpurun_out/r02_bench_v14.err; echo "bench rc=$?"
head -c 2500 gpurun_out/r02_bench_v14.json; echo
