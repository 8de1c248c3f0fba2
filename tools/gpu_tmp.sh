#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_leaf.py tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py tests/test_gpu_tc.py -q -x > gpurun_out/r02_v5_leaf_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_v5_leaf_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_v5_bench_quick.json 2> gpurun_out/r02_v5_bench_quick.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_v5_bench_quick.json')); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
timeout 300 python tools/step_profile.py 2>&1 | grep -i "leaf\|total" | head -6
