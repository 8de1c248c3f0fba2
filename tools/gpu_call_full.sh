#!/bin/bash
# full cycle: smoke, GPU test-suite, bench (with the secondary configs), ncu launch list of the bench command
mkdir -p gpurun_out
TAG=${1:-r02_v2}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'traffic', d['roofline']['traffic'])
for k,v in d['tc_kernels'].items(): print(k, v.get('ms'), v.get('frac'))
print({k: (v.get('forward', v).get('ms') if isinstance(v, dict) else v) for k, v in d['extra'].items()})
"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"; head -c 400 gpurun_out/${TAG}_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
