#!/bin/bash
# full cycle: GPU test-suite, bench, ncu launch list of the bench command, ncu --set full of the tensor-core kernels of one step
mkdir -p gpurun_out
TAG=${1:-r02_v2}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
for k,v in d['tc_kernels'].items(): print(k, v.get('ms'), v.get('frac'))
"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'tsum_|sum_tc_bwd|tc_gprep|tc_dw_finish|tc_wimage|tc_logz' -c 14 -o gpurun_out/${TAG}_tc python tools/one_step.py > gpurun_out/${TAG}_ncu_tc.log 2>&1; echo "ncu full rc=$?"
timeout 120 python tools/ncu_summary.py gpurun_out/${TAG}_tc.ncu-rep gpurun_out/${TAG}_ncu_tc_summary.csv > /dev/null 2>&1; echo "summary rc=$?"
