#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api_rows.py tests/test_checkpoint_interchange.py tests/test_gpu_tc_general.py tests/test_gpu_tc.py -q -m gpu > gpurun_out/r02_newtests_v40.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert" gpurun_out/r02_newtests_v40.log | head -30
bash tools/gpu_variants.sh r02_tc_v40_time
