#!/bin/bash
# tensor-core kernel cycle: parity tests of the tcgen05 path, then kernel timings (product build + any experiment builds)
mkdir -p gpurun_out
TAG=${1:-r02_tc}
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_general.py -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tc pytest rc=$?"
tail -5 gpurun_out/${TAG}_tests.log
bash tools/gpu_variants.sh ${TAG}_time all
