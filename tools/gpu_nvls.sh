#!/bin/bash
# A/B of the NCCL all-reduce algorithm for the flat gradient exchange at N GPUs: default choice vs NVLS (in-switch reduction)
N=$1; mkdir -p gpurun_out
for algo in NVLS ""; do
  echo "== NCCL_ALGO=${algo:-default}"
  env ${algo:+NCCL_ALGO=$algo} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2> gpurun_out/_nvls.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('weak', round(d['value']), round(d['ms_per_step'], 3))"
  env ${algo:+NCCL_ALGO=$algo} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-cpu-baseline --scaling strong 2>> gpurun_out/_nvls.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('strong', round(d['value']), round(d['ms_per_step'], 3))"
done 2>&1 | tee gpurun_out/r02_nvls_n${N}.txt
tail -3 gpurun_out/_nvls.err
