#!/bin/bash
# N-GPU bench under a few data-parallel settings (overlap on/off, NCCL CTA cap); prints value / ms per step.
N=${1:-2}
run() {
  echo "== $*"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']))"
}
run SPN_DP_OVERLAP=1
run SPN_DP_OVERLAP=0
run SPN_DP_OVERLAP=1 NCCL_MAX_CTAS=4
run SPN_DP_OVERLAP=0 NCCL_MAX_CTAS=4
run SPN_DP_OVERLAP=1 NCCL_MAX_CTAS=16 NCCL_MIN_CTAS=16
