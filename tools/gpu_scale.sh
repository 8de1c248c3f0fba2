#!/bin/bash
# weak + strong scaling of the bench on N GPUs of one box (run with gpurun --gpus N): bash tools/gpu_scale.sh N TAG
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-cpu-baseline "$@"; }
run > gpurun_out/${TAG}_bench_n${N}_weak.json 2> gpurun_out/${TAG}_bench_n${N}_weak.err; echo "weak rc=$?"
run --scaling strong > gpurun_out/${TAG}_bench_n${N}_strong.json 2> gpurun_out/${TAG}_bench_n${N}_strong.err; echo "strong rc=$?"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 2 --warmup 3 --no-extra --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_nccl_n${N}.log; echo "nccl rc=$?"
{ grep -E "NVLS multicast support|Connected all|nRanks" gpurun_out/${TAG}_nccl_n${N}.log | grep "\[0\]" | head -6; grep -E "AllReduce: " gpurun_out/${TAG}_nccl_n${N}.log | grep "\[0\]" | sed -E 's/^.*NCCL INFO //' | sort | uniq -c | sort -rn | head -20; } > gpurun_out/${TAG}_nccl_n${N}_summary.txt
rm -f gpurun_out/${TAG}_nccl_n${N}.log
python - <<PY
import json
for k in ("weak","strong"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_n${N}_%s.json" % k))
        print(k, "N=", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "samples/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"])
    except Exception as e: print(k, "failed", e)
PY
cat gpurun_out/${TAG}_nccl_n${N}_summary.txt | cut -c1-160
