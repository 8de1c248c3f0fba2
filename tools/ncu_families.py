"""One invocation of each kernel family that bench.py's headline line does not cover, for ncu captures (per-kernel DRAM
bytes / dram_throughput_pct asked for by the north star): the per-conditional (CSPN) streaming kernels sum_ps_* and
leaf_ps_* at the cfg 5-narrow shape, and the sampling kernels sample_sum_* / sample_leaf_* at the cfg 4 shape.

    ncu --set full --clock-control none -k regex:'sum_ps|leaf_ps|sample_' -o OUT python tools/ncu_families.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bc)
def _once(fn, reps=1, warmup=0):   # run every workload exactly once
    fn()
    return 1.0, 1.0


bc.timed = _once

w = int(sys.argv[1]) if len(sys.argv) > 1 else 256
print(bc.cfg5_narrow_train(w))
print(bc.cfg5_narrow(w))
m = bc.ratspn(784, 5, 40, 40, 40)
with torch.no_grad():
    m.sample(mode="index", n=16384)
    m.sample(mode="index", n=16384, is_mpe=True)
torch.cuda.synchronize()
