import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import tools.bench_configs as bc
import ratspn_b200 as rb
# rebuild the cfg5-narrow model via the bench helper internals
res = bc.cfg5_narrow.__code__
import types
# simplest: run the bench function under the profiler
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = bc.cfg5_narrow()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
for k, ms, n in rows[:12]:
    print(f"{ms:9.3f} ms total x{n:<3d} {ms/n:8.3f} ms each  {k[:100]}")
print(out)
