"""Per-kernel time table of one bench step (torch.profiler / CUPTI; not a bench number)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH_PER_GPU
x = torch.rand(B, 1, bench.CFG["F"], 1, 1, device=dev)
def step():
    for p in model.parameters():
        p.grad = None
    loss = -model(x).mean()
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per step: {tot:.3f} ms")
for k, ms, n in rows[:30]:
    print(f"{ms:9.3f} ms  x{n:<3d} {100 * ms / tot:5.1f}%  {k[:110]}")
