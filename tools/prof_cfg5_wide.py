"""Kernel-level time breakdown of cfg5-wide (LL forward / forward + backward), torch profiler (CUPTI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util, torch
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
from torch.profiler import ProfilerActivity, profile
import ratspn_b200 as rb
orig_timed = bc.timed
state = {}
def timed(fn, reps=10, warmup=3):
    fn(); torch.cuda.synchronize()
    if "done" not in state or len(state["done"]) < 2:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        print("---- call", len(state.setdefault("done", [])))
        for e in rows[:14]:
            if e.device_type.name == "CUDA":
                print(f"  {e.device_time_total / 1e3:9.3f} ms x{e.count:3d}  {e.key[:100]}")
        state["done"].append(1)
    return orig_timed(fn, reps=2, warmup=0)
bc.timed = timed
print(bc.cfg5_wide()["forward"])
