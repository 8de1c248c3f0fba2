import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from torch.profiler import profile, ProfilerActivity
import ratspn_b200 as rb
DEV = torch.device("cuda", 0)
torch.manual_seed(0); np.random.seed(0)
cfg = rb.CspnConfig()
cfg.F_cond = (376,)
cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 17, 3, 4, 3, 3, 1
cfg.dropout = 0.0; cfg.leaf_base_class = rb.RatNormal; cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
cfg.tanh_squash = True
cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [64], [64], [64]
cfg.cond_layers_inner_act = torch.nn.LeakyReLU
m = rb.CSPN(cfg).to(DEV)
B = 256
obs = torch.randn(B, 376, device=DEV)
failed = torch.rand(B, 17, device=DEV) < 0.05
ev = torch.where(failed, torch.atanh(torch.rand(B, 17, device=DEV) * 1.8 - 0.9), torch.full((B, 17), float("nan"), device=DEV)).view(1, B, 17, 1, 1)
def phases():
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.zero_grad(set_to_none=True)
    ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
    torch.cuda.synchronize(); t["sample_onehot"] = time.perf_counter() - t0; t0 = time.perf_counter()
    H, _ = m.recursive_entropy_approx(condition=obs, sample_size=5, marginal_mask=failed)
    torch.cuda.synchronize(); t["entropy_fwd"] = time.perf_counter() - t0; t0 = time.perf_counter()
    (-(0.1 * H + ctx.sample.reshape(B, 17).sum(-1)).mean()).backward()
    torch.cuda.synchronize(); t["backward"] = time.perf_counter() - t0
    return t
for _ in range(3): phases()
print({k: round(v * 1e3, 2) for k, v in phases().items()}, "ms (wall, synchronised)")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    phases(); torch.cuda.synchronize()
ev_ = prof.key_averages()
nk = sum(e.count for e in ev_ if e.device_type.name == "CUDA")
dt = sum(e.device_time_total for e in ev_ if e.device_type.name == "CUDA") / 1e3
print("kernel launches", nk, "device time ms", round(dt, 3))
rows = sorted(((e.key, e.device_time_total / 1e3, e.count) for e in ev_ if e.device_type.name == "CUDA"), key=lambda r: -r[1])
for k, ms, n in rows[:14]: print(f"{ms:8.3f} ms x{n:<4d} {k[:90]}")
cpu = sorted(((e.key, e.self_cpu_time_total / 1e3, e.count) for e in ev_ if e.device_type.name == "CPU"), key=lambda r: -r[1])
print("host-side ops by self CPU time:")
for k, ms, n in cpu[:22]: print(f"{ms:8.3f} ms x{n:<4d} {k[:90]}")
