"""Capture the cfg3 policy update (sample + entropy + backward, LL fwd+bwd) into one CUDA graph and time replays."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
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
xs = torch.randn(1, B, 17, 1, 1, device=DEV).clamp(-2, 2)
params = [p for p in m.parameters() if p.requires_grad]
for p in params:
    p.grad = torch.zeros_like(p)

def update():
    for p in params:
        p.grad.zero_()
    ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
    H, _ = m.recursive_entropy_approx(condition=obs, sample_size=5, marginal_mask=failed)
    act = ctx.sample.reshape(B, 17)
    (-(0.1 * H + act.sum(-1)).mean()).backward()
    (-m(xs, condition=obs).mean()).backward()
    return act

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (time.perf_counter() - t0) / reps * 1e3

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): update()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    act = update()
g.replay(); torch.cuda.synchronize()
a0 = act.clone(); g0 = params[0].grad.clone()
g.replay(); torch.cuda.synchronize()
print("replays differ in samples (fresh noise per replay):", bool((a0 != act).any()))
print("graphed ms/update (device events, wall):", timeit(g.replay))
del act
print("eager   ms/update (device events, wall):", timeit(update))
