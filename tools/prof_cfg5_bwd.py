"""cfg5-narrow (per-sample weights) forward + backward: time and per-kernel breakdown (CUPTI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import ratspn_b200 as rb
DEV = torch.device("cuda", 0)
w = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0); np.random.seed(0)
cfg = rb.CspnConfig()
cfg.F_cond = (4,)
cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 8, 6, 16, 16, 1
cfg.dropout = 0.0; cfg.leaf_base_class = rb.RatNormal; cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
cfg.tanh_squash = False
cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
m = rb.RatSpn.__new__(rb.CSPN); rb.RatSpn.__init__(m, cfg); m.replace_layer_params(); m = m.to(DEV)
params, nbytes = [], 0
for layer in list(m._inner_layers) + [m.root]:
    if isinstance(layer, rb.Sum):
        shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
        layer.weight_param = torch.log_softmax(torch.randn(shape, device=DEV), dim=2).requires_grad_(True)
        params.append(layer.weight_param); nbytes += layer.weight_param.numel() * 4
    else:
        layer.num_conditionals = w
base = m._leaf.base_leaf
base.mean_param = torch.randn(w, 1024, 16, 8, device=DEV).requires_grad_(True)
base.std_param = (torch.rand(w, 1024, 16, 8, device=DEV) * 0.9 + 0.05).requires_grad_(True)
params += [base.mean_param, base.std_param]; nbytes += 2 * base.mean_param.numel() * 4
x = torch.randn(1, w, 1024, 1, 1, device=DEV)
def step():
    for p in params: p.grad = None
    (-rb.RatSpn.forward(m, x).mean()).backward()
for _ in range(2): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"w={w}: fwd+bwd {ms:.2f} ms; parameter bytes {nbytes/1e9:.2f} GB; algorithmic traffic ~3x (read w fwd, read w + write dw bwd) = {3*nbytes/ms/1e6:.0f} GB/s")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
for k, t, n in rows[:12]:
    print(f"{t:9.3f} ms x{n:<3d} {k[:100]}")
