"""cfg5-narrow (per-sample weights): recursive entropy approximation, time and per-kernel breakdown (CUPTI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import ratspn_b200 as rb
DEV = torch.device("cuda", 0)
w = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
torch.manual_seed(0); np.random.seed(0)
cfg = rb.CspnConfig()
cfg.F_cond = (4,)
cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 8, 6, 16, 16, 1
cfg.dropout = 0.0; cfg.leaf_base_class = rb.RatNormal; cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
cfg.tanh_squash = False
cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
class _Injected(rb.CSPN):
    device = property(lambda self: DEV)
m = rb.RatSpn.__new__(_Injected); rb.RatSpn.__init__(m, cfg); m.replace_layer_params(); m = m.to(DEV)
params = []
for layer in list(m._inner_layers) + [m.root]:
    if isinstance(layer, rb.Sum):
        shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
        layer.weight_param = torch.log_softmax(torch.randn(shape, device=DEV), dim=2).requires_grad_(grad)
        params.append(layer.weight_param)
    else:
        layer.num_conditionals = w
m._sampling_root.weight_param = torch.zeros(w, 1, 1, 1, 1, device=DEV)
base = m._leaf.base_leaf
base.mean_param = torch.randn(w, 1024, 16, 8, device=DEV).requires_grad_(grad)
base.std_param = (torch.rand(w, 1024, 16, 8, device=DEV) * 0.9 + 0.05).requires_grad_(grad)
params += [base.mean_param, base.std_param]
def run():
    if grad:
        for p in params: p.grad = None
        H, _ = rb.RatSpn.recursive_entropy_approx(m, sample_size=5)
        H.sum().backward()
    else:
        with torch.no_grad():
            H, _ = rb.RatSpn.recursive_entropy_approx(m, sample_size=5)
    return H
for _ in range(2): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); H = run(); b.record(); torch.cuda.synchronize()
print(f"w={w} grad={grad}: recursive entropy {a.elapsed_time(b):.2f} ms; H mean {float(H.mean()):.3f}; peak mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
print("kernels:", sum(r[2] for r in rows), "device ms:", round(sum(r[1] for r in rows), 2))
for k, t, n in rows[:12]:
    print(f"{t:9.3f} ms x{n:<3d} {k[:100]}")
