"""Debug aid: tensor-core leaf (leaf_tc.cu) vs the exact fp32 kernels on one shape, with an error breakdown per axis."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratspn_b200 import ops  # noqa: E402

F, I, R, B, card = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (784, 40, 4, 300, 25))]
d0 = -(-F // card)
dev = "cuda"
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 1, F, 1, 1, generator=g).to(dev)
perm = torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1).to(torch.int32).to(dev)
mu0 = torch.randn(1, F, I, R, generator=g).to(dev)
sg0 = (torch.rand(1, F, I, R, generator=g) * 0.9 + 0.05).to(dev)
go = torch.randn(d0, R, B, I, generator=g).to(dev)


def breakdown(name, a, b, axes):
    err = (a - b).abs()
    print(f"{name}: max err {err.max().item():.3e}  max |ref| {b.abs().max().item():.3e}  bad(>1e-3*max) {(err > 1e-3 * b.abs().max()).float().mean().item():.4f}"
          f"  nan {torch.isnan(a).sum().item()}")
    for i, ax in enumerate(axes):
        dims = [j for j in range(err.dim()) if j != i]
        v = err.amax(dim=dims)
        print(f"   by {ax}: ", " ".join(f"{t:.1e}" for t in v.flatten()[:48].tolist()))


outs = {}
for tc in (False, True):
    ops.LEAF_TENSOR_CORES = tc
    mu, sg = mu0.clone().requires_grad_(True), sg0.clone().requires_grad_(True)
    out = ops.leaf_forward(x, mu, sg, perm, card, False, "drbc")
    torch.cuda.synchronize()
    outs[tc] = [out.detach().clone(), None, None, mu, sg, out]
    print("forward done tc =", tc, flush=True)
breakdown("forward", outs[True][0], outs[False][0], ["scope", "rep", "row", "chan"])
sys.stdout.flush()
for tc in (False, True):
    ops.LEAF_TENSOR_CORES = tc
    _, _, _, mu, sg, out = outs[tc]
    out.backward(go)
    torch.cuda.synchronize()
    outs[tc][1], outs[tc][2] = mu.grad.clone(), sg.grad.clone()
    print("backward done tc =", tc, flush=True)
breakdown("dmu", outs[True][1][0], outs[False][1][0], ["feature", "chan", "rep"])
breakdown("dsigma", outs[True][2][0], outs[False][2][0], ["feature", "chan", "rep"])
