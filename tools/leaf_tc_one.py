"""One forward + parameter-gradient backward of the cfg 2 leaf layer on the tensor-core kernels, then on the exact fp32
kernels (for ncu: ncu --set full -k regex:leaf python tools/leaf_tc_one.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratspn_b200 import ops  # noqa: E402

dev = "cuda"
F, I, R, B, card = 784, 40, 40, 4096, 25
d0 = -(-F // card)
g = torch.Generator().manual_seed(0)
x = torch.rand(B, 1, F, 1, 1, generator=g).to(dev)
perm = torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1).to(torch.int32).to(dev)
mu = torch.randn(1, F, I, R, generator=g).to(dev).requires_grad_(True)
sigma = torch.sigmoid(torch.randn(1, F, I, R, generator=g)).to(dev).requires_grad_(True)
go = torch.randn(d0, R, B, I, device=dev) / B
for tc in (True, False):
    ops.LEAF_TENSOR_CORES = tc
    out = ops.leaf_forward(x, mu, sigma, perm, card, False, "drbc")
    out.backward(go)
    torch.cuda.synchronize()
