"""Times the leaf layer of BASELINE config 2 (F = 784, D = 5, I = 40, R = 40, 4096 rows) alone: tensor-core kernels
(leaf_tc.cu) against the exact fp32 kernels (leaf_tiled.cu), forward and parameter-gradient backward, CUDA events.
The working set (839 MB of activations) exceeds the L2, so no flush is needed between launches."""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
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
out_bytes = d0 * R * B * I * 4
res = {}
for tc in (True, False):
    ops.LEAF_TENSOR_CORES = tc
    tf, tb = [], []
    for it in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        mu.grad = sigma.grad = None
        e[0].record()
        out = ops.leaf_forward(x, mu, sigma, perm, card, False, "drbc")
        e[1].record()
        out.backward(go)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
    key = "tensor_cores" if tc else "exact_fp32"
    res[key] = {"fwd_ms": sum(tf) / len(tf), "bwd_ms": sum(tb) / len(tb),
                "fwd_write_GBps": out_bytes / (sum(tf) / len(tf)) / 1e6, "bwd_read_GBps": out_bytes / (sum(tb) / len(tb)) / 1e6}
    res[key + "_out"] = out.detach()
    res[key + "_dmu"] = mu.grad.detach().clone()
    res[key + "_dsigma"] = sigma.grad.detach().clone()
dev_out = (res["tensor_cores_out"] - res["exact_fp32_out"]).abs().max().item()
summary = {"shape": {"F": F, "I": I, "R": R, "rows": B, "card": card}, "tensor_cores": res["tensor_cores"],
           "exact_fp32": res["exact_fp32"], "max_abs_dev_out": dev_out,
           "max_abs_out": res["exact_fp32_out"].abs().max().item(),
           "rel_dev_dmu": ((res["tensor_cores_dmu"] - res["exact_fp32_dmu"]).abs().max() / res["exact_fp32_dmu"].abs().max()).item(),
           "rel_dev_dsigma": ((res["tensor_cores_dsigma"] - res["exact_fp32_dsigma"]).abs().max() / res["exact_fp32_dsigma"].abs().max()).item(),
           "note": "fwd includes the x transpose and (tensor cores) the parameter-image kernel; bwd = parameter gradients"}
print(json.dumps(summary))
