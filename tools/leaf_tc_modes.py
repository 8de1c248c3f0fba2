"""Experiment (needs the experiment build: SPN_LEAF_TC_PROF=1 SPN_LIB_OUT=build/variants/libspn_prof.so python
conditional-ratspn-for-rl_b200/build.py, then SPN_LIB_PATH=build/variants/libspn_prof.so): leaf_tc kernels with parts switched
off (spn_debug_leaf_tc_mode bits: 1 no MMA, 2 no x reads, 4 no A-operand writes, 8 no bulk store (forward), 16 no B-operand
build (backward)); results are garbage, only the times matter."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratspn_b200 import ops, _abi  # noqa: E402
from ratspn_b200._abi import call, ptr, stream_ptr  # noqa: E402

dev = "cuda"
F, I, R, B, card = 784, 40, 40, 4096, 25
d0 = -(-F // card)
g = torch.Generator().manual_seed(0)
x = torch.rand(B, 1, F, 1, 1, generator=g).to(dev)
perm = torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1).to(torch.int32).to(dev)
mu = torch.randn(1, F, R, I, generator=g).to(dev)
sigma = torch.sigmoid(torch.randn(1, F, R, I, generator=g)).to(dev)
go = torch.randn(d0, R, B, I, device=dev) / B
lib = _abi.lib()
xT = torch.empty(lib.spn_leaf_workspace_floats(B, F), device=dev)
st = stream_ptr(x.device)
call("spn_leaf_transpose", ptr(x), F, 1, B, F, ptr(xT), st)
img = torch.empty(lib.spn_leaf_tc_image_bytes(I, R, card, d0), device=dev, dtype=torch.uint8)
out = torch.empty(d0, R, B, I, device=dev)
dmu, dsg = torch.empty_like(mu), torch.empty_like(sigma)
sp = (R * I, 1, I)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


modes = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 1 | 2 | 4, 31]
for mode in modes:
    lib.spn_debug_leaf_tc_mode(mode)
    tf = timed(lambda: call("spn_leaf_tc_fwd", ptr(xT), ptr(perm), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0, ptr(img), ptr(out), st))
    tb = timed(lambda: call("spn_leaf_tc_bwd_params", ptr(xT), ptr(go), ptr(perm), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0,
                            ptr(dmu), ptr(dsg), st))
    print(f"mode {mode:2d}: fwd (image + kernel) {tf:.3f} ms   bwd {tb:.3f} ms", flush=True)
lib.spn_debug_leaf_tc_mode(0)
