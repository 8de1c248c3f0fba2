"""Experiment: per-phase clocks of the leaf_tc kernels (build with SPN_LEAF_TC_PROF=1 SPN_LIB_OUT=..., run with SPN_LIB_PATH=...).
Prints, for block 0, the clocks one warp of each role spent per phase, per tile (forward) / per stage (backward)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratspn_b200 import _abi  # noqa: E402
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
buf = (ctypes.c_longlong * 64)()


def fwd():
    call("spn_leaf_tc_fwd", ptr(xT), ptr(perm), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0, ptr(img), ptr(out), st)


def bwd():
    call("spn_leaf_tc_bwd_params", ptr(xT), ptr(go), ptr(perm), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0, ptr(dmu), ptr(dsg), st)


for mode in [int(v) for v in sys.argv[1:]] or [0]:
    lib.spn_debug_leaf_tc_mode(mode)
    for fn in (fwd, bwd):
        fn(); torch.cuda.synchronize()
        lib.spn_debug_leaf_tc_prof(buf, 1)
        fn(); torch.cuda.synchronize()
        lib.spn_debug_leaf_tc_prof(buf, 1)
        v = list(buf)
        items0 = (d0 * R - 1) // 148 + 1
        if fn is fwd:
            tiles = items0 * 32
            print(f"mode {mode} forward, block 0: {tiles} tiles; clocks per tile")
            print("  producer: other %.0f | x tile wait %.0f | ring read %.0f | stage wait %.0f | split+tcgen05.st %.0f | fence+arrive %.0f" % tuple(t / tiles for t in v[0:6]))
            print("  mma:      other %.0f | image wait %.0f | acc wait %.0f | stage wait %.0f | issue+commit %.0f" % tuple(t / tiles for t in v[8:13]))
            print("  epilogue: other %.0f | acc wait %.0f | tcgen05.ld %.0f | release+staging wait %.0f | staging+store %.0f" % tuple(t / tiles for t in v[16:21]))
        else:
            stages = items0 * 64
            print(f"mode {mode} backward, block 0: {stages} stages; clocks per stage")
            print("  producer: other(x loads) %.0f | stage wait %.0f | A operand %.0f | g wait %.0f | B operand %.0f | fence+arrive %.0f" % tuple(t / stages for t in v[24:30]))
            print("  mma:      other %.0f | operand wait %.0f | issue %.0f" % tuple(t / stages for t in v[32:35]))
        sys.stdout.flush()
lib.spn_debug_leaf_tc_mode(0)
