import sys, torch
sys.path.insert(0, '.')
from ratspn_b200 import ops
c, d_in, R, B = 40, 32, 40, 4096
x = (torch.randn(d_in, R, B, c, device='cuda') * 3 - 20)
lw = torch.log_softmax(torch.randn(1, 16, c*c, c, R, device='cuda'), dim=2)
go = torch.randn(16, R, B, c, device='cuda')
wimg = ops.tc_prepare(lw)
fhand = ops.tc_fhand_buffer(x, lw)
out = ops.tc_sum_forward(x, wimg, lw, None, fhand)
def timeit(name, fn, flops=None, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{name}: {ms:.3f} ms" + (f"  {flops/ms/1e9:.1f} TFLOP/s" if flops else ""), flush=True)
fl = 2 * B * 16 * 1600 * 40 * R
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fwd"): timeit("tc fwd S_1", lambda: ops.tc_sum_forward(x, wimg, lw), fl)
if which in ("all", "fwd"): timeit("tc fwd S_1 + hand-off", lambda: ops.tc_sum_forward(x, wimg, lw, None, fhand), fl)
hand = ops.tc_hand_buffer(x, lw)
if which in ("all", "bwdx", "bwdw"): timeit("tc bwd_x S_1", lambda: ops.tc_sum_backward_x(x, out, go, wimg, lw, fhand, hand), 2 * fl)
if which in ("all", "bwdw"): timeit("tc bwd_w S_1", lambda: ops.tc_sum_backward_w(fhand, hand, x, out, go, lw), fl)
if which in ("all",): timeit("prepare", lambda: ops.tc_prepare(lw))
print("aborts", ops.tc_aborts())
if "kern" in sys.argv[1:] or which == "all":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            ops.tc_sum_forward(x, wimg, lw, None, fhand)
            ops.tc_sum_backward_x(x, out, go, wimg, lw, fhand, hand)
            ops.tc_sum_backward_w(fhand, hand, x, out, go, lw)
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_type.name == "CUDA":
            print(f"  {e.device_time_total / e.count / 1e3:8.3f} ms x{e.count:3d}  {e.key[:90]}")
import os, ctypes
if os.environ.get("SPN_TC_PROFILE"):
    from ratspn_b200 import _abi
    buf = (ctypes.c_ulonglong * 12)()
    names = ["er", "wait_eL", "wait_acc(ch>0)", "tmem_ld", "release", "ffma2", "tile_end_rest", "tiles", "wait_acc(ch=0)", "wait_stage", "tile_end_bar1"]
    _abi.lib().spn_debug_tc_prof(buf, 1)
    ops.tc_sum_backward_x(x, out, go, wimg, lw, fhand, hand); torch.cuda.synchronize()
    _abi.lib().spn_debug_tc_prof(buf, 0)
    n = max(1, buf[7])
    print("dx epilogue clk per tile (CTA 0, warp 0):", " ".join(f"{names[i]}={buf[i]/n:.0f}" for i in (0, 1, 8, 2, 3, 4, 9, 5, 10, 6)), f"tiles={buf[7]}")
