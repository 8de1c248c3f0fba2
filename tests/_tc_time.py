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
hand = ops.tc_hand_buffer(x, out)
if which in ("all", "bwdx", "bwdw"): timeit("tc bwd_x S_1", lambda: ops.tc_sum_backward_x(x, out, go, wimg, lw, fhand, hand), 2 * fl)
if which in ("all", "bwdw"): timeit("tc bwd_w S_1", lambda: ops.tc_sum_backward_w(fhand, hand, x, out, go, lw), fl)
if which in ("all",): timeit("prepare", lambda: ops.tc_prepare(lw))
print("aborts", ops.tc_aborts())
import os, ctypes
if os.environ.get("SPN_TC_DEBUG"):
    from ratspn_b200 import _abi
    buf = (ctypes.c_ulonglong * 24)()
    _abi.lib().spn_debug_tc_prof(buf, 1)
    ops.tc_sum_forward(x, wimg, lw); torch.cuda.synchronize()
    _abi.lib().spn_debug_tc_prof(buf, 0)
    n = max(1, buf[5])
    names = ["g.wait_vec", "g.wait_empty", "g.gen", "g.wait_st", "g.stages", "g.tiles", "l.prologue", "l.wait_acc", "l.epilogue",
             "-", "-", "-", "m.wait_acc_empty", "m.wait_full", "m.issue", "m.wait_w", "-", "m.tiles"]
    print(" ".join(f"{names[i]}={buf[i]/n:.0f}" for i in range(18) if names[i] != "-"))
