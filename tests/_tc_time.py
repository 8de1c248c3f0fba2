import sys, torch, time
sys.path.insert(0, '.')
from ratspn_b200 import ops
c, d_in, R, B = 40, 32, 40, 4096
x = (torch.randn(d_in, R, B, c, device='cuda') * 3 - 20)
lw = torch.log_softmax(torch.randn(1, 16, c*c, c, R, device='cuda'), dim=2)
wimg = ops.tc_prepare(lw)
for _ in range(2): out = ops.tc_sum_forward(x, wimg, lw)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): out = ops.tc_sum_forward(x, wimg, lw)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
fl = 2 * B * 16 * 1600 * 40 * R
print(f"tc fwd S_1: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s  aborts={ops.tc_aborts()}")
a.record(); 
for _ in range(5): wimg = ops.tc_prepare(lw)
b.record(); torch.cuda.synchronize()
print(f"prepare: {a.elapsed_time(b)/5:.3f} ms")
