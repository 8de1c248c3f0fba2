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
hand = ops.tc_hand_buffer(x, lw)
dx = ops.tc_sum_backward_x(x, out, go, wimg, lw, fhand, hand)
torch.cuda.synchronize()
print("ok")
