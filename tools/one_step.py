"""One forward+backward step of the bench model (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH_PER_GPU
x = torch.rand(B, 1, bench.CFG["F"], 1, 1, device=dev)
loss = -model(x).mean()
loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
