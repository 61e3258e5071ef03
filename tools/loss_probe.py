"""Small driver for profiling the loss kernels alone: python tools/loss_probe.py [V H W]"""
import sys
import torch
sys.path.insert(0, ".")
from pose_splatter_b200 import losses

V, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1536, 256, 288)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
rgb = torch.rand(V, H, W, 3, generator=g).to(dev)
alpha = torch.rand(V, H, W, generator=g).to(dev)
timg = torch.rand(V, 3, H, W, generator=g).to(dev)
mask = (torch.rand(V, H, W, generator=g) > 0.5).float().to(dev)
for _ in range(3):
    out = losses._launch(rgb, alpha, timg, mask, 1.0, 0.5, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = losses._launch(rgb, alpha, timg, mask, 1.0, 0.5, True)
e1.record()
torch.cuda.synchronize()
print(f"view_loss V={V} {W}x{H}: {e0.elapsed_time(e1) / 5:.3f} ms per call")
