"""Small driver for profiling the parameter-head kernels alone: python tools/head_probe.py [frames N]"""
import sys
import torch
sys.path.insert(0, ".")
from pose_splatter_b200 import param_head

F, N = (int(x) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (256, 16000)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)
n = F * N
net = torch.randn(n, 14, generator=g).to(dev).requires_grad_(True)
probs = (torch.rand(n, generator=g) * 0.7 + 0.27).to(dev)
grid = (torch.rand(n, 3, generator=g) * 0.2 - 0.1).to(dev)
cot = torch.randn(n, 14, generator=g).to(dev)
angles = torch.rand(F, generator=g, dtype=torch.float64) * 6.28 - 3.14
p3 = torch.rand(F, 3, generator=g) * 0.1 - 0.05
rf = torch.arange(F).repeat_interleave(N).int().to(dev)
scale = torch.tensor([-5.5], device=dev)
for _ in range(3):
    net.grad = None
    rows = param_head.gaussian_rows("3d", net, probs, scale, 0.18 / 112, 0.25, grid_sel=grid, angle=angles, p_3d=p3, row_frame=rf)
    rows.backward(cot)
torch.cuda.synchronize()
print("ok", tuple(rows.shape))
