"""Single-view latency of the drop-in class (the reference's own call shape: one render() + backward per view):
wall clock, device time, and where the host time goes.  `python tools/latency_probe.py [3d|2d]`"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pose_splatter_b200 import create_renderer, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "3d"
wl = "c2" if mode == "3d" else "c3"
dev = torch.device("cuda", 0)
d = synth.make_views(wl, 1, 1, seed=1)
W, H = d["width"], d["height"]
r = create_renderer(mode, W, H, device="cuda")
r.set_background_color(torch.ones(3))
p = d["params"][0].to(dev).requires_grad_(True)
vm, K = d["viewmats"][0].to(dev), d["Ks"][0].to(dev)


def step(backward=True):
    rgb, a = r.render(p, vm, K)
    if backward:
        (rgb.sum() + a.sum()).backward()
        p.grad = None


for bw in (False, True):
    for _ in range(10):
        step(bw)
    torch.cuda.synchronize()
    lat = []
    for _ in range(50):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(bw)
        torch.cuda.synchronize()
        lat.append(1e3 * (time.perf_counter() - t0))
    lat.sort()
    # throughput of back-to-back calls (no sync in between): the host-side cost per call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        step(bw)
    t_enq = 1e3 * (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    t_all = 1e3 * (time.perf_counter() - t0) / 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        step(bw)
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode} {'fwd+bwd' if bw else 'fwd only'}: latency median {lat[len(lat) // 2]:.3f} ms min {lat[0]:.3f} | back-to-back "
          f"host enqueue {t_enq:.3f} ms/call, wall {t_all:.3f} ms/call, device span {e0.elapsed_time(e1) / 200:.3f} ms/call")

pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step(True)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
