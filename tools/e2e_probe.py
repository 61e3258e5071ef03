"""Where does the end-to-end step time go?  Same pipeline as bench.py's e2e leg, one switch at a time.
    python tools/e2e_probe.py [frames]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pose_splatter_b200 import batched, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
sets = [synth.make_views("c2", F, 6, seed=k) for k in range(3)]
W, H = sets[0]["width"], sets[0]["height"]
V = len(sets[0]["view_frame"])
host = [{k: s[k].pin_memory() for k in ("params", "view_frame", "viewmats", "Ks")} for s in sets]
bg = torch.ones(3, device=dev)
w_rgb, w_a = synth.cotangents(V, H, W, seed=7)
w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)
out_host = [torch.empty_like(host[0]["params"]).pin_memory() for _ in range(2)]
h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


loss_host = torch.empty(2).pin_memory()
gnorm_host = [torch.empty(F).pin_memory() for _ in range(2)]


def run(steps, copy_in, copy_out, kernels, extras=0):
    main = torch.cuda.current_stream(dev)
    resident = {k: v.to(dev) for k, v in host[0].items()}
    g_res = torch.zeros_like(resident["params"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nxt = None
    for k in range(steps):
        if copy_in:
            if nxt is None:
                with torch.cuda.stream(h2d):
                    nxt = ({n: v.to(dev, non_blocking=True) for n, v in host[k % 3].items()}, torch.cuda.Event())
                    nxt[1].record(h2d)
            t, ev = nxt
            with torch.cuda.stream(h2d):
                nxt = ({n: v.to(dev, non_blocking=True) for n, v in host[(k + 1) % 3].items()}, torch.cuda.Event())
                nxt[1].record(h2d)
            main.wait_event(ev)
            for v in t.values():
                v.record_stream(main)
        else:
            t = resident
        if kernels:
            rgb, alpha, g = batched.render_views_vjp("3d", t["params"], t["view_frame"], W, H, bg, w_rgb, w_a, t["viewmats"], t["Ks"])
        else:
            g = g_res
        if extras >= 1 and kernels:
            loss = torch.dot(rgb.reshape(-1), w_rgb.reshape(-1)) + torch.dot(alpha.reshape(-1), w_a.reshape(-1))
            lossd = loss.detach().reshape(1)
        if extras >= 2 and kernels:
            gnorm = torch.linalg.vector_norm(g.reshape(g.shape[0], -1), dim=1)
        if copy_out:
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                out_host[k % 2].copy_(g, non_blocking=True)
                if extras >= 2 and kernels:
                    gnorm_host[k % 2].copy_(gnorm, non_blocking=True)
                if extras >= 1 and kernels:
                    loss_host[k % 2:k % 2 + 1].copy_(lossd, non_blocking=True)
            g.record_stream(d2h)
            if extras >= 2 and kernels:
                gnorm.record_stream(d2h)
            if extras >= 1 and kernels:
                lossd.record_stream(d2h)
    main.wait_stream(d2h)
    main.wait_stream(h2d)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name, cfg in (("kernels only", (False, False, True)), ("h2d only", (True, False, False)), ("d2h only", (False, True, False)),
                  ("h2d + d2h", (True, True, False)), ("h2d + kernels", (True, False, True)), ("kernels + d2h", (False, True, True)),
                  ("h2d + kernels + d2h", (True, True, True))):
    run(3, *cfg)
    print(f"{name:24s} {run(10, *cfg):8.3f} ms/step", flush=True)
for extras in (1, 2):
    run(3, True, True, True, extras)
    print(f"full + extras {extras}          {run(10, True, True, True, extras):8.3f} ms/step", flush=True)
