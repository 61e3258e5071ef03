"""Full-size golden fixtures from the UNMODIFIED reference 2D class (src/gaussian_renderer.py) on CPU.

Run by hand in the build container, where /root/reference exists (about 3 minutes on 8 cores):
    python tests/golden/make_golden_fullsize.py
Nothing at test time reads /root/reference; the .npz files are committed.

Round-1 fixtures stop at 192x171 / N=512.  These pin the sizes the metric is quoted on:
  ref2d_c3_workload_576x512_n16000   forward (no_grad) of camera 0 of the c3 bench workload itself
                                     (pose_splatter_b200.synth, BASELINE.json configs[2]); N = 16000
  ref2d_c3_spread_576x512_n16000     forward (no_grad), 16000 Gaussians spread over the whole image (every tile busy,
                                     long lists everywhere); N = 16000
  ref2d_c3_grad_576x512_n1024        forward + autograd backward at 576x512, N = 1024 (= min_n, src/model.py:32; the
                                     reference needs ~N*H*W*40 B with grad)
  ref2d_c5_workload_1152x1024_n4096  forward (no_grad) at the full resolution of BASELINE.json configs[4], first 4096 rows of a
                                     c5 workload view
The renderer is built exactly as the a6000_2d template does (sigma_cutoff=3.0, kernel_size=5, batch_size=5,
configs/templates/a6000_2d.json:50-55).
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference")
from src.gaussian_renderer import create_renderer  # noqa: E402  (the reference, unmodified)

sys.path.insert(1, str(ROOT))
from pose_splatter_b200 import synth  # noqa: E402  (input generator only)
sys.path.insert(2, str(Path(__file__).resolve().parent))
from make_golden import cotangents, random_params  # noqa: E402

OUT = Path(__file__).resolve().parent


def reference(params, W, H, bg, grad_seed=None):
    r = create_renderer("2d", W, H, device="cpu", sigma_cutoff=3.0, kernel_size=5, batch_size=5)
    r.set_background_color(bg)
    if grad_seed is None:
        with torch.no_grad():
            rgb, alpha = r.render(params, None, None)
        return rgb.numpy(), alpha.numpy(), None
    p = params.clone().requires_grad_(True)
    rgb, alpha = r.render(p, None, None)
    w_rgb, w_a = cotangents(grad_seed, H, W)
    ((rgb * w_rgb).sum() + (alpha * w_a).sum()).backward()
    return rgb.detach().numpy(), alpha.detach().numpy(), p.grad.numpy()


def workload_view(wl, seed, cam, n=None):
    d = synth.make_views(wl, 1, 6, seed=seed)
    p = d["params"][cam]
    return (p if n is None else p[:n]).contiguous(), d["width"], d["height"]


if __name__ == "__main__":
    torch.set_num_threads(8)
    white = torch.ones(3)
    jobs = []
    p, W, H = workload_view("c3", 41, 0)
    jobs.append(("ref2d_c3_workload_576x512_n16000", p, W, H, white, None))
    jobs.append(("ref2d_c3_spread_576x512_n16000", random_params(21, 16000, 576, 512).float(), 576, 512,
                 torch.tensor([0.1, 0.3, 0.7]), None))
    jobs.append(("ref2d_c3_grad_576x512_n1024", random_params(23, 1024, 576, 512, 0.5, 6.0).float(), 576, 512, white, 1023))
    p, W, H = workload_view("c5_2d", 43, 1, n=4096)
    jobs.append(("ref2d_c5_workload_1152x1024_n4096", p, W, H, white, None))
    only = set(sys.argv[1:])
    for name, params, W, H, bg, gseed in jobs:
        if only and name not in only:
            continue
        t0 = time.time()
        rgb, alpha, grad = reference(params, W, H, bg, gseed)
        extra = {} if grad is None else dict(grad=grad, seed_w=gseed)
        np.savez_compressed(OUT / f"{name}.npz", params=params.numpy(), bg=bg.numpy(), W=W, H=H, rgb=rgb, alpha=alpha, **extra)
        print(f"{name}: N={params.shape[0]} {W}x{H} alpha max {alpha.max():.4f} covered px {(alpha > 1e-3).sum()} "
              f"{time.time() - t0:.1f} s, {(OUT / (name + '.npz')).stat().st_size / 1e6:.2f} MB", flush=True)
