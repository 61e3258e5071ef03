"""Generate golden fixtures by RUNNING THE REFERENCE 2D class (src/gaussian_renderer.py) on CPU.

Run by hand in the build container, where /root/reference exists:
    python tests/golden/make_golden.py
The .npz files it writes are committed; nothing at test time reads /root/reference.
Each fixture holds the inputs (params, background, cotangent seed) and the reference's
fp32 outputs: rgb, alpha and the autograd gradient of
    L = sum(w_rgb * rgb) + sum(w_a * alpha)        
w.r.t. gaussian_params.  The 3D path has no fixture: its arithmetic (gsplat) is not in the
reference tree and not installable here (parity unpinned).
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from src.gaussian_renderer import create_renderer  # noqa: E402  (the reference, unmodified)

OUT = Path(__file__).resolve().parent


def cotangents(seed, H, W):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(H, W, 3, generator=g) - 0.3, torch.rand(H, W, generator=g) - 0.3


def random_params(seed, N, W, H, sig_lo=0.5, sig_hi=3.5):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, generator=g) * W
    v = torch.rand(N, generator=g) * H
    ls = torch.log(torch.rand(N, 2, generator=g) * (sig_hi - sig_lo) + sig_lo)
    th = torch.rand(N, generator=g) * 2 * np.pi
    col = torch.rand(N, 3, generator=g)
    logit = torch.randn(N, generator=g)
    return torch.cat([u[:, None], v[:, None], ls, th[:, None], col, logit[:, None]], 1)


def adversarial_params(seed, N, W, H):
    """sub-pixel sigma, near-opaque, colours outside [0,1], off-screen means, big blobs."""
    p = random_params(seed, N, W, H)
    g = torch.Generator().manual_seed(seed + 1)
    k = N // 8
    p[0:k, 2:4] = -5.5 + 0.3 * torch.randn(k, 2, generator=g)          # sigma ~ e^-5.5 px
    p[0:k // 2, 0:2] = torch.round(p[0:k // 2, 0:2])                    # exactly on pixel centres
    p[k:2 * k, 8] = float(np.log((1 - 1e-6) / 1e-6))                    # opacity 1 - 1e-6
    p[2 * k:3 * k, 5:8] = torch.rand(k, 3, generator=g) * 1.6 - 0.3     # colours outside [0,1]
    p[3 * k:4 * k, 0] = -40.0 + 20 * torch.randn(k, generator=g)       # off-screen left
    p[4 * k:5 * k, 2:4] = np.log(12.0) + 0.2 * torch.randn(k, 2, generator=g)  # big blobs
    p[5 * k:6 * k, 8] = -12.0                                           # nearly transparent
    p[6 * k:7 * k, 4] = 50.0 * torch.randn(k, generator=g)              # large angles
    return p


def run_reference(params, W, H, bg, seed_w, batch_size):
    r = create_renderer("2d", W, H, device="cpu", sigma_cutoff=3.0, kernel_size=5, batch_size=batch_size)
    r.set_background_color(bg)
    p = params.clone().requires_grad_(True)
    rgb, alpha = r.render(p, None, None)
    w_rgb, w_a = cotangents(seed_w, H, W)
    ((rgb * w_rgb).sum() + (alpha * w_a).sum()).backward()
    return rgb.detach().numpy(), alpha.detach().numpy(), p.grad.numpy()


CASES = {
    # name: (maker, seed, N, W, H, bg, batch_size)
    "ref2d_random_96x80": (random_params, 0, 300, 96, 80, (1.0, 1.0, 1.0), 5),
    "ref2d_adversarial_70x50": (adversarial_params, 3, 256, 70, 50, (0.2, 0.9, 0.5), 7),
    "ref2d_c1_192x171": (random_params, 7, 512, 192, 171, (1.0, 1.0, 1.0), 5),
    "ref2d_dense_small_sigma_33x47": (lambda s, n, w, h: random_params(s, n, w, h, 0.05, 0.6), 11, 200, 33, 47, (0.0, 0.0, 0.0), 1),
}

if __name__ == "__main__":
    torch.set_num_threads(8)
    for name, (maker, seed, N, W, H, bg, bs) in CASES.items():
        params = maker(seed, N, W, H).float()
        bgt = torch.tensor(bg, dtype=torch.float32)
        rgb, alpha, grad = run_reference(params, W, H, bgt, 1000 + seed, bs)
        np.savez_compressed(OUT / f"{name}.npz", params=params.numpy(), bg=bgt.numpy(), W=W, H=H,
                            seed_w=1000 + seed, rgb=rgb.astype(np.float16) if False else rgb,
                            alpha=alpha, grad=grad)
        print(name, "N", N, f"{W}x{H}", "alpha max", float(alpha.max()), "grad absmax", float(np.abs(grad).max()))
    # reference's own known-answer cases (tests/test_gaussian_renderer.py:58-183): record actual values
    r = create_renderer("2d", 256, 256, device="cpu")
    single = torch.tensor([[128.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0]])
    rgb, alpha = r.render(single, None, None)
    two = torch.tensor([[64.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0], [192.0, 128.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 2.0]])
    rgb2, alpha2 = r.render(two, None, None)
    np.savez_compressed(OUT / "ref2d_known_answers.npz", single_rgb_row128=rgb[128].numpy(), single_alpha_row128=alpha[128].numpy(),
                        two_rgb_row128=rgb2[128].numpy(), two_alpha_row128=alpha2[128].numpy())
    print("known answers written")
