"""Golden fixture for the 3D ADAPTER stage, produced by the reference's own code.

gsplat itself is absent from the reference tree, but everything GaussianRenderer3D.render does
before and after calling it is plain torch (src/gaussian_renderer.py:175-211).  This script imports the UNMODIFIED
reference class with `gsplat.rendering.rasterization` replaced by a recorder, and stores

  * the tensors the reference hands to gsplat for seeded (partly adversarial) rows: means, quats, scales, opacities,
    colors, viewmats, Ks, backgrounds + the keyword set (`packed`, width, height),
  * the autograd Jacobian of those activated values w.r.t. the raw [N,14] rows (block diagonal per row: [N,14,14]),
    taken through the reference's own graph,
  * what the adapter returns from gsplat's 3-tuple (rgb[0], alpha[0, ..., 0]).

Run by hand in the build container (needs /root/reference):  python tests/golden/make_golden_adapter3d.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
OUT = Path(__file__).resolve().parent

captured = {}


def recorder(**kw):
    """Stands in for gsplat.rendering.rasterization: records the call, returns tensors that depend linearly on
    every differentiable input so that autograd reaches the adapter."""
    captured.clear()
    captured.update(kw)
    H, W = kw["height"], kw["width"]
    act = torch.cat([kw["means"], kw["scales"], kw["quats"], kw["colors"], kw["opacities"][:, None]], 1)  # [N,14]
    captured["act"] = act
    rgb = torch.zeros(1, H, W, 3) + kw["backgrounds"][:, None, None, :]
    alpha = torch.zeros(1, H, W, 1)
    return rgb, alpha, {}


fake = types.ModuleType("gsplat")
fake_r = types.ModuleType("gsplat.rendering")
fake_r.rasterization = recorder
fake.rendering = fake_r
sys.modules["gsplat"] = fake
sys.modules["gsplat.rendering"] = fake_r
sys.path.insert(0, "/root/reference")
from src.gaussian_renderer import create_renderer  # noqa: E402  (the reference, unmodified)


def rows(seed, n):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n, 14, generator=g)
    p[:, 3:6] = -5.5 + 0.3 * torch.randn(n, 3, generator=g)        # log-scales as the model emits them (src/model.py:86,219)
    p[:, 10:13] = torch.rand(n, 3, generator=g)
    k = n // 8
    p[0:k, 10:13] = torch.rand(k, 3, generator=g) * 1.8 - 0.4      # colours outside [0,1]: clamp and its zero gradient
    p[0, 10:13] = torch.tensor([0.0, 1.0, 0.5])                    # exactly on the clamp bounds
    p[k:2 * k, 13] = torch.tensor([-30.0, -15.0, 15.0, 30.0]).repeat(k)[:k]  # saturated sigmoid
    p[2 * k:3 * k, 6:10] *= 1e-6                                   # tiny quaternions: the + 1e-8 matters
    p[2 * k, 6:10] = 0.0                                           # zero quaternion
    p[3 * k:4 * k, 3:6] = torch.randn(k, 3, generator=g) * 3.0     # wide range of scales
    p[4 * k:5 * k, 6:10] *= 50.0
    return p.float()


if __name__ == "__main__":
    W, H, N = 64, 48, 64
    r = create_renderer("3d", W, H, device="cpu")
    bg = torch.tensor([0.2, 0.7, 0.4])
    r.set_background_color(bg)
    p = rows(5, N)
    viewmat = torch.eye(4)
    viewmat[:3, 3] = torch.tensor([0.1, -0.2, 1.0])
    K = torch.tensor([[100.0, 0.0, 32.0], [0.0, 110.0, 24.0], [0.0, 0.0, 1.0]])
    pr = p.clone().requires_grad_(True)
    rgb, alpha = r.render(pr, viewmat, K)
    act = captured["act"]
    jac = torch.zeros(N, 14, 14)  # jac[i, a, j] = d act[i, a] / d p[i, j]; rows are independent in the adapter
    for a in range(14):
        (g,) = torch.autograd.grad(act[:, a].sum(), pr, retain_graph=True)
        jac[:, a, :] = g
    kw = {k: v for k, v in captured.items() if k != "act"}
    np.savez_compressed(
        OUT / "adapter3d_reference.npz", params=p.numpy(), viewmat=viewmat.numpy(), K=K.numpy(), bg=bg.numpy(), W=W, H=H,
        means=kw["means"].detach().numpy(), quats=kw["quats"].detach().numpy(), scales=kw["scales"].detach().numpy(),
        opacities=kw["opacities"].detach().numpy(), colors=kw["colors"].detach().numpy(),
        viewmats=kw["viewmats"].numpy(), Ks=kw["Ks"].numpy(), backgrounds=kw["backgrounds"].numpy(),
        width=kw["width"], height=kw["height"], packed=kw["packed"],
        keywords=np.array(sorted(kw.keys())), act=act.detach().numpy(), jac=jac.numpy(),
        out_rgb_shape=np.array(rgb.shape), out_alpha_shape=np.array(alpha.shape), out_rgb00=rgb[0, 0].detach().numpy())
    print("keywords passed to gsplat:", sorted(kw.keys()))
    print("act", act.shape, "finite", bool(torch.isfinite(act).all()), "jac finite", bool(torch.isfinite(jac).all()))
    print("returned", tuple(rgb.shape), tuple(alpha.shape))
