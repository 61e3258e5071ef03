"""The `gsplat.rendering.rasterization` shim against the oracle, with the keyword arguments of the
reference's legacy call (src/model.py:342-361: several cameras per call, radius_clip=2.0, absgrad=True,
packed=False, no backgrounds, activated inputs, un-normalised quaternions)."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGB_TOL, column_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(N, seed):
    from pose_splatter_b200 import synth
    p = synth.gaussians_3d(N, seed)
    p[:, 3:6] = torch.exp(p[:, 3:6] + 0.5)
    p[:, 10:13] = p[:, 10:13] * 1.2           # gsplat does not clamp colours
    p[:, 13] = torch.sigmoid(p[:, 13])
    return p


def test_shim_matches_oracle_multi_camera_with_gradients():
    from gsplat.rendering import rasterization
    from oracle import oracle as ora
    from pose_splatter_b200 import synth
    W, H, C = 144, 128, 3
    vm, Ks = synth.ring_cameras(6, ds=8.0)
    vm, Ks = vm[:C], Ks[:C]
    rows = _inputs(1500, 3)
    leaves = [rows[:, a:b].clone().to(DEV).requires_grad_(True) for a, b in ((0, 3), (6, 10), (3, 6), (13, 14), (10, 13))]
    means, quats, scales, opac, cols = leaves
    rgb, alpha, meta = rasterization(means, quats, scales, opac.squeeze(-1), cols, vm.to(DEV), Ks.to(DEV), W, H,
                                     packed=False, near_plane=0.01, far_plane=1e10, render_mode="RGB", sh_degree=None,
                                     sparse_grad=False, absgrad=True, rasterize_mode="classic", radius_clip=2.0)
    assert rgb.shape == (C, H, W, 3) and alpha.shape == (C, H, W, 1) and meta["n_cameras"] == C
    w_rgb, w_a = synth.cotangents(C, H, W, seed=12)
    ((rgb * w_rgb.to(DEV)).sum() + (alpha[..., 0] * w_a.to(DEV)).sum()).backward()
    want = ora.render_views("3d", rows.numpy()[None], np.zeros(C, np.int32), W, H, np.zeros(3, np.float32), vm.numpy(),
                            Ks.numpy(), w_rgb.numpy(), w_a.numpy(), radius_clip=2.0, activated=True)
    assert np.abs(rgb.detach().cpu().numpy() - want["rgb"]).max() <= RGB_TOL
    assert np.abs(alpha[..., 0].detach().cpu().numpy() - want["alpha"]).max() <= RGB_TOL
    got = np.concatenate([means.grad.cpu().numpy(), scales.grad.cpu().numpy(), quats.grad.cpu().numpy(),
                          cols.grad.cpu().numpy(), opac.grad.cpu().numpy()], 1)
    rel = column_rel_err(got, want["d_params"][0])
    assert rel.max() <= GRAD_TOL, rel


def test_shim_backgrounds_and_legacy_composite_agree():
    """`backgrounds=` composites like gsplat; the legacy caller composites itself (src/model.py:363-364)."""
    from gsplat.rendering import rasterization
    from pose_splatter_b200 import synth
    W, H = 96, 80
    vm, Ks = synth.ring_cameras(6, ds=12.0)
    rows = _inputs(600, 8).to(DEV)
    args = (rows[:, 0:3], rows[:, 6:10], rows[:, 3:6], rows[:, 13], rows[:, 10:13], vm[:2].to(DEV), Ks[:2].to(DEV), W, H)
    bg = torch.tensor([[0.2, 0.4, 0.6], [1.0, 1.0, 1.0]], device=DEV)
    rgb0, a0, _ = rasterization(*args, radius_clip=2.0)
    rgb1, a1, _ = rasterization(*args, radius_clip=2.0, backgrounds=bg)
    assert torch.equal(a0, a1)
    assert torch.allclose(rgb1, rgb0 + (1 - a0) * bg.view(2, 1, 1, 3), atol=1e-6)


def test_shim_rejects_what_it_does_not_implement():
    from gsplat.rendering import rasterization
    z = torch.zeros(4, 3, device=DEV)
    with pytest.raises(NotImplementedError):
        rasterization(z, torch.ones(4, 4, device=DEV), z + 1, torch.ones(4, device=DEV), z, torch.eye(4, device=DEV)[None],
                      torch.eye(3, device=DEV)[None], 32, 32, sh_degree=3)
    with pytest.raises(NotImplementedError):
        rasterization(z, torch.ones(4, 4, device=DEV), z + 1, torch.ones(4, device=DEV), z, torch.eye(4, device=DEV)[None],
                      torch.eye(3, device=DEV)[None], 32, 32, render_mode="RGB+D")
