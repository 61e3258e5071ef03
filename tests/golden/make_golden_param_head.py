"""Generates tests/golden/param_head_{3d,2d}.npz by running the REFERENCE's own code (src/model.py:177-298,368-421)
on seeded inputs, with autograd gradients for a seeded cotangent.  Run in the build container only
(/root/reference is absent on the GPU box):  python tests/golden/make_golden_param_head.py
src/model.py imports gsplat / torch_scatter / zarr / h5py at module level; none is touched by the functions used
here, so they are stubbed."""
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch

sys.dont_write_bytecode = True
for m in ["torch_scatter", "zarr", "h5py", "matplotlib", "matplotlib.pyplot", "torchmetrics", "torchmetrics.image", "gsplat", "gsplat.rendering"]:
    sys.modules[m] = mock.MagicMock()
sys.path.insert(0, "/root/reference")
import src.model as M  # noqa: E402

OUT = Path(__file__).resolve().parent


def run(mode, seed, n_vox=600, angle=0.83, p_3d=(0.03, -0.02, 0.011)):
    g = torch.Generator().manual_seed(seed)
    P = 14 if mode == "3d" else 9
    vol0 = torch.randn(n_vox, generator=g) * 2.0
    vol0[:7] = 40.0      # probs -> 1: the clamp at 1 - 1e-6 binds
    vol0[7:12] = 0.25 + float(np.log(0.25 / 0.75)) + 1e-7  # probs just above the threshold: the clamp at 1e-6 binds
    feats = torch.randn(7, n_vox, generator=g)
    volume = torch.cat([vol0[None], feats], 0).requires_grad_(True)
    W1 = torch.randn(7 + 1, P, generator=g)
    W1[:, 8:11 if mode == "3d" else 5:8] *= 3.0  # colours reach the 0.99 clip
    W1 = W1.requires_grad_(True)
    scale = torch.nn.Parameter(-5.5 * torch.ones(1))
    grid = torch.rand(n_vox, 3, generator=g) * 0.2 - 0.1
    fake = types.SimpleNamespace(mask_threshold=0.25, prob_threshold=0.25, mask_threshold_delta=0.05, max_n=16000, min_n=16,
                                 gaussian_param_net=lambda x: x @ W1, gaussian_mode=mode, color_clip=(0, 0.99), scale=scale,
                                 grid=grid, voxel_size=0.18 / 64)
    rows = M.PoseSplatter.get_gaussian_params_from_volume_unified(fake, volume)
    pre = rows.detach().clone()
    if mode == "3d":
        rows = M.PoseSplatter.apply_pose_transform_3d(fake, rows, angle, torch.tensor(p_3d))
    cot = torch.randn(rows.shape, generator=g)
    (rows * cot).sum().backward()
    probs = torch.sigmoid(volume[0].detach() - 0.25)
    sel = probs > 0.25
    net_out = (volume.detach()[:, sel].T @ W1.detach())
    # the reference's gradients arrive at W1 and volume; a test maps d_net_out / d_probs of the tail onto them:
    # d_W1 = x^T d_net_out,  d_volume[:, sel] = W1 d_net_out^T (+ d_probs * probs (1 - probs) on row 0)
    x = volume.detach()[:, sel].T
    np.savez(OUT / f"param_head_{mode}.npz", net_out=net_out.numpy(), probs_sel=probs[sel].numpy(), grid_sel=grid[sel].numpy(),
             scale0=np.float32(-5.5), voxel_size=np.float32(0.18 / 64), pt=np.float32(0.25), angle=np.float64(angle),
             p_3d=np.asarray(p_3d, np.float32), rows_pre_pose=pre.numpy(), rows=rows.detach().numpy(), cot=cot.numpy(),
             x=x.numpy(), W1=W1.detach().numpy(), d_W1=W1.grad.numpy(), d_volume=volume.grad.numpy(), sel=sel.numpy(), d_scale=scale.grad.numpy())
    print(mode, "N =", int(sel.sum()), "rows", tuple(rows.shape))


if __name__ == "__main__":
    run("3d", 1)
    run("2d", 2)
