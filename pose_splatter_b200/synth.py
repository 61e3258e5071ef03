"""Synthetic cameras and Gaussians for parity tests and bench.py.

No dataset, checkpoint or camera file ships with the reference, so every measured
workload is generated here from seeds: the documented 6-camera rig
(docs/reports/CAMERA_COORDINATE_SYSTEMS.md:55-64 intrinsics, normalised so cameras sit
at distance 1, src/utils.py:93-100) and Gaussians distributed like the model's
parameter head emits them (src/model.py:86,218-223).
"""
from __future__ import annotations

import math

import numpy as np
import torch

RIG_FX, RIG_FY, RIG_CX, RIG_CY = 1632.31, 1639.30, 601.35, 491.22
RIG_W, RIG_H = 1152, 1024

# BASELINE.json configs: name -> (mode, downsample, N)
WORKLOADS = {
    "c1": dict(mode="2d", ds=6, width=192, height=171, n=4096),
    "c2": dict(mode="3d", ds=4, width=288, height=256, n=16000),
    "c3": dict(mode="2d", ds=2, width=576, height=512, n=16000),
    "c4": dict(mode="3d", ds=4, width=288, height=256, n=16000),  # full-sequence inference: 3600 frames x 6 cameras, forward only
    "c5_3d": dict(mode="3d", ds=1, width=1152, height=1024, n=16000),
    "c5_2d": dict(mode="2d", ds=1, width=1152, height=1024, n=16000),
}


def ring_cameras(n_cams: int = 6, ds: float = 4.0, radius: float = 1.0):
    """World->camera [C,4,4] (OpenCV: x right, y down, z forward) and intrinsics [C,3,3]."""
    viewmats = np.zeros((n_cams, 4, 4), np.float64)
    Ks = np.zeros((n_cams, 3, 3), np.float64)
    for c in range(n_cams):
        az = 2.0 * math.pi * c / n_cams
        el = math.radians(20.0 if c % 2 == 0 else 35.0)
        pos = radius * np.array([math.cos(az) * math.cos(el), math.sin(az) * math.cos(el), math.sin(el)])
        fwd = -pos / np.linalg.norm(pos)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])
        viewmats[c, :3, :3] = R
        viewmats[c, :3, 3] = -R @ pos
        viewmats[c, 3, 3] = 1.0
        Ks[c] = np.array([[RIG_FX / ds, 0, RIG_CX / ds], [0, RIG_FY / ds, RIG_CY / ds], [0, 0, 1]])
    return torch.from_numpy(viewmats).float(), torch.from_numpy(Ks).float()


def gaussians_3d(n: int, seed: int, grid: int = 112) -> torch.Tensor:
    """[N,14] rows in the renderer's layout (means, log_scales, quats wxyz, colours, logit opacity)."""
    g = torch.Generator().manual_seed(seed)
    voxel = 0.22 / grid
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True) * torch.rand(n, 1, generator=g) ** (1.0 / 3.0)
    means = d * torch.tensor([0.05, 0.02, 0.015]) + 2 * voxel * torch.tanh(torch.randn(n, 3, generator=g))
    yaw = float(torch.rand(1, generator=g)) * 2 * math.pi
    shift = (torch.rand(3, generator=g) - 0.5) * 0.1 * torch.tensor([1.0, 1.0, 0.2])
    cy, sy = math.cos(yaw), math.sin(yaw)
    rot = torch.tensor([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    means = means @ rot.T + shift
    log_scales = -5.5 + 0.3 * torch.randn(n, 3, generator=g)
    quats = torch.randn(n, 4, generator=g)
    colours = torch.sigmoid(torch.randn(n, 3, generator=g)).clamp(0.0, 0.99)
    op = torch.rand(n, 1, generator=g) * (1 - 2e-6) + 1e-6
    return torch.cat([means, log_scales, quats, colours, torch.logit(op)], 1).float()


def project_to_2d(params3d: torch.Tensor, viewmat: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """[N,14] -> [N,9] for one camera: what the reference's convert_3d_to_2d_params stub
    (src/gaussian_renderer.py:567-590) describes. fp64 pinhole/EWA, 0.3 px^2 blur like the 3D path."""
    p = params3d.double()
    q = p[:, 6:10] / p[:, 6:10].norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(-1, 3, 3)
    M = R * p[:, 3:6].exp()[:, None, :]
    cov = M @ M.transpose(1, 2)
    V = viewmat.double()
    Kd = K.double()
    pc = p[:, 0:3] @ V[:3, :3].T + V[:3, 3]
    covc = V[:3, :3] @ cov @ V[:3, :3].T
    fx, fy, cx, cy = Kd[0, 0], Kd[1, 1], Kd[0, 2], Kd[1, 2]
    rz = 1.0 / pc[:, 2]
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * pc[:, 0] * rz * rz, zero, fy * rz, -fy * pc[:, 1] * rz * rz], 1).view(-1, 2, 3)
    c2 = J @ covc @ J.transpose(1, 2)
    c2[:, 0, 0] += 0.3
    c2[:, 1, 1] += 0.3
    evals, evecs = torch.linalg.eigh(c2)
    theta = torch.atan2(evecs[:, 1, 1], evecs[:, 0, 1])  # direction of the major axis
    sx, sy = evals[:, 1].sqrt(), evals[:, 0].sqrt()
    u = fx * pc[:, 0] * rz + cx
    v = fy * pc[:, 1] * rz + cy
    out = torch.stack([u, v, sx.log(), sy.log(), theta], 1)
    return torch.cat([out, p[:, 10:14]], 1).float()


def make_views(workload: str, n_frames: int, n_cams: int = 6, seed: int = 0, n: int | None = None):
    """Batched inputs for one step: params [F,N,P], view->frame map [V], viewmats [V,4,4], Ks [V,3,3].

    3D: one [N,14] per frame shared by its cameras.  2D: one [N,9] per (frame, camera) view
    (each camera sees a different projection), so F == V there.
    """
    wl = WORKLOADS[workload]
    n = wl["n"] if n is None else n
    vm, Ks = ring_cameras(n_cams, wl["ds"])
    frames3d = [gaussians_3d(n, 1000 * seed + f) for f in range(n_frames)]
    if wl["mode"] == "3d":
        params = torch.stack(frames3d)
        view_frame = torch.arange(n_frames).repeat_interleave(n_cams).int()
    else:
        params = torch.stack([project_to_2d(frames3d[f], vm[c], Ks[c]) for f in range(n_frames) for c in range(n_cams)])
        view_frame = torch.arange(n_frames * n_cams).int()
    return dict(mode=wl["mode"], width=wl["width"], height=wl["height"], params=params.contiguous(),
                view_frame=view_frame, viewmats=vm.repeat(n_frames, 1, 1).contiguous(),
                Ks=Ks.repeat(n_frames, 1, 1).contiguous(), n_cams=n_cams)


def cotangents(n_views: int, height: int, width: int, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n_views, height, width, 3, generator=g) - 0.3,
            torch.rand(n_views, height, width, generator=g) - 0.3)
