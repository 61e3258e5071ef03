"""Independent fp64 torch restatement of the 3D path with autograd (TEST INFRASTRUCTURE ONLY).

Adapter: src/gaussian_renderer.py:183-211.  Core: gsplat 1.5.x rasterization(packed=False,
classic, RGB) semantics as listed in SURVEY.md 8c-c5 (gsplat itself is absent: PARITY
UNPINNED).  Dense over pixels, Gaussians visited in global depth order with the same
per-pixel skip / stop rules; no tiles.  A Gaussian whose rectangle misses a tile has
alpha < 1/255 there in exact arithmetic, so the dense image equals the tiled one up to
threshold ties.  Used to validate ps_oracle.c (forward values and all gradients).
"""
from __future__ import annotations

import math

import torch


def quat_to_rotmat(q):
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)


def project(params, viewmat, K, width, height, near=0.01, far=1e10, eps2d=0.3, radius_clip=0.0, activated=False):
    means, log_s, quats, cols, logit = params[:, 0:3], params[:, 3:6], params[:, 6:10], params[:, 10:13], params[:, 13]
    if activated:  # values as gsplat.rendering.rasterization receives them (src/model.py:342-361)
        scales, opac = log_s, logit
    else:          # adapter activations, src/gaussian_renderer.py:183-193
        scales = log_s.exp()
        quats = quats / (quats.norm(dim=-1, keepdim=True) + 1e-8)
        cols = cols.clamp(0.0, 1.0)
        opac = logit.sigmoid()
    qh = quats / quats.norm(dim=-1, keepdim=True)
    R = quat_to_rotmat(qh)
    M = R * scales[:, None, :]
    cov = M @ M.transpose(1, 2)
    Rw, t = viewmat[:3, :3], viewmat[:3, 3]
    pc = means @ Rw.T + t
    covc = Rw @ cov @ Rw.T
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    tanx, tany = 0.5 * width / fx, 0.5 * height / fy
    lim_xp, lim_xn = (width - cx) / fx + 0.3 * tanx, cx / fx + 0.3 * tanx
    lim_yp, lim_yn = (height - cy) / fy + 0.3 * tany, cy / fy + 0.3 * tany
    x, y, z = pc.unbind(-1)
    rz = 1.0 / z
    tx = z * torch.minimum(lim_xp, torch.maximum(-lim_xn, x * rz))
    ty = z * torch.minimum(lim_yp, torch.maximum(-lim_yn, y * rz))
    zero = torch.zeros_like(z)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], -1).view(-1, 2, 3)
    cov2 = J @ covc @ J.transpose(1, 2)
    mean2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], -1)
    a = cov2[:, 0, 0] + eps2d
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + eps2d
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], -1)
    with torch.no_grad():
        valid = (z >= near) & (z <= far) & (det > 0) & (opac >= 1.0 / 255.0)
        ext = torch.sqrt(2.0 * torch.log(opac.clamp_min(1e-30) * 255.0).clamp_min(0)).clamp_max(3.33)
        bh = 0.5 * (a + c)
        v1 = bh + torch.sqrt((bh * bh - det).clamp_min(0.01))
        r1 = ext * torch.sqrt(v1)
        rx = torch.ceil(torch.minimum(ext * torch.sqrt(a.clamp_min(0)), r1))
        ry = torch.ceil(torch.minimum(ext * torch.sqrt(c.clamp_min(0)), r1))
        valid &= ~((rx <= radius_clip) & (ry <= radius_clip))
        valid &= ~((mean2d[:, 0] + rx <= 0) | (mean2d[:, 0] - rx >= width) |
                   (mean2d[:, 1] + ry <= 0) | (mean2d[:, 1] - ry >= height))
    return dict(mean2d=mean2d, depth=z, conic=conic, opacity=opac, colour=cols, valid=valid, rx=rx, ry=ry)


def render(params, viewmat, K, width, height, background, **kw):
    """params [N,14] (any float dtype) -> rgb [H,W,3], alpha [H,W]; differentiable."""
    pr = project(params, viewmat, K, width, height, **kw)
    dt = params.dtype
    order = torch.argsort(pr["depth"].detach(), stable=True)
    ys = torch.arange(height, dtype=dt).view(height, 1) + 0.5
    xs = torch.arange(width, dtype=dt).view(1, width) + 0.5
    T = torch.ones(height, width, dtype=dt)
    done = torch.zeros(height, width, dtype=torch.bool)
    img = torch.zeros(height, width, 3, dtype=dt)
    ncon = torch.zeros(height, width, dtype=torch.int32)
    for i in order.tolist():
        if not bool(pr["valid"][i]):
            continue
        # tile rectangle of this Gaussian (16 px tiles): outside it the Gaussian is not listed
        with torch.no_grad():
            mx, my = float(pr["mean2d"][i, 0]), float(pr["mean2d"][i, 1])
            rx, ry = float(pr["rx"][i]), float(pr["ry"][i])
            tw, th = (width + 15) // 16, (height + 15) // 16
            tx0 = min(max(math.floor((mx - rx) / 16), 0), tw); tx1 = min(max(math.ceil((mx + rx) / 16), 0), tw)
            ty0 = min(max(math.floor((my - ry) / 16), 0), th); ty1 = min(max(math.ceil((my + ry) / 16), 0), th)
            listed = torch.zeros(height, width, dtype=torch.bool)
            listed[ty0 * 16:ty1 * 16, tx0 * 16:tx1 * 16] = True
        dx = pr["mean2d"][i, 0] - xs
        dy = pr["mean2d"][i, 1] - ys
        A, B, C = pr["conic"][i]
        sigma = 0.5 * (A * dx * dx + C * dy * dy) + B * dx * dy
        alpha = torch.clamp_max(pr["opacity"][i] * torch.exp(-sigma), 0.999)
        cand = listed & (sigma >= 0) & (alpha >= 1.0 / 255.0) & ~done
        nT = T * (1 - alpha)
        stop = cand & (nT <= 1e-4)
        done = done | stop
        use = cand & ~stop
        vis = torch.where(use, alpha * T, torch.zeros_like(T))
        img = img + vis.unsqueeze(-1) * pr["colour"][i].view(1, 1, 3)
        T = torch.where(use, nT, T)
        ncon += use.to(torch.int32)
    rgb = img + T.unsqueeze(-1) * background.to(dt).view(1, 1, 3)
    return rgb, 1 - T, ncon
