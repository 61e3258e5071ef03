"""Dense torch restatement of the reference 2D splatter (TEST INFRASTRUCTURE ONLY).

Follows src/gaussian_renderer.py:314-334 (activations, background composite) and :379-425
(per-chunk Gaussian weights, then one Gaussian at a time: contribution = g * (1 - A),
canvas += contribution * colour, A += contribution).  Every Gaussian touches every pixel,
rows are composited in parameter order, pixel centres are integers.  Works in any float
dtype (fp64 gives the tolerance head-room reference; fp32 on all host cores is the
"port" CPU baseline that bench.py --impl reference times).
"""
from __future__ import annotations

import torch


def render_dense(params: torch.Tensor, width: int, height: int, background: torch.Tensor, chunk: int = 5):
    """params [N,9] -> (rgb [H,W,3], alpha [H,W]); differentiable w.r.t. params."""
    if params.shape[1] != 9:
        raise ValueError(f"Expected 9 parameters per Gaussian, got {params.shape[1]}")
    dt, dev = params.dtype, params.device
    ys = torch.arange(height, dtype=dt, device=dev).view(1, height, 1)
    xs = torch.arange(width, dtype=dt, device=dev).view(1, 1, width)
    centre, log_sigma, theta = params[:, 0:2], params[:, 2:4], params[:, 4]
    sigma = log_sigma.exp()
    colour = params[:, 5:8].clamp(0.0, 1.0)
    opacity = params[:, 8].sigmoid()
    image = torch.zeros(height, width, 3, dtype=dt, device=dev)
    cover = torch.zeros(height, width, dtype=dt, device=dev)
    n = params.shape[0]
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        off_x = xs - centre[lo:hi, 0].view(-1, 1, 1)
        off_y = ys - centre[lo:hi, 1].view(-1, 1, 1)
        ct = theta[lo:hi].cos().view(-1, 1, 1)
        st = theta[lo:hi].sin().view(-1, 1, 1)
        along = ct * off_x + st * off_y
        across = -st * off_x + ct * off_y
        two_var = 2 * sigma[lo:hi] ** 2 + 1e-8
        expo = along ** 2 / two_var[:, 0].view(-1, 1, 1) + across ** 2 / two_var[:, 1].view(-1, 1, 1)
        weight = torch.exp(-expo) * opacity[lo:hi].view(-1, 1, 1)
        for k in range(hi - lo):
            add = weight[k] * (1.0 - cover)
            image = image + add.unsqueeze(-1) * colour[lo + k].view(1, 1, 3)
            cover = cover + add
    rgb = image + (1.0 - cover).unsqueeze(-1) * background.to(dt).view(1, 1, 3)
    return rgb, cover


def render_dense_with_grad(params, width, height, background, w_rgb, w_a, chunk: int = 5):
    """Forward + autograd backward of L = sum(w_rgb*rgb) + sum(w_a*alpha). Returns rgb, alpha, dL/dparams."""
    p = params.detach().clone().requires_grad_(True)
    rgb, alpha = render_dense(p, width, height, background, chunk)
    loss = (rgb * w_rgb.to(rgb.dtype)).sum() + (alpha * w_a.to(rgb.dtype)).sum()
    loss.backward()
    return rgb.detach(), alpha.detach(), p.grad.detach()
