"""`gsplat.rendering.rasterization` on top of libpsplat.so.

The reference calls gsplat directly, outside its renderer API, in the legacy `PoseSplatter.splat`
(src/model.py:342-361: C cameras in one call, `radius_clip=2.0`, `absgrad=True`, no `backgrounds`, un-normalised
quaternions, already-activated scales / opacities / colours) and in src/plots.py:41-60,93-112,168-187.  This shim
keeps that call signature and the 3-tuple return `(rgb[C,H,W,3], alpha[C,H,W,1], meta)` and runs the same kernels
as GaussianRenderer3D with PS_FLAG_ACTIVATED_INPUTS (no exp / q/(|q|+1e-8) / clamp / sigmoid; gradients w.r.t. the
values passed in).  Supported: the classic RGB rasterization of 3-channel colours with a pinhole camera -- what the
reference uses; anything else raises NotImplementedError instead of silently doing something different.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from pose_splatter_b200.batched import render_views


def rasterization(means: torch.Tensor, quats: torch.Tensor, scales: torch.Tensor, opacities: torch.Tensor,
                  colors: torch.Tensor, viewmats: torch.Tensor, Ks: torch.Tensor, width: int, height: int,
                  near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
                  sh_degree: Optional[int] = None, packed: bool = False, tile_size: int = 16,
                  backgrounds: Optional[torch.Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
                  absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
                  distributed: bool = False, camera_model: str = "pinhole", **unsupported
                  ) -> Tuple[torch.Tensor, torch.Tensor, Dict]:
    if unsupported:
        raise NotImplementedError(f"gsplat shim: unsupported arguments {sorted(unsupported)}")
    if sh_degree is not None or render_mode != "RGB" or rasterize_mode != "classic" or camera_model != "pinhole" \
            or tile_size != 16 or distributed or sparse_grad:
        raise NotImplementedError("gsplat shim: only classic RGB rasterization with a pinhole camera and 16 px tiles "
                                  "(what pose-splatter calls) is implemented")
    if means.dim() != 2 or means.shape[1] != 3 or colors.shape != means.shape:
        raise NotImplementedError("gsplat shim: means [N,3] and colours [N,3] (shared by all cameras) are required")
    N, C = means.shape[0], viewmats.shape[0]
    if quats.shape != (N, 4) or scales.shape != (N, 3) or opacities.shape != (N,):
        raise ValueError("gsplat shim: expected quats [N,4], scales [N,3], opacities [N]")
    if viewmats.shape != (C, 4, 4) or Ks.shape != (C, 3, 3):
        raise ValueError("gsplat shim: expected viewmats [C,4,4] and Ks [C,3,3]")
    rows = torch.cat([means, scales, quats, colors, opacities[:, None]], dim=1).unsqueeze(0)  # [1,N,14], activated
    view_frame = torch.zeros(C, dtype=torch.int32, device=means.device)
    zero_bg = torch.zeros(3, dtype=torch.float32, device=means.device)
    rgb, alpha = render_views("3d", rows, view_frame, int(width), int(height), zero_bg, viewmats, Ks,
                              near_plane=float(near_plane), far_plane=float(far_plane),
                              radius_clip=float(radius_clip), eps2d=float(eps2d), activated=True)
    alpha = alpha.unsqueeze(-1)
    if backgrounds is not None:  # gsplat: render + (1 - alpha) * background, per camera
        rgb = rgb + (1.0 - alpha) * backgrounds.to(rgb.dtype).reshape(C, 1, 1, 3)
    meta = {"width": int(width), "height": int(height), "tile_size": 16, "n_cameras": C, "backend": "libpsplat"}
    return rgb, alpha, meta
