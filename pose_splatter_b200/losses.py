"""The per-view training loss of the reference, fused with its backward on the GPU.

Reference: scripts/training/train_script.py:30-36 (`get_iou_loss`), :121-133 (iou + ssim_lambda * (1 - SSIM) +
img_lambda * L1 / sum(mask), then `total_loss.backward()`), SSIM = torchmetrics
`StructuralSimilarityIndexMeasure(data_range=1.0)` (:270).  There the three terms are three autograd graphs over
`[H, W, 3]` tensors plus three `.item()` syncs per step; here one C-ABI call (`ps_view_loss`) produces the loss
terms of V views and -- in the same pass -- the cotangents `d_rgb`, `d_alpha` that the renderer's backward takes.
No CPU or PyTorch fallback: CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _capi

LOSS_NAMES = ("iou", "ssim", "img")  # scripts/training/train_script.py:25


def _launch(rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda, want_grad):
    if rgb.device.type != "cuda":
        raise RuntimeError(f"pose_splatter_b200.losses runs on CUDA tensors only (got {rgb.device}); there is no CPU path")
    V, H, W, C = rgb.shape
    if C != 3 or alpha.shape != (V, H, W):
        raise ValueError(f"Expected rgb [V,H,W,3] and alpha [V,H,W], got {tuple(rgb.shape)} and {tuple(alpha.shape)}")
    if target_img.shape != (V, 3, H, W) or target_mask.shape != (V, H, W):
        raise ValueError("Predicted and target masks must have the same shape.")  # the reference's message (:32)
    dev = rgb.device
    f32 = dict(dtype=torch.float32, device=dev)
    rgb_c, alpha_c = rgb.detach().float().contiguous(), alpha.detach().float().contiguous()
    timg, tmask = target_img.to(**f32).contiguous(), target_mask.to(**f32).contiguous()
    losses = torch.empty((V, 3), **f32)
    d_rgb = torch.empty_like(rgb_c) if want_grad else None
    d_alpha = torch.empty_like(alpha_c) if want_grad else None
    _capi.check(_capi.load().ps_view_loss(_capi.context(dev), V, H, W, _capi.ptr(rgb_c), _capi.ptr(alpha_c),
                                          _capi.ptr(timg), _capi.ptr(tmask), float(ssim_lambda), float(img_lambda),
                                          _capi.ptr(losses), _capi.ptr(d_rgb), _capi.ptr(d_alpha),
                                          _capi.stream_ptr(dev)), "ps_view_loss")
    return losses, d_rgb, d_alpha


class _ViewLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda):
        want = rgb.requires_grad or alpha.requires_grad
        losses, d_rgb, d_alpha = _launch(rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda, want)
        if want:
            ctx.save_for_backward(d_rgb, d_alpha)
        ctx.mark_non_differentiable(losses)
        return losses.sum(1), losses

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        d_rgb, d_alpha = ctx.saved_tensors
        g = g_total.to(d_rgb.dtype)
        return d_rgb * g[:, None, None, None], d_alpha * g[:, None, None], None, None, None, None


def view_loss(rgb, alpha, target_img, target_mask, ssim_lambda: float, img_lambda: float):
    """Loss of V rendered views against their targets.

    rgb [V,H,W,3], alpha [V,H,W] as `render_views` returns them; target_img [V,3,H,W], target_mask [V,H,W] as the
    reference's loader yields them.  Returns (total [V], parts [V,3]): total = iou + ssim + img per view, differentiable
    w.r.t. rgb and alpha (the gradient was computed by the same kernels as the loss); parts = the three terms in the
    order of LOSS_NAMES, for logging.
    """
    return _ViewLoss.apply(rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda)


class _IouLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, mask):
        dev = alpha.device
        V, H, W = alpha.shape
        a = alpha.detach().float().contiguous()
        m = mask.detach().to(dtype=torch.float32, device=dev).contiguous()
        losses = torch.empty(V, dtype=torch.float32, device=dev)
        d_alpha = torch.empty_like(a) if alpha.requires_grad else None
        _capi.check(_capi.load().ps_iou_loss(_capi.context(dev), V, H, W, _capi.ptr(a), _capi.ptr(m), _capi.ptr(losses),
                                             _capi.ptr(d_alpha), _capi.stream_ptr(dev)), "ps_iou_loss")
        if d_alpha is not None:
            ctx.save_for_backward(d_alpha)
        return losses

    @staticmethod
    def backward(ctx, g):
        (d_alpha,) = ctx.saved_tensors
        return d_alpha * g.to(d_alpha.dtype)[:, None, None], None


def get_iou_loss(predicted_mask, target_mask, eps=1e-6):
    """Name and meaning of scripts/training/train_script.py:30-36 for one [H,W] (or [V,H,W]) pair of masks: its own
    two-launch path (ps_iou_loss), any image size, finite for an empty target mask like the reference."""
    if predicted_mask.shape != target_mask.shape:
        raise ValueError("Predicted and target masks must have the same shape.")
    if eps != 1e-6:
        raise ValueError("the fused kernel implements the reference's eps = 1e-6")
    if predicted_mask.device.type != "cuda":
        raise RuntimeError(f"pose_splatter_b200.losses runs on CUDA tensors only (got {predicted_mask.device}); there is no CPU path")
    a = predicted_mask if predicted_mask.dim() == 3 else predicted_mask[None]
    m = target_mask if target_mask.dim() == 3 else target_mask[None]
    return _IouLoss.apply(a, m).mean()
