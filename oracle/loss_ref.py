"""CPU restatement of the per-view training loss that follows render() every step.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else).  The product path is
pose_splatter_b200/csrc/ps_loss.cu behind ps_view_loss (include/psplat.h).

Follows scripts/training/train_script.py of the reference:
  :30-36    get_iou_loss(alpha, target_mask, eps=1e-6)   soft IoU over the last two dims, 1 - mean
  :129      ssim_lambda * (1 - ssim(target_img[None], rgb[None]))   ssim = torchmetrics
            StructuralSimilarityIndexMeasure(data_range=1.0)  (:270) -- preds = target image, target = render
  :130      img_lambda * |target_img - rgb|.sum() / target_mask.sum()
  :133      total = iou + img + ssim, total.backward()

PARITY UNPINNED for the SSIM term: torchmetrics is a third-party dependency (requirements.txt:7,
`torchmetrics>=0.11.0`, i.e. unpinned) that is absent from /root/reference and from this image.  Its published
algorithm (torchmetrics/functional/image/ssim.py, 1.x) is restated here: 11x11 Gaussian window (sigma 1.5, taps
exp(-(d/sigma)^2/2) normalised), k1 = 0.01, k2 = 0.03, c = (k * data_range)^2, reflection padding by 5 followed by a
valid convolution and a crop of the same 5 pixels -- so only windows that lie entirely inside the image count
and the padding never reaches the result --, variances clamped at 0, mean over channels and the (H-10)x(W-10)
window centres.  The IoU and L1 terms are the reference's own torch code and need no third party.
"""
from __future__ import annotations

import torch

K1, K2 = 0.01, 0.03
WIN, SIGMA, PAD = 11, 1.5, 5
IOU_EPS = 1e-6


def gaussian_taps(dtype=torch.float64):
    d = torch.arange((1 - WIN) / 2, (1 + WIN) / 2, 1, dtype=dtype)
    g = torch.exp(-((d / SIGMA) ** 2) / 2)
    return g / g.sum()


def ssim_valid(pred: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """pred, target [3, H, W] -> scalar SSIM (mean over channels and valid window centres)."""
    g = gaussian_taps(pred.dtype)
    k2d = torch.outer(g, g)[None, None].expand(3, 1, WIN, WIN)
    c1, c2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    stack = torch.stack([pred, target, pred * pred, target * target, pred * target])  # [5, 3, H, W]
    out = torch.nn.functional.conv2d(stack, k2d, groups=3)  # valid windows only: [5, 3, H-10, W-10]
    mu_p, mu_t, e_pp, e_tt, e_pt = out
    s_pp = torch.clamp(e_pp - mu_p * mu_p, min=0.0)
    s_tt = torch.clamp(e_tt - mu_t * mu_t, min=0.0)
    s_pt = e_pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_pp + s_tt + c2))
    return m.mean()


def view_loss(rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda, dtype=torch.float64):
    """One view.  rgb [H,W,3] (as render() returns it), alpha [H,W], target_img [3,H,W], target_mask [H,W].
    Returns (iou, ssim, img) loss terms as 0-d tensors of `dtype` (differentiable w.r.t. rgb / alpha)."""
    rgb_c = rgb.to(dtype).permute(2, 0, 1)
    a = alpha.to(dtype)
    t = target_img.to(dtype)
    m = target_mask.to(dtype)
    inter = (a * m).sum()
    union = (a + m - a * m).sum()
    iou = 1 - (inter + IOU_EPS) / (union + IOU_EPS)
    ssim = ssim_lambda * (1.0 - ssim_valid(t, rgb_c))
    img = img_lambda * torch.abs(t - rgb_c).sum() / m.sum()
    return iou, ssim, img


def views_loss_and_grads(rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda, dtype=torch.float64):
    """Batch of V views: losses [V,3] and the gradients of sum_v (iou+ssim+img)_v w.r.t. rgb [V,H,W,3], alpha [V,H,W]."""
    rgb = rgb.detach().to(dtype).requires_grad_(True)
    alpha = alpha.detach().to(dtype).requires_grad_(True)
    rows = []
    total = 0
    for v in range(rgb.shape[0]):
        parts = view_loss(rgb[v], alpha[v], target_img[v], target_mask[v], ssim_lambda, img_lambda, dtype)
        rows.append(torch.stack([p.detach() for p in parts]))
        total = total + parts[0] + parts[1] + parts[2]
    total.backward()
    return torch.stack(rows), rgb.grad, alpha.grad
