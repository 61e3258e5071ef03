"""CPU checks of the loss oracle (oracle/loss_ref.py) and of the analytic adjoint formulas the
CUDA kernels implement (pose_splatter_b200/csrc/ps_loss.cu), restated here in NumPy fp64 and compared with autograd
of the oracle.  torchmetrics is absent: the SSIM restatement is pinned only by its own invariants (parity unpinned)."""
import numpy as np
import torch

from oracle import loss_ref


def _case(seed, V=2, H=23, W=31):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(V, H, W, 3, generator=g, dtype=torch.float64)
    rgb[:, :6, :9] = 1.0  # a flat (background) corner: variances ~ 0, the clamp decides
    alpha = torch.rand(V, H, W, generator=g, dtype=torch.float64)
    timg = torch.rand(V, 3, H, W, generator=g, dtype=torch.float64)
    timg[:, :, :4, :5] = 1.0
    mask = (torch.rand(V, H, W, generator=g) > 0.6).double()
    return rgb, alpha, timg, mask


def test_gaussian_taps():
    g = loss_ref.gaussian_taps()
    assert g.shape == (11,) and abs(float(g.sum()) - 1.0) < 1e-15
    assert torch.allclose(g, g.flip(0)) and float(g[5]) == float(g.max())
    assert abs(float(g[4] / g[5]) - np.exp(-0.5 / 2.25)) < 1e-15


def test_ssim_invariants():
    rgb, _, timg, _ = _case(0)
    x = timg[0]
    assert abs(float(loss_ref.ssim_valid(x, x)) - 1.0) < 1e-12          # identical images
    a, b = float(loss_ref.ssim_valid(x, rgb[0].permute(2, 0, 1))), float(loss_ref.ssim_valid(rgb[0].permute(2, 0, 1), x))
    assert abs(a - b) < 1e-14 and a < 0.2                                 # symmetric; unrelated noise scores low
    c = torch.full((3, 20, 20), 0.25, dtype=torch.float64)
    d = torch.full((3, 20, 20), 0.75, dtype=torch.float64)
    want = (2 * 0.25 * 0.75 + 1e-4) / (0.25 ** 2 + 0.75 ** 2 + 1e-4)      # constant images: luminance term only
    assert abs(float(loss_ref.ssim_valid(c, d)) - want) < 1e-12


def test_iou_and_l1_terms_match_reference_formulas():
    rgb, alpha, timg, mask = _case(1, V=1)
    iou, ssim, img = loss_ref.view_loss(rgb[0], alpha[0], timg[0], mask[0], 0.0, 0.7)
    a, m = alpha[0], mask[0]
    want_iou = 1 - ((a * m).sum() + 1e-6) / ((a + m - a * m).sum() + 1e-6)   # train_script.py:30-36
    want_img = 0.7 * (timg[0] - rgb[0].permute(2, 0, 1)).abs().sum() / m.sum()  # :130
    assert abs(float(iou - want_iou)) < 1e-14 and abs(float(img - want_img)) < 1e-12 and float(ssim) == 0.0


def _filt(x, g):
    """'valid' correlation of [..., H, W] with the separable window g x g"""
    H, W = x.shape[-2:]
    k = len(g)
    h = sum(g[t] * x[..., :, t:W - k + 1 + t] for t in range(k))
    return sum(g[t] * h[..., t:H - k + 1 + t, :] for t in range(k))


def _filt_full(a, g):
    """transpose of _filt: [.., H-10, W-10] adjoints back onto [.., H, W] (zero-extended, symmetric taps)"""
    pad = len(g) - 1
    z = np.pad(a, [(0, 0)] * (a.ndim - 2) + [(pad, pad), (pad, pad)])
    return _filt(z, g)


def test_kernel_adjoint_formulas_match_autograd():
    """The formulas of ssim_fwd_kernel / loss_bwd_kernel, in NumPy fp64, against autograd of the oracle."""
    ssim_lambda, img_lambda = 0.8, 0.35
    rgb, alpha, timg, mask = _case(2)
    losses, g_rgb, g_alpha = loss_ref.views_loss_and_grads(rgb, alpha, timg, mask, ssim_lambda, img_lambda)
    g = loss_ref.gaussian_taps().numpy()
    c1, c2 = 1e-4, 9e-4
    for v in range(rgb.shape[0]):
        q = rgb[v].permute(2, 0, 1).numpy()
        p = timg[v].numpy()
        a, m = alpha[v].numpy(), mask[v].numpy()
        H, W = a.shape
        mp, mq = _filt(p, g), _filt(q, g)
        vpp, vqq, vpq = _filt(p * p, g) - mp * mp, _filt(q * q, g) - mq * mq, _filt(p * q, g) - mp * mq
        free = vqq > 0
        N1, N2 = 2 * mp * mq + c1, 2 * vpq + c2
        D1, D2 = mp * mp + mq * mq + c1, np.maximum(vpp, 0) + np.maximum(vqq, 0) + c2
        S = N1 * N2 / (D1 * D2)
        count = 3 * (H - 10) * (W - 10)
        coef = -ssim_lambda / count
        dmu = (2 * mp * N2 - 2 * mp * N1) / (D1 * D2) - S * (2 * mq / D1 + np.where(free, -2 * mq, 0.0) / D2)
        A1, A2, A3 = coef * dmu, np.where(free, coef * (-S / D2), 0.0), coef * 2 * N1 / (D1 * D2)
        msum = m.sum()
        d_q = _filt_full(A1, g) + 2 * q * _filt_full(A2, g) + p * _filt_full(A3, g) - img_lambda / msum * np.sign(p - q)
        I, U = (a * m).sum() + 1e-6, (a + m - a * m).sum() + 1e-6
        d_a = -m / U + I * (1 - m) / U ** 2
        assert abs(ssim_lambda * (1 - S.mean()) - float(losses[v, 1])) < 1e-13
        assert np.abs(d_q.transpose(1, 2, 0) - g_rgb[v].numpy()).max() < 1e-12 * max(1.0, np.abs(d_q).max())
        assert np.abs(d_a - g_alpha[v].numpy()).max() < 1e-15
