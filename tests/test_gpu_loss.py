"""GPU parity of the fused per-view loss (ps_view_loss) against the fp64 oracle restatement of
scripts/training/train_script.py:30-36,129-133 (oracle/loss_ref.py), through the C ABI.

Tolerances: loss terms 2e-5 absolute (fp32 window sums against fp64); gradients 1e-3 of the tensor's max-abs
gradient (the north_star bar for gradients) plus 2e-7 absolute (fp32 rounding where the exact gradient is 0)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOSS_TOL, GRAD_TOL = 2e-5, 1e-3


def _inputs(seed, V, H, W, flat=True):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(V, H, W, 3, generator=g)
    alpha = torch.rand(V, H, W, generator=g)
    timg = torch.rand(V, 3, H, W, generator=g)
    if flat:  # white background regions like a real render: zero variance, the clamp decides
        rgb[:, : H // 3] = 1.0
        timg[:, :, : H // 4, : W // 2] = 1.0
        alpha[:, : H // 3] = 0.0
    mask = (torch.rand(V, H, W, generator=g) > 0.5).float()
    return rgb, alpha, timg, mask


def _check(rgb, alpha, timg, mask, sl, il):
    from oracle import loss_ref
    from pose_splatter_b200 import losses
    r = rgb.to(DEV).requires_grad_(True)
    a = alpha.to(DEV).requires_grad_(True)
    total, parts = losses.view_loss(r, a, timg.to(DEV), mask.to(DEV), sl, il)
    total.sum().backward()
    want, g_rgb, g_alpha = loss_ref.views_loss_and_grads(rgb, alpha, timg, mask, sl, il)
    assert np.abs(parts.cpu().numpy() - want.numpy()).max() <= LOSS_TOL, (parts.cpu(), want)
    assert np.abs(total.detach().cpu().numpy() - want.sum(1).numpy()).max() <= 3 * LOSS_TOL
    for got, ref, name in ((r.grad, g_rgb, "d_rgb"), (a.grad, g_alpha, "d_alpha")):
        err = (got.cpu().double() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= GRAD_TOL * scale + 2e-7, f"{name}: max err {err:.3e} vs max-abs {scale:.3e}"


@pytest.mark.parametrize("V,H,W", [(1, 11, 11), (2, 23, 31), (3, 64, 40), (1, 100, 145), (6, 256, 288), (2, 141, 403), (1, 300, 200)])
def test_view_loss_matches_oracle(V, H, W):
    _check(*_inputs(10 + H, V, H, W), 0.8, 0.35)


def test_view_loss_reference_lambdas_and_exact_targets():
    """lambdas of the reference configs; a target equal to the render: ssim term 0, L1 sign term 0"""
    rgb, alpha, timg, mask = _inputs(3, 2, 48, 64)
    timg = rgb.permute(0, 3, 1, 2).contiguous()
    _check(rgb, alpha, timg, mask, 1.0, 0.5)


def test_losses_only_path_and_scaling_of_the_upstream_gradient():
    from pose_splatter_b200 import losses
    rgb, alpha, timg, mask = _inputs(5, 2, 32, 32)
    r, a = rgb.to(DEV), alpha.to(DEV)
    with torch.no_grad():  # validation: scripts/training/train_script.py:39-66
        total0, parts0 = losses.view_loss(r, a, timg.to(DEV), mask.to(DEV), 0.8, 0.35)
    r1, a1 = r.clone().requires_grad_(True), a.clone().requires_grad_(True)
    total1, parts1 = losses.view_loss(r1, a1, timg.to(DEV), mask.to(DEV), 0.8, 0.35)
    assert torch.equal(parts0, parts1)
    w = torch.tensor([2.0, -0.5], device=DEV)
    (total1 * w).sum().backward()
    r2, a2 = r.clone().requires_grad_(True), a.clone().requires_grad_(True)
    losses.view_loss(r2, a2, timg.to(DEV), mask.to(DEV), 0.8, 0.35)[0].sum().backward()
    assert torch.allclose(r1.grad, r2.grad * w[:, None, None, None]) and torch.allclose(a1.grad, a2.grad * w[:, None, None])


def test_get_iou_loss_name_and_errors():
    from pose_splatter_b200 import losses
    g = torch.Generator().manual_seed(0)
    a, m = torch.rand(20, 30, generator=g), (torch.rand(20, 30, generator=g) > 0.5).float()
    got = losses.get_iou_loss(a.to(DEV), m.to(DEV))
    inter, union = (a * m).sum(), (a + m - a * m).sum()
    assert abs(float(got) - float(1 - (inter + 1e-6) / (union + 1e-6))) < 1e-5
    with pytest.raises(ValueError, match="same shape"):
        losses.get_iou_loss(a.to(DEV), m[:10].to(DEV))
    with pytest.raises(RuntimeError, match="11x11"):
        losses.view_loss(torch.rand(1, 8, 8, 3, device=DEV), torch.rand(1, 8, 8, device=DEV),
                         torch.rand(1, 3, 8, 8, device=DEV), torch.ones(1, 8, 8, device=DEV), 1.0, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.view_loss(torch.rand(1, 16, 16, 3), torch.rand(1, 16, 16), torch.rand(1, 3, 16, 16), torch.ones(1, 16, 16), 1.0, 1.0)


def test_training_step_render_loss_backward_chain():
    """render_views -> view_loss -> backward to gaussian_params, against the oracle renderer chained with the
    oracle loss (the whole of scripts/training/train_script.py:107-134 except the network)."""
    from oracle import loss_ref, oracle as ora
    from pose_splatter_b200 import batched, losses, synth
    d = synth.make_views("c2", n_frames=1, n_cams=3, seed=4, n=1200)
    W, H = d["width"], d["height"]
    V = len(d["view_frame"])
    g = torch.Generator().manual_seed(9)
    timg = torch.rand(V, 3, H, W, generator=g)
    mask = (torch.rand(V, H, W, generator=g) > 0.7).float()
    p = d["params"].to(DEV).requires_grad_(True)
    bg = torch.ones(3, device=DEV)
    rgb, alpha = batched.render_views("3d", p, d["view_frame"].to(DEV), W, H, bg, d["viewmats"].to(DEV), d["Ks"].to(DEV))
    total, parts = losses.view_loss(rgb, alpha, timg.to(DEV), mask.to(DEV), 0.8, 0.5)
    total.sum().backward()
    args = ("3d", d["params"].numpy(), d["view_frame"].numpy(), W, H, np.ones(3, np.float32), d["viewmats"].numpy(), d["Ks"].numpy())
    z_rgb, z_a = np.zeros((V, H, W, 3), np.float32), np.zeros((V, H, W), np.float32)
    img = ora.render_views(*args, z_rgb, z_a)
    want, g_rgb, g_alpha = loss_ref.views_loss_and_grads(torch.from_numpy(img["rgb"]), torch.from_numpy(img["alpha"]), timg, mask, 0.8, 0.5)
    back = ora.render_views(*args, g_rgb.float().numpy(), g_alpha.float().numpy())
    assert np.abs(parts.cpu().numpy() - want.numpy()).max() <= LOSS_TOL
    got, ref = p.grad.cpu().numpy().reshape(-1, 14), back["d_params"].reshape(-1, 14)
    rel = np.abs(got - ref).max(0) / np.maximum(np.abs(ref).max(0), 1e-20)
    assert rel.max() <= GRAD_TOL, rel


def test_empty_batch_and_degenerate_masks():
    from pose_splatter_b200 import losses
    total, parts = losses.view_loss(torch.zeros(0, 16, 16, 3, device=DEV), torch.zeros(0, 16, 16, device=DEV),
                                    torch.zeros(0, 3, 16, 16, device=DEV), torch.zeros(0, 16, 16, device=DEV), 1.0, 1.0)
    assert total.shape == (0,) and parts.shape == (0, 3)
    # an all-zero target mask: the reference divides by mask.sum() = 0 (:130) -> inf / nan, not an exception
    rgb, alpha, timg, mask = _inputs(1, 1, 16, 16)
    total, parts = losses.view_loss(rgb.to(DEV), alpha.to(DEV), timg.to(DEV), torch.zeros_like(mask).to(DEV), 1.0, 1.0)
    assert not torch.isfinite(parts[0, 2]) and torch.isfinite(parts[0, :2]).all()


def test_iou_loss_own_path_any_size_and_empty_mask():
    """get_iou_loss has its own two-launch path (ps_iou_loss): images smaller than the SSIM window, an all-zero
    target mask gives the reference's finite value (its eps), gradient = autograd of the reference formula."""
    from pose_splatter_b200 import losses
    g = torch.Generator().manual_seed(3)
    for H, W, empty in ((7, 5, False), (40, 33, True), (64, 48, False)):
        a = torch.rand(3, H, W, generator=g)
        m = torch.zeros(3, H, W) if empty else (torch.rand(3, H, W, generator=g) > 0.6).float()
        ar = a.double().requires_grad_(True)
        inter = (ar * m).sum(dim=(-2, -1))
        union = (ar + m - ar * m).sum(dim=(-2, -1))
        want = (1 - (inter + 1e-6) / (union + 1e-6)).mean()  # scripts/training/train_script.py:30-36
        want.backward()
        ad = a.to(DEV).requires_grad_(True)
        got = losses.get_iou_loss(ad, m.to(DEV))
        got.backward()
        assert torch.isfinite(got) and abs(float(got) - float(want)) < 2e-6
        assert float((ad.grad.cpu().double() - ar.grad).abs().max()) <= 1e-3 * float(ar.grad.abs().max()) + 1e-12
    # a zero-weight term of the fused loss is exactly zero, also for an empty mask
    rgb = torch.rand(1, 16, 16, 3, generator=g).to(DEV)
    total, parts = losses.view_loss(rgb, torch.rand(1, 16, 16, generator=g).to(DEV), torch.rand(1, 3, 16, 16, generator=g).to(DEV),
                                    torch.zeros(1, 16, 16, device=DEV), 0.0, 0.0)
    assert torch.isfinite(parts).all() and float(parts[0, 1]) == 0.0 and float(parts[0, 2]) == 0.0
