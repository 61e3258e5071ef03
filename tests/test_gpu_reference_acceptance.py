"""The reference's own renderer tests (tests/test_gaussian_renderer.py, tests/test_renderer_simple.py)
restated against the drop-in on a CUDA device: same inputs, same assertions, plus the
reference's actual values for those inputs (tests/golden/ref2d_known_answers.npz)."""
import math

import numpy as np
import pytest
import torch

from helpers import GOLDEN, RGB_TOL

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _r2d(w=256, h=256, **kw):
    from src.gaussian_renderer import create_renderer
    return create_renderer("2d", w, h, device=DEV, **kw)


def test_single_red_gaussian_centre_and_corner():  # ref tests/test_gaussian_renderer.py:58-87
    r = _r2d()
    params = torch.tensor([[128.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0]], device=DEV)
    rgb, alpha = r.render(params, None, None)
    assert rgb.shape == (256, 256, 3) and alpha.shape == (256, 256)
    assert rgb[128, 128, 0] > 0.5 and rgb[128, 128, 1] < 0.1 and rgb[128, 128, 2] < 0.1
    assert alpha[128, 128] > 0.5 and alpha[0, 0] < 0.1
    z = np.load(GOLDEN / "ref2d_known_answers.npz")
    assert np.abs(rgb[128].cpu().numpy() - z["single_rgb_row128"]).max() <= RGB_TOL
    assert np.abs(alpha[128].cpu().numpy() - z["single_alpha_row128"]).max() <= RGB_TOL


def test_off_screen_gaussian_contributes_nothing():  # ref :89-105
    r = _r2d()
    params = torch.tensor([[-100.0, -100.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0]], device=DEV)
    _, alpha = r.render(params, None, None)
    assert alpha.max() < 0.01


def test_two_gaussians_keep_their_colours():  # ref :107-125
    r = _r2d()
    params = torch.tensor([[64.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0],
                           [192.0, 128.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 2.0]], device=DEV)
    rgb, alpha = r.render(params, None, None)
    assert rgb[128, 64, 0] > 0.5 and rgb[128, 64, 2] < 0.1
    assert rgb[128, 192, 2] > 0.5 and rgb[128, 192, 0] < 0.1
    z = np.load(GOLDEN / "ref2d_known_answers.npz")
    assert np.abs(rgb[128].cpu().numpy() - z["two_rgb_row128"]).max() <= RGB_TOL


def test_rotation_changes_the_anisotropy():  # ref :127-159
    r = _r2d()
    base = [128.0, 128.0, math.log(5.0), math.log(2.0)]
    p0 = torch.tensor([base + [0.0, 1.0, 1.0, 1.0, 2.0]], device=DEV)
    p90 = torch.tensor([base + [math.pi / 2, 1.0, 1.0, 1.0, 2.0]], device=DEV)
    _, a0 = r.render(p0, None, None)
    _, a90 = r.render(p90, None, None)
    assert a0[128, 138] > a0[138, 128]      # theta = 0: elongated along x
    assert a90[138, 128] > a90[128, 138]    # theta = pi/2: elongated along y


def test_wrong_row_width_raises_value_error():  # ref :161-169, :231-244
    from src.gaussian_renderer import create_renderer
    with pytest.raises(ValueError, match="Expected 9 parameters"):
        _r2d().render(torch.randn(10, 14, device=DEV), None, None)
    with pytest.raises(ValueError, match="Expected 14 parameters"):
        create_renderer("3d", 64, 64, device=DEV).render(torch.randn(10, 9, device=DEV), torch.eye(4, device=DEV),
                                                         torch.eye(3, device=DEV))


def test_empty_parameter_set_renders_background():  # ref :171-183
    r = _r2d(64, 64)
    r.set_background_color(torch.tensor([0.5, 0.5, 0.5]))
    rgb, alpha = r.render(torch.zeros(0, 9, device=DEV), None, None)
    assert torch.allclose(rgb, torch.full_like(rgb, 0.5), atol=1e-5) and torch.allclose(alpha, torch.zeros_like(alpha))


def test_3d_randn_rows_eye_viewmat_shapes():  # ref :207-229
    from src.gaussian_renderer import create_renderer
    r = create_renderer("3d", 256, 256, device=DEV)
    assert r.get_num_params() == 14
    params = torch.randn(100, 14, device=DEV)
    K = torch.tensor([[256.0, 0, 128], [0, 256.0, 128], [0, 0, 1]], device=DEV)
    rgb, alpha = r.render(params, torch.eye(4, device=DEV), K)
    assert rgb.shape == (256, 256, 3) and alpha.shape == (256, 256)
    assert torch.isfinite(rgb).all() and torch.isfinite(alpha).all()
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0


@pytest.mark.parametrize("mode", ["2d", "3d"])
def test_output_shapes_non_square(mode):  # ref :307-334
    from src.gaussian_renderer import create_renderer
    r = create_renderer(mode, 128, 96, device=DEV)
    params = torch.randn(50, r.get_num_params(), device=DEV)
    K = torch.tensor([[100.0, 0, 64], [0, 100.0, 48], [0, 0, 1]], device=DEV)
    rgb, alpha = r.render(params, torch.eye(4, device=DEV), K)
    assert rgb.shape == (96, 128, 3) and alpha.shape == (96, 128)


def test_renderer_is_differentiable_like_the_reference_module():
    """src/model.py:164-168 + train_script.py:134: loss.backward() reaches gaussian_params."""
    r = _r2d(96, 80)
    z = np.load(GOLDEN / "ref2d_random_96x80.npz")
    r.set_background_color(torch.from_numpy(z["bg"]))
    p = torch.from_numpy(z["params"]).to(DEV).requires_grad_(True)
    rgb, alpha = r.render(p, None, None)
    (rgb.mean() + alpha.mean()).backward()
    assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0


def test_state_dict_has_only_background_color_and_moves_with_module():
    r = _r2d(32, 32)
    assert list(r.state_dict().keys()) == ["background_color"]
    assert r.background_color.is_cuda
