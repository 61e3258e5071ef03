"""CPU-side checks of the drop-in's interface (no kernels run): the parts of the reference's
tests/test_gaussian_renderer.py that do not render (:28-56, :161-169, :256-290) and the
no-fallback rule."""
import pytest
import torch

from src.gaussian_renderer import GaussianRenderer, GaussianRenderer2D, GaussianRenderer3D, create_renderer


def test_abstract_base_cannot_be_instantiated():
    with pytest.raises(TypeError):
        GaussianRenderer(256, 256)

    class Incomplete(GaussianRenderer):
        pass

    with pytest.raises(TypeError):
        Incomplete(256, 256)


def test_2d_attributes_and_parameter_count():
    r = GaussianRenderer2D(256, 200, device="cpu")
    assert (r.get_num_params(), r.width, r.height, r.device) == (9, 256, 200, "cpu")
    assert r.background_color.shape == (3,) and float(r.background_color.abs().sum()) == 0.0
    assert (r.kernel_size, r.sigma_cutoff, r.batch_size) == (5, 3.0, 1)


def test_3d_needs_no_gsplat_and_has_14_params():
    r = GaussianRenderer3D(64, 64, device="cpu")
    assert r.get_num_params() == 14


def test_factory_semantics():
    assert isinstance(create_renderer("2D", 8, 8, device="cpu"), GaussianRenderer2D)
    assert isinstance(create_renderer("3d", 8, 8, device="cpu"), GaussianRenderer3D)
    r = create_renderer("2d", 8, 8, device="cpu", sigma_cutoff=4.0, kernel_size=7, batch_size=5)
    assert (r.sigma_cutoff, r.kernel_size, r.batch_size) == (4.0, 7, 5)
    create_renderer("3d", 8, 8, device="cpu", sigma_cutoff=4.0)  # kwargs dropped for 3D like the reference
    with pytest.raises(ValueError, match="Unknown renderer mode"):
        create_renderer("4d", 8, 8, device="cpu")


def test_width_height_may_be_none_at_construction():
    r = create_renderer("3d", None, None, device="cpu")  # scripts/preprocessing/calculate_visual_features.py:206-215
    assert r.width is None


def test_background_color_validation_and_checkpoint_key():
    r = create_renderer("2d", 8, 8, device="cpu")
    with pytest.raises(ValueError, match="Expected color shape"):
        r.set_background_color(torch.ones(4))
    r.set_background_color(torch.tensor([1.0, 0.5, 0.25]))
    assert list(r.state_dict().keys()) == ["background_color"]
    r2 = create_renderer("2d", 8, 8, device="cpu")
    r2.load_state_dict(r.state_dict(), strict=True)
    assert torch.equal(r2.background_color, r.background_color)


def test_wrong_row_width_raises_before_any_launch():
    with pytest.raises(ValueError, match="Expected 9 parameters"):
        create_renderer("2d", 8, 8, device="cpu").render(torch.zeros(3, 14), None, None)
    with pytest.raises(ValueError, match="Expected 14 parameters"):
        create_renderer("3d", 8, 8, device="cpu").render(torch.zeros(3, 9), torch.eye(4), torch.eye(3))


def test_no_cpu_fallback():
    """A CPU tensor must fail loudly, never be served by another implementation."""
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        create_renderer("2d", 8, 8, device="cpu").render(torch.zeros(3, 9), None, None)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        create_renderer("3d", 8, 8, device="cpu").render(torch.zeros(3, 14), torch.eye(4), torch.eye(3))


def test_product_never_imports_the_oracle():
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parent.parent / "pose_splatter_b200"
    for f in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")) + list(root.rglob("*.h")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<].*oracle", text, re.M), f


def test_new_rows_fail_loudly_without_cuda():
    """losses / param_head have no CPU path (the oracle is test infrastructure, never a fallback)"""
    import pytest
    import torch
    from pose_splatter_b200 import batched, losses, param_head
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.view_loss(torch.rand(1, 16, 16, 3), torch.rand(1, 16, 16), torch.rand(1, 3, 16, 16), torch.ones(1, 16, 16), 1.0, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        param_head.gaussian_rows("3d", torch.zeros(4, 14), torch.zeros(4), torch.zeros(1), 0.1, 0.25, grid_sel=torch.zeros(4, 3))
    with pytest.raises(ValueError, match="Expected 9 parameters"):
        param_head.gaussian_rows("2d", torch.zeros(4, 14), torch.zeros(4), torch.zeros(1), 0.1, 0.25)
    with pytest.raises(ValueError, match="Unknown renderer mode"):
        param_head.gaussian_rows("4d", torch.zeros(4, 14), torch.zeros(4), torch.zeros(1), 0.1, 0.25)
    with pytest.raises(RuntimeError, match="CUDA"):
        batched.render_views_vjp("3d", torch.zeros(1, 4, 14), torch.zeros(1, dtype=torch.int32), 16, 16, torch.ones(3),
                                 torch.zeros(1, 16, 16, 3), torch.zeros(1, 16, 16), torch.eye(4)[None], torch.eye(3)[None])
    with pytest.raises(ValueError, match="same shape"):
        losses.get_iou_loss(torch.rand(4, 4), torch.rand(4, 5))


def test_shims_and_host_modules_never_import_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/"""
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for f in list((root / "pose_splatter_b200").glob("*.py")) + list((root / "gsplat").glob("*.py")) + list((root / "src").glob("*.py")):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f
