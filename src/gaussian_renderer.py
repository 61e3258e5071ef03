"""Import shim: `from src.gaussian_renderer import create_renderer` (src/model.py:15 of the
reference) resolves to the B200-native drop-in."""
from pose_splatter_b200.gaussian_renderer import (  # noqa: F401
    GaussianRenderer, GaussianRenderer2D, GaussianRenderer3D, create_renderer)
