"""pose_splatter_b200: B200-native (sm_100a) Gaussian-splatting renderer behind pose-splatter's renderer API."""
from .gaussian_renderer import GaussianRenderer, GaussianRenderer2D, GaussianRenderer3D, create_renderer
from .batched import render_views, render_views_rgba8, render_views_vjp

__all__ = ["GaussianRenderer", "GaussianRenderer2D", "GaussianRenderer3D", "create_renderer", "render_views", "render_views_rgba8", "render_views_vjp"]
