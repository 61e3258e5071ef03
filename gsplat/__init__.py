"""Import shim: `from gsplat.rendering import rasterization` (src/model.py:10 of the reference) resolves to the
B200-native renderer.  Real gsplat is not a dependency of this repository; where it is installed, keep this
directory off PYTHONPATH (or ahead of it, to route the legacy call sites through libpsplat.so)."""
from . import rendering  # noqa: F401

__version__ = "1.5.0+psplat"
