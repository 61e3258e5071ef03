"""Import shim: `from gsplat.rendering import rasterization` (src/model.py:10 of the reference) resolves to the
B200-native renderer.  Real gsplat is not a dependency of this repository; where it is installed, keep this
directory off PYTHONPATH (or ahead of it, to route the legacy call sites through libpsplat.so)."""
import importlib.machinery
import os
import sys
import warnings

from . import rendering  # noqa: F401

__version__ = "1.5.0+psplat"


def _shadowed_install():
    """Path of a real gsplat found further down sys.path (this shim takes precedence over it), or None."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for entry in sys.path:
        if not entry or os.path.abspath(entry) == here:
            continue
        try:
            spec = importlib.machinery.PathFinder.find_spec("gsplat", [entry])
        except Exception:
            spec = None
        if spec is not None and spec.origin and os.path.abspath(os.path.dirname(spec.origin)) != os.path.dirname(os.path.abspath(__file__)):
            return spec.origin
    return None


_other = _shadowed_install()
if _other:
    warnings.warn(f"pose_splatter_b200's gsplat shim shadows an installed gsplat ({_other}): `gsplat.rendering.rasterization` is served "
                  "by libpsplat.so. Put the repository root after site-packages on sys.path to use the installed gsplat instead.",
                  stacklevel=2)
