"""libpsplat.so loads on a CPU-only box and exports every function include/psplat.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "psplat.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from pose_splatter_b200 import _capi
    assert sorted(_capi.EXPORTS) == _declared()


def test_library_exports_every_declared_symbol():
    from pose_splatter_b200 import _capi
    lib = _capi.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.ps_abi_version() == 4


def test_context_creation_fails_loudly_without_gpu():
    import torch
    from pose_splatter_b200 import _capi
    if torch.cuda.is_available():
        return
    lib = _capi.load()
    h = ctypes.c_void_p()
    assert lib.ps_ctx_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.ps_last_error()
