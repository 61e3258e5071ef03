"""ctypes binding of libpsplat.so (include/psplat.h).  PyTorch supplies device memory and the
stream; every computation happens inside the library.  There is no fallback: if the library is
missing, or no CUDA device is present, this module raises."""
from __future__ import annotations

import ctypes
import threading
from pathlib import Path

import torch

MODE_2D, MODE_3D = 2, 3
FLAG_SAVE_FOR_BACKWARD, FLAG_KEEP_BINNING = 1, 2
TAPS = dict(isect_keys=1, flatten_ids=2, tile_offsets=3, last_ids=4, tiles_touched=5, rec0=6, rec1=7, rec2=8, depth=9)

LIB_PATH = Path(__file__).resolve().parent / "libpsplat.so"

EXPORTS = ("ps_abi_version", "ps_last_error", "ps_ctx_create", "ps_ctx_destroy", "ps_forward", "ps_forward_rgba8", "ps_backward", "ps_backward_peer", "ps_peer_sum",
           "ps_saved_info_get", "ps_saved_copy", "ps_saved_release", "ps_ctx_launch_count", "ps_math_probe",
           "ps_ctx_set_profiling", "ps_ctx_stage_times", "ps_ctx_raster_stats", "ps_fp32_peak_probe", "ps_view_loss",
           "ps_param_head_forward", "ps_param_head_backward", "ps_adapter3d_probe", "ps_iou_loss")

STAGES = ("project", "rank", "scan", "partition", "sort", "raster_fwd", "raster_bwd", "project_bwd", "blocks")
FLAG_RASTER_STATS = 4
FLAG_ACTIVATED_INPUTS = 8


class RenderDesc(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("n_frames", ctypes.c_int32), ("n_gauss", ctypes.c_int32), ("n_views", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("near_plane", ctypes.c_float), ("far_plane", ctypes.c_float),
                ("radius_clip", ctypes.c_float), ("eps2d", ctypes.c_float)]


class SavedInfo(ctypes.Structure):
    _fields_ = [("n_isect", ctypes.c_int64), ("tile_bits", ctypes.c_int32), ("view_bits", ctypes.c_int32),
                ("tiles_x", ctypes.c_int32), ("tiles_y", ctypes.c_int32), ("n_views", ctypes.c_int32),
                ("n_gauss", ctypes.c_int32), ("n_frames", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("n_lists", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


_lib = None
_ctx = {}
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """Load libpsplat.so (built ahead of time by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C {LIB_PATH.parent / 'csrc'}` "
            "(pose_splatter_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, ip = ctypes.c_void_p, ctypes.c_int
    lib.ps_abi_version.restype = ip
    lib.ps_last_error.restype = ctypes.c_char_p
    lib.ps_ctx_create.argtypes = [ip, ctypes.POINTER(vp)]
    lib.ps_ctx_destroy.argtypes = [vp]
    lib.ps_forward.argtypes = [vp, ctypes.POINTER(RenderDesc), vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(vp), vp]
    lib.ps_forward_rgba8.argtypes = [vp, ctypes.POINTER(RenderDesc), vp, vp, vp, vp, vp, vp, vp]
    lib.ps_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ps_backward_peer.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ip, ip, vp]
    lib.ps_peer_sum.argtypes = [vp, vp, ip, ctypes.c_size_t, vp, vp]
    lib.ps_saved_info_get.argtypes = [vp, ctypes.POINTER(SavedInfo)]
    lib.ps_saved_copy.argtypes = [vp, vp, ip, vp, ctypes.c_size_t, vp]
    lib.ps_saved_release.argtypes = [vp, vp, vp]
    lib.ps_ctx_launch_count.argtypes = [vp]
    lib.ps_ctx_launch_count.restype = ctypes.c_int64
    lib.ps_math_probe.argtypes = [vp, vp, ip, vp, vp]
    lib.ps_adapter3d_probe.argtypes = [vp, vp, ip, vp, vp, vp, vp]
    lib.ps_ctx_set_profiling.argtypes = [vp, ip]
    lib.ps_ctx_stage_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ip]
    lib.ps_ctx_raster_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), ip, vp]
    lib.ps_fp32_peak_probe.argtypes = [vp, ctypes.POINTER(ctypes.c_double), vp]
    cf, cd = ctypes.c_float, ctypes.c_double
    lib.ps_param_head_forward.argtypes = [vp, ip, ip, vp, vp, vp, vp, cf, cf, cf, cf, ip, cd, ctypes.POINTER(cf), vp, vp, ip, vp, vp]
    lib.ps_param_head_backward.argtypes = [vp, ip, ip, vp, vp, cf, cf, cf, cf, ip, cd, vp, vp, ip, vp, vp, vp, vp, vp]
    lib.ps_iou_loss.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, vp]
    lib.ps_view_loss.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().ps_last_error().decode(errors="replace")
        raise RuntimeError(f"libpsplat {what} failed (code {rc}): {msg}")


def context(device: torch.device) -> ctypes.c_void_p:
    """One ps_ctx per CUDA device, created on first use."""
    if device.type != "cuda":
        raise RuntimeError(f"pose_splatter_b200 renders on CUDA devices only (got {device}); there is no CPU path")
    index = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        if index not in _ctx:
            handle = ctypes.c_void_p()
            check(load().ps_ctx_create(index, ctypes.byref(handle)), "ps_ctx_create")
            _ctx[index] = handle
        return _ctx[index]


def launch_count(device: torch.device) -> int:
    return int(load().ps_ctx_launch_count(context(device)))


def set_profiling(device: torch.device, on: bool):
    check(load().ps_ctx_set_profiling(context(device), int(on)), "ps_ctx_set_profiling")


def stage_times(device: torch.device, reset: bool = True):
    """{stage: (total ms, calls)} measured with CUDA events on the launching stream."""
    ms = (ctypes.c_double * len(STAGES))()
    calls = (ctypes.c_int64 * len(STAGES))()
    check(load().ps_ctx_stage_times(context(device), ms, calls, int(reset)), "ps_ctx_stage_times")
    return {name: (float(ms[i]), int(calls[i])) for i, name in enumerate(STAGES)}


def raster_stats(device: torch.device, reset: bool = True):
    """Pair counters of the rasterizer kernels that ran with FLAG_RASTER_STATS: {"fwd": {...}, "bwd": {...}}."""
    out = (ctypes.c_uint64 * 8)()
    check(load().ps_ctx_raster_stats(context(device), out, int(reset), stream_ptr(device)), "ps_ctx_raster_stats")
    names = ("pairs_evaluated", "pairs_contributing", "entries_walked", "entries_staged")
    return {"fwd": {n: int(out[i]) for i, n in enumerate(names)}, "bwd": {n: int(out[4 + i]) for i, n in enumerate(names)}}


def fp32_peak_tflops(device: torch.device) -> float:
    out = ctypes.c_double()
    check(load().ps_fp32_peak_probe(context(device), ctypes.byref(out), stream_ptr(device)), "ps_fp32_peak_probe")
    return float(out.value)


def stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None
