"""Batched multi-view rendering on top of the C ABI: the unit of work the GPU is fed with.

The reference renders one (frame, camera) view per call (scripts/training/train_script.py:107,
src/model.py:164-168); one 288x256 view has ~10 busy tiles and cannot occupy 148 SMs
(SURVEY.md fact 9), so the product API takes V views at once.  `render_views` is differentiable
w.r.t. `params` only (viewmat / K / background are non-trainable in the reference:
src/model.py:82-83, src/gaussian_renderer.py:53).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _capi

DEFAULTS_3D = dict(near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, activated=False)


def _desc(mode, width, height, n_frames, n_gauss, n_views, flags, opts):
    o = dict(DEFAULTS_3D)
    o.update(opts or {})
    if o["activated"]:
        flags |= _capi.FLAG_ACTIVATED_INPUTS
    return _capi.RenderDesc(_capi.MODE_3D if mode == "3d" else _capi.MODE_2D, int(width), int(height), int(n_frames),
                            int(n_gauss), int(n_views), int(flags), o["near_plane"], o["far_plane"], o["radius_clip"],
                            o["eps2d"])


class SavedForward:
    """Owns a ps_saved handle; released when dropped."""

    def __init__(self, handle, device):
        self.handle, self.device = handle, device

    def info(self) -> _capi.SavedInfo:
        out = _capi.SavedInfo()
        _capi.check(_capi.load().ps_saved_info_get(self.handle, ctypes.byref(out)), "ps_saved_info_get")
        return out

    def tap(self, name: str) -> torch.Tensor:
        """Copy one binning / raster intermediate to a torch tensor (bit-exact parity checks)."""
        info = self.info()
        M, VN = int(info.n_isect), info.n_views * info.n_gauss
        shapes = dict(isect_keys=((M,), torch.int64), flatten_ids=((M,), torch.int32),
                      tile_offsets=((info.n_views * info.tiles_x * info.tiles_y + 1,), torch.int32),
                      last_ids=((info.n_views, info.height, info.width), torch.int32),
                      tiles_touched=((VN,), torch.int32), rec0=((VN, 4), torch.float32), rec1=((VN, 4), torch.float32),
                      rec2=((VN, 4), torch.float32), depth=((VN if info.mode == 3 else 0,), torch.int32))
        shape, dtype = shapes[name]
        out = torch.empty(shape, dtype=dtype, device=self.device)
        if out.numel():
            _capi.check(_capi.load().ps_saved_copy(_capi.context(self.device), self.handle, _capi.TAPS[name],
                                                   _capi.ptr(out), out.numel() * out.element_size(),
                                                   _capi.stream_ptr(self.device)), "ps_saved_copy")
        return out

    def release(self):
        if self.handle is not None:
            _capi.load().ps_saved_release(_capi.context(self.device), self.handle, _capi.stream_ptr(self.device))
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _check_inputs(mode, params, view_frame, viewmats, Ks, background):
    P = 14 if mode == "3d" else 9
    if params.dim() != 3 or params.shape[2] != P:
        raise ValueError(f"Expected {P} parameters per Gaussian, got {params.shape[-1]}")
    if not params.is_cuda:
        raise RuntimeError("pose_splatter_b200 renders on CUDA devices only; there is no CPU path")
    if mode == "3d" and (viewmats is None or Ks is None):
        raise ValueError("3D rendering needs viewmats [V,4,4] and Ks [V,3,3]")
    if background.shape != (3,):
        raise ValueError(f"Expected color shape (3,), got {background.shape}")
    if view_frame.dim() != 1:
        raise ValueError(f"Expected view_frame [V], got {tuple(view_frame.shape)}")
    V = int(view_frame.shape[0])
    if mode == "3d" and (tuple(viewmats.shape) != (V, 4, 4) or tuple(Ks.shape) != (V, 3, 3)):
        raise ValueError(f"Expected viewmats [{V},4,4] and Ks [{V},3,3], got {tuple(viewmats.shape)} and {tuple(Ks.shape)}")
    if not view_frame.is_cuda and V > 0 and (int(view_frame.min()) < 0 or int(view_frame.max()) >= params.shape[0]):
        # a host-side map is checked for free; a device-side one is guarded inside the kernels (such views render nothing)
        raise ValueError(f"view_frame entries must lie in [0, {params.shape[0]})")


def forward_raw(mode, params, view_frame, viewmats, Ks, background, width, height, flags=0, want_counts=False, opts=None):
    """Direct call of ps_forward. Returns rgb [V,H,W,3], alpha [V,H,W], n_contrib or None, SavedForward or None."""
    dev = params.device
    F, N, _ = params.shape
    V = int(view_frame.shape[0])
    rgb = torch.empty(V, height, width, 3, dtype=torch.float32, device=dev)
    alpha = torch.empty(V, height, width, dtype=torch.float32, device=dev)
    counts = torch.empty(V, height, width, dtype=torch.int32, device=dev) if want_counts else None
    desc = _desc(mode, width, height, F, N, V, flags, opts)
    handle = ctypes.c_void_p()
    _capi.check(_capi.load().ps_forward(_capi.context(dev), ctypes.byref(desc), _capi.ptr(params), _capi.ptr(view_frame),
                                        _capi.ptr(viewmats), _capi.ptr(Ks), _capi.ptr(background), _capi.ptr(rgb),
                                        _capi.ptr(alpha), _capi.ptr(counts), ctypes.byref(handle),
                                        _capi.stream_ptr(dev)), "ps_forward")
    saved = SavedForward(handle, dev) if handle.value else None
    return rgb, alpha, counts, saved


def render_views_rgba8(mode: str, params: torch.Tensor, view_frame: torch.Tensor, width: int, height: int,
                       background: torch.Tensor, viewmats: Optional[torch.Tensor] = None, Ks: Optional[torch.Tensor] = None,
                       **opts) -> torch.Tensor:
    """Inference render straight to uint8 RGBA [V,H,W,4], quantised like the reference's evaluation writer
    (scripts/utils/evaluate_model.py:110-113).  Not differentiable."""
    mode = mode.lower()
    _check_inputs(mode, params, view_frame, viewmats, Ks, background)
    dev = params.device
    p = params.detach().contiguous().float()
    vf, vm, Kd, bg = _prep(view_frame, torch.int32, dev), _prep(viewmats, torch.float32, dev), _prep(Ks, torch.float32, dev), \
        _prep(background, torch.float32, dev)
    F, N, _ = p.shape
    V = int(vf.shape[0])
    out = torch.empty(V, height, width, 4, dtype=torch.uint8, device=dev)
    desc = _desc(mode, width, height, F, N, V, 0, opts)
    _capi.check(_capi.load().ps_forward_rgba8(_capi.context(dev), ctypes.byref(desc), _capi.ptr(p), _capi.ptr(vf), _capi.ptr(vm),
                                              _capi.ptr(Kd), _capi.ptr(bg), _capi.ptr(out), _capi.stream_ptr(dev)),
                "ps_forward_rgba8")
    return out


def backward_raw(saved: SavedForward, params, view_frame, viewmats, Ks, background, d_rgb, d_alpha):
    d_params = torch.empty_like(params)
    dev = params.device
    _capi.check(_capi.load().ps_backward(_capi.context(dev), saved.handle, _capi.ptr(params), _capi.ptr(view_frame),
                                         _capi.ptr(viewmats), _capi.ptr(Ks), _capi.ptr(background), _capi.ptr(d_rgb),
                                         _capi.ptr(d_alpha), _capi.ptr(d_params), _capi.stream_ptr(dev)), "ps_backward")
    return d_params


def _prep(t, dtype, dev):
    if t is None:
        return None
    return t.detach().to(device=dev, dtype=dtype).contiguous()


def backward_peer_raw(saved: SavedForward, params, viewmats, Ks, background, d_rgb, d_alpha, rank_ptrs, my_rank, world):
    """ps_backward_peer: finished rows are pushed into the staging buffer of each frame's owner rank (peer memory).
    rank_ptrs: int64 device tensor [world] holding every rank's staging-buffer address."""
    dev = params.device
    _capi.check(_capi.load().ps_backward_peer(_capi.context(dev), saved.handle, _capi.ptr(params), _capi.ptr(viewmats),
                                              _capi.ptr(Ks), _capi.ptr(background), _capi.ptr(d_rgb), _capi.ptr(d_alpha),
                                              _capi.ptr(rank_ptrs), int(my_rank), int(world), _capi.stream_ptr(dev)),
                "ps_backward_peer")


def peer_sum_raw(stage: torch.Tensor, out: torch.Tensor):
    """out [Fo,N,P] = sum over the world slots of this rank's staging buffer [world,Fo,N,P]."""
    dev = stage.device
    _capi.check(_capi.load().ps_peer_sum(_capi.context(dev), _capi.ptr(stage), int(stage.shape[0]), out.numel(), _capi.ptr(out),
                                         _capi.stream_ptr(dev)), "ps_peer_sum")
    return out


class _RenderViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, view_frame, viewmats, Ks, background, mode, width, height, opts):
        p = params.detach().contiguous().float()
        need = params.requires_grad
        rgb, alpha, _, saved = forward_raw(mode, p, view_frame, viewmats, Ks, background, width, height,
                                           _capi.FLAG_SAVE_FOR_BACKWARD if need else 0, False, opts)
        ctx.saved_fwd = saved
        ctx.aux = (p, view_frame, viewmats, Ks, background)  # the library keeps its own copy of the colour for the backward
        ctx.in_dtype = params.dtype
        return rgb, alpha

    @staticmethod
    def backward(ctx, d_rgb, d_alpha):
        p, view_frame, viewmats, Ks, background = ctx.aux
        d_params = backward_raw(ctx.saved_fwd, p, view_frame, viewmats, Ks, background,
                                d_rgb.contiguous().float(), d_alpha.contiguous().float())
        ctx.saved_fwd.release()
        return d_params.to(ctx.in_dtype), None, None, None, None, None, None, None, None


def render_views_vjp(mode: str, params: torch.Tensor, view_frame: torch.Tensor, width: int, height: int,
                     background: torch.Tensor, d_rgb: torch.Tensor, d_alpha: torch.Tensor,
                     viewmats: Optional[torch.Tensor] = None, Ks: Optional[torch.Tensor] = None, **opts):
    """Forward and vector-Jacobian product in one call, without an autograd graph: for losses whose cotangents do not
    depend on the render (the benchmark's L = sum(w_rgb * rgb) + sum(w_a * alpha)) or are produced by a
    fused loss.  Returns rgb [V,H,W,3], alpha [V,H,W], d_params [F,N,P] = (d_rgb, d_alpha)^T d(rgb, alpha)/d params."""
    mode = mode.lower()
    if mode not in ("2d", "3d"):
        raise ValueError(f"Unknown renderer mode: '{mode}'. Expected '2d' or '3d'.")
    _check_inputs(mode, params, view_frame, viewmats, Ks, background)
    dev = params.device
    p = params.detach().contiguous().float()
    vf, vm, Kd, bg = _prep(view_frame, torch.int32, dev), _prep(viewmats, torch.float32, dev), _prep(Ks, torch.float32, dev), \
        _prep(background, torch.float32, dev)
    V = int(vf.shape[0])
    if d_rgb.shape != (V, height, width, 3) or d_alpha.shape != (V, height, width):
        raise ValueError(f"Expected cotangents [V,H,W,3] and [V,H,W], got {tuple(d_rgb.shape)} and {tuple(d_alpha.shape)}")
    rgb, alpha, _, saved = forward_raw(mode, p, vf, vm, Kd, bg, width, height, _capi.FLAG_SAVE_FOR_BACKWARD, False, opts)
    d_params = backward_raw(saved, p, vf, vm, Kd, bg, _prep(d_rgb, torch.float32, dev), _prep(d_alpha, torch.float32, dev))
    saved.release()
    return rgb, alpha, d_params.to(params.dtype)


def render_views(mode: str, params: torch.Tensor, view_frame: torch.Tensor, width: int, height: int,
                 background: torch.Tensor, viewmats: Optional[torch.Tensor] = None, Ks: Optional[torch.Tensor] = None,
                 **opts):
    """Render V views in one launch sequence.

    params [F,N,14|9] (raw rows, activations applied inside), view_frame [V] int, viewmats [V,4,4],
    Ks [V,3,3], background [3]  ->  rgb [V,H,W,3], alpha [V,H,W].  Differentiable w.r.t. params.
    """
    mode = mode.lower()
    if mode not in ("2d", "3d"):
        raise ValueError(f"Unknown renderer mode: '{mode}'. Expected '2d' or '3d'.")
    _check_inputs(mode, params, view_frame, viewmats, Ks, background)
    dev = params.device
    view_frame = _prep(view_frame, torch.int32, dev)
    viewmats = _prep(viewmats, torch.float32, dev)
    Ks = _prep(Ks, torch.float32, dev)
    background = _prep(background, torch.float32, dev)
    return _RenderViews.apply(params, view_frame, viewmats, Ks, background, mode, int(width), int(height), opts)
