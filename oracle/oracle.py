"""ctypes front-end of the CPU oracle (oracle/ps_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Per-view functions mirror the stages of the hot path:
project -> emit keys -> stable sort -> tile ranges -> rasterize fwd -> rasterize bwd ->
projection/activation bwd.  See ps_oracle.c for the reference citations of each stage.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

TILE = 16
DEFAULTS_3D = dict(near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, activated=False)


def build(force: bool = False) -> Path:
    so = _HERE / "libps_oracle.so"
    src = _HERE / "ps_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.ora_tile_bits.restype = ctypes.c_int
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def _f(a):
    return _p(a, ctypes.c_float)


def _i(a):
    return _p(a, ctypes.c_int32)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def tile_grid(W, H):
    return (W + TILE - 1) // TILE, (H + TILE - 1) // TILE


def tile_bits(W, H):
    return int(lib().ora_tile_bits(int(W), int(H)))


def math_probe(x):
    x = _c32(x)
    outs = [np.empty_like(x) for _ in range(5)]
    lib().ora_math_probe(_f(x), ctypes.c_int(x.size), *[_f(o) for o in outs])
    return dict(zip(("exp", "log", "sigmoid", "sin", "cos"), outs))


def project(mode, params, W, H, viewmat=None, K=None, **opts):
    """Stage K1: per-Gaussian table for one view. Returns dict of arrays."""
    params = _c32(params)
    N = params.shape[0]
    geom = np.zeros((N, 8), np.float32)
    rgb = np.zeros((N, 3), np.float32)
    rect = np.zeros((N, 4), np.int32)
    trect = np.zeros((N, 4), np.int32)
    low = np.zeros(N, np.uint32)
    tiles = np.zeros(N, np.int32)
    if mode == "3d":
        o = dict(DEFAULTS_3D)
        o.update(opts)
        V = _c32(viewmat).reshape(16)
        Kf = _c32(K).reshape(9)
        lib().ora3d_project(_f(params), N, _f(V), _f(Kf), W, H,
                            ctypes.c_float(o["near_plane"]), ctypes.c_float(o["far_plane"]),
                            ctypes.c_float(o["radius_clip"]), ctypes.c_float(o["eps2d"]),
                            _f(geom), _f(rgb), _i(rect), _i(trect), _p(low, ctypes.c_uint32), _i(tiles),
                            ctypes.c_int(int(bool(o["activated"]))))
    else:
        lib().ora2d_project(_f(params), N, W, H, _f(geom), _f(rgb), _i(rect), _i(trect),
                            _p(low, ctypes.c_uint32), _i(tiles))
    return dict(geom=geom, rgb=rgb, rect=rect, tile_rect=trect, low=low, tiles=tiles)


def adapter3d(params, v_act=None):
    """The 3D adapter stage alone (src/gaussian_renderer.py:183-193): activated [N,14] = means | scales | quats |
    colours | opacity as gsplat's rasterization() receives them; with v_act [N,14] also J^T v_act (float64)."""
    params = _c32(params)
    N = params.shape[0]
    act = np.zeros((N, 14), np.float32)
    dp = ctypes.POINTER(ctypes.c_double)
    if v_act is None:
        lib().ora3d_adapter(_f(params), N, _f(act), None, None)
        return act
    v = np.ascontiguousarray(v_act, np.float64)
    d = np.zeros((N, 14), np.float64)
    lib().ora3d_adapter(_f(params), N, _f(act), v.ctypes.data_as(dp), d.ctypes.data_as(dp))
    return act, d


def bin_view(tab, W, H, view=0, n_stride=None):
    """Stages K2-K4 for one view: keys/vals (sorted) and tile offsets [n_tiles+1]."""
    N = tab["tiles"].shape[0]
    M = int(tab["tiles"].sum())
    tb = tile_bits(W, H)
    tw, th = tile_grid(W, H)
    keys = np.zeros(M, np.int64)
    vals = np.zeros(M, np.int32)
    if n_stride is None:
        n_stride = N
    lib().ora_emit(N, W, int(view), int(n_stride), _i(tab["tile_rect"]), _p(tab["low"], ctypes.c_uint32),
                   tb, _p(keys, ctypes.c_int64), _i(vals))
    unsorted = (keys.copy(), vals.copy())
    lib().ora_sort_pairs(_p(keys, ctypes.c_int64), _i(vals), ctypes.c_size_t(M))
    offsets = np.zeros((view + 1) * tw * th + 1, np.int32)
    lib().ora_tile_ranges(_p(keys, ctypes.c_int64), ctypes.c_size_t(M), int(view) + 1, tw * th, tb, _i(offsets))
    offsets = offsets[view * tw * th:]
    return dict(keys=keys, vals=vals, offsets=offsets, M=M, unsorted=unsorted, tile_bits=tb)


def raster_fwd(mode, tab, binned, W, H, bg, id_base=0):
    rgb = np.zeros((H, W, 3), np.float32)
    alpha = np.zeros((H, W), np.float32)
    ncon = np.zeros((H, W), np.int32)
    last = np.zeros((H, W), np.int32)
    bg = _c32(bg)
    if mode == "3d":
        lib().ora3d_raster_fwd(W, H, _f(tab["geom"]), _f(tab["rgb"]), _i(binned["vals"]), int(id_base),
                               _i(binned["offsets"]), _f(bg), _f(rgb), _f(alpha), _i(ncon), _i(last))
    else:
        lib().ora2d_raster_fwd(W, H, _f(tab["geom"]), _f(tab["rgb"]), _i(tab["rect"]), _i(binned["vals"]),
                               int(id_base), _i(binned["offsets"]), _f(bg), _f(rgb), _f(alpha), _i(ncon), _i(last))
    return dict(rgb=rgb, alpha=alpha, n_contrib=ncon, last=last)


def raster_bwd(mode, tab, binned, fwd, W, H, bg, w_rgb, w_a, id_base=0):
    N = tab["tiles"].shape[0]
    acc = np.zeros((N, 9), np.float64)
    bg = _c32(bg)
    w_rgb = _c32(w_rgb)
    w_a = _c32(w_a)
    dp = ctypes.POINTER(ctypes.c_double)
    if mode == "3d":
        lib().ora3d_raster_bwd(W, H, _f(tab["geom"]), _f(tab["rgb"]), _i(binned["vals"]), int(id_base),
                               _i(binned["offsets"]), _f(bg), _i(fwd["last"]), _f(w_rgb), _f(w_a),
                               acc.ctypes.data_as(dp))
    else:
        lib().ora2d_raster_bwd(W, H, _f(tab["geom"]), _f(tab["rgb"]), _i(tab["rect"]), _i(binned["vals"]),
                               int(id_base), _i(binned["offsets"]), _f(bg), _i(fwd["last"]), _f(w_rgb), _f(w_a),
                               acc.ctypes.data_as(dp))
    return acc


def project_bwd(mode, params, tab, acc, W, H, viewmat=None, K=None, d_params=None, **opts):
    params = _c32(params)
    N, P = params.shape
    if d_params is None:
        d_params = np.zeros((N, P), np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    if mode == "3d":
        o = dict(DEFAULTS_3D)
        o.update(opts)
        V = _c32(viewmat).reshape(16)
        Kf = _c32(K).reshape(9)
        lib().ora3d_project_bwd(_f(params), N, _f(V), _f(Kf), W, H,
                                ctypes.c_float(o["near_plane"]), ctypes.c_float(o["far_plane"]),
                                ctypes.c_float(o["radius_clip"]), ctypes.c_float(o["eps2d"]),
                                acc.ctypes.data_as(dp), d_params.ctypes.data_as(dp), ctypes.c_int(int(bool(o["activated"]))))
    else:
        lib().ora2d_project_bwd(_f(params), N, _f(tab["geom"]), acc.ctypes.data_as(dp), d_params.ctypes.data_as(dp))
    return d_params


def render(mode, params, W, H, bg, viewmat=None, K=None, w_rgb=None, w_a=None, **opts):
    """One view end to end.  With cotangents (w_rgb [H,W,3], w_a [H,W]) also returns d_params (float64)."""
    tab = project(mode, params, W, H, viewmat, K, **opts)
    binned = bin_view(tab, W, H)
    fwd = raster_fwd(mode, tab, binned, W, H, bg)
    out = dict(tab=tab, binned=binned, **fwd)
    if w_rgb is not None:
        acc = raster_bwd(mode, tab, binned, fwd, W, H, bg, w_rgb, w_a)
        out["acc"] = acc
        out["d_params"] = project_bwd(mode, params, tab, acc, W, H, viewmat, K, **opts)
    return out


def render_views(mode, params_f, view_frame, W, H, bg, viewmats=None, Ks=None, w_rgb=None, w_a=None, **opts):
    """Batched restatement: params_f [F,N,P], view_frame [V] -> per-view outputs and d_params [F,N,P].

    Keys carry the view id above the tile bits exactly like the product's batched launch.
    """
    params_f = _c32(params_f)
    F, N, P = params_f.shape
    V = len(view_frame)
    tb = tile_bits(W, H)
    outs = []
    d_params = np.zeros((F, N, P), np.float64) if w_rgb is not None else None
    all_keys, all_vals, all_off = [], [], []
    base = 0
    for v in range(V):
        f = int(view_frame[v])
        vm = viewmats[v] if viewmats is not None else None
        Kv = Ks[v] if Ks is not None else None
        o = render(mode, params_f[f], W, H, bg, vm, Kv,
                   None if w_rgb is None else w_rgb[v], None if w_a is None else w_a[v], **opts)
        if d_params is not None:
            d_params[f] += o["d_params"]
        all_keys.append(o["binned"]["keys"] | (np.int64(v) << np.int64(32 + tb)))
        all_vals.append(o["binned"]["vals"] + np.int32(v * N))
        all_off.append(o["binned"]["offsets"][:-1] + base)
        base += o["binned"]["M"]
        outs.append(o)
    return dict(views=outs, d_params=d_params,
                keys=np.concatenate(all_keys) if V else np.zeros(0, np.int64),
                vals=np.concatenate(all_vals) if V else np.zeros(0, np.int32),
                offsets=np.concatenate(all_off + [np.array([base], np.int32)]).astype(np.int32),
                rgb=np.stack([o["rgb"] for o in outs]) if V else None,
                alpha=np.stack([o["alpha"] for o in outs]) if V else None,
                n_contrib=np.stack([o["n_contrib"] for o in outs]) if V else None)
