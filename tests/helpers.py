"""Shared helpers of the test-suite (oracle access lives here: tests may use oracle/, the product may not)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"

RGB_TOL = 1e-4    # north_star: rendered RGB / alpha within 1e-4 max-abs
GRAD_TOL = 1e-3   # north_star: gradients within 1e-3 relative (per parameter column, max-abs normalised)


def golden_cotangents(seed, H, W):
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand(H, W, 3, generator=g) - 0.3, torch.rand(H, W, generator=g) - 0.3


def column_rel_err(got, want):
    got = np.asarray(got, np.float64).reshape(-1, got.shape[-1])
    want = np.asarray(want, np.float64).reshape(-1, want.shape[-1])
    scale = np.maximum(np.abs(want).max(0), 1e-20)
    return (np.abs(got - want).max(0) / scale)


def host_contract():
    """g++ build of the product's arithmetic contract header for CPU-side bit-exactness checks."""
    so = ROOT / "tests" / "_host_contract.so"
    src = ROOT / "tests" / "host_contract.cpp"
    hdr = ROOT / "pose_splatter_b200" / "csrc" / "ps_contract.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)], check=True)
    return ctypes.CDLL(str(so))


def hc_project(mode, params, W, H, V=None, K=None, near=0.01, far=1e10, clip=0.0, eps=0.3, activated=False):
    L = host_contract()
    fp, ip, up = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint32))
    params = np.ascontiguousarray(params, np.float32)
    N = len(params)
    r = np.zeros((N, 12), np.float32)
    tile = np.zeros((N, 4), np.int32)
    low = np.zeros(N, np.uint32)
    V = np.zeros(16, np.float32) if V is None else np.ascontiguousarray(V, np.float32).reshape(16)
    K = np.zeros(9, np.float32) if K is None else np.ascontiguousarray(K, np.float32).reshape(9)
    L.hc_project(3 if mode == "3d" else 2, params.ctypes.data_as(fp), N, V.ctypes.data_as(fp), K.ctypes.data_as(fp),
                 W, H, ctypes.c_float(near), ctypes.c_float(far), ctypes.c_float(clip), ctypes.c_float(eps),
                 r.ctypes.data_as(fp), tile.ctypes.data_as(ip), low.ctypes.data_as(up), int(bool(activated)))
    return r, tile, low


def records_from_oracle(mode, tab, table=False):
    """Oracle table -> the product's [N,12] record layout (rec0|rec1|rec2) for bit-exact comparison.

    table=False: the contract-level PsRecord (x, y, rx, ry | A, B, C, o | r, g, b, depth).  table=True: the layout
    the projection kernel stores in HBM for the 3D rasterizer: (x, y, thr, o | A/2, B, C/2, 0 | r, g, b, 0) with exact
    halvings and thr = log(255 * opacity); radii are covered by the tile-rectangle taps, depth bits by tap "depth"."""
    g, rgb, rect = tab["geom"], tab["rgb"], tab["rect"]
    N = g.shape[0]
    r = np.zeros((N, 12), np.float32)
    listed = tab["tiles"] > 0
    if mode == "3d":
        r[:, 0:2] = g[:, 0:2]
        r[:, 2:4] = rect[:, 0:2].astype(np.float32)
        r[:, 4:8] = g[:, 2:6]
        r[:, 8:11] = rgb
        r[:, 11] = g[:, 6]
        if table:
            from oracle import oracle as ora
            thr = ora.math_probe((g[:, 5] * np.float32(255.0)).astype(np.float32))["log"]
            r[:, 2] = np.where(g[:, 6] != 0, thr, np.float32(0.0))  # set only for Gaussians that survive every cull
            r[:, 3] = g[:, 5]
            r[:, 4] = np.float32(0.5) * g[:, 2]
            r[:, 6] = np.float32(0.5) * g[:, 4]
            r[:, 7] = 0.0
            r[:, 11] = 0.0
    else:
        r[:, 0:2] = g[:, 0:2]
        r[:, 4:8] = g[:, 2:6]
        r[:, 8:11] = rgb
        if table:   # (u, v, L, o | cos, sin, iax, iay | r, g, b, 0)
            r[:, 2] = g[:, 7]
            r[:, 3] = g[:, 6]
            r[:, 11] = 0.0
        else:       # contract-level PsRecord: pixel rectangle bits in r0, opacity in r2
            lo = (rect[:, 0].astype(np.uint32) | (rect[:, 1].astype(np.uint32) << 16))
            hi = (rect[:, 2].astype(np.uint32) | (rect[:, 3].astype(np.uint32) << 16))
            r[:, 2] = lo.view(np.float32)
            r[:, 3] = hi.view(np.float32)
            r[:, 11] = g[:, 6]
    return r, listed


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
