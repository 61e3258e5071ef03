"""The conservative footprint tests of pose_splatter_b200/csrc/ps_cull.cuh, built for the host (tests/host_cull.cpp): they may
only drop (pixel block, splat) pairs that cannot pass the rasterizers' own per-pixel test.  Checked by brute force over the
32 pixel centres of every block with the contract's sigma arithmetic: exact blocks are a subset of ps_block_mask8 (the block
split, 3D pixel centres at +0.5 and 2D at integers) and of the bounding-rectangle masks of PS_BIN_MODE=bytes|split; the
exact-mask test must also stay tight (it decides how many entries the rasterizers stage)."""
import ctypes
import subprocess

import numpy as np

from helpers import ROOT


def _lib():
    so, src = ROOT / "tests" / "_host_cull.so", ROOT / "tests" / "host_cull.cpp"
    hdrs = [ROOT / "pose_splatter_b200" / "csrc" / h for h in ("ps_cull.cuh", "ps_contract.cuh")]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)], check=True)
    return ctypes.CDLL(str(so))


def _masks(lib, splats, half, tx, ty, W, H):
    n = len(splats)
    outs = [np.zeros(n, np.uint32) for _ in range(3)]
    fp, up = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)
    lib.hc_block_masks(splats.ctypes.data_as(fp), n, ctypes.c_float(half), tx, ty, W, H, *[o.ctypes.data_as(up) for o in outs])
    return outs


def _splats(rng, n, tx, ty, needle):
    """Records as the 3D projection stores them: mean, halved conic of (R diag(s^2) R^T + 0.3 I)^-1, thr = log(255 o)."""
    sx, sy = np.exp(rng.uniform(-2.5, 3.0, n)), np.exp(rng.uniform(-2.5, 3.0, n))
    if needle:
        sx = sx * 0.02
    th = rng.uniform(0, np.pi, n)
    c, s = np.cos(th), np.sin(th)
    a, b, d = c * c * sx * sx + s * s * sy * sy + 0.3, c * s * (sx * sx - sy * sy), s * s * sx * sx + c * c * sy * sy + 0.3
    det = a * d - b * b
    thr = np.log(255 * np.exp(rng.uniform(np.log(1 / 255), 0, n)))
    reach = 10 + 3.5 * np.sqrt(np.maximum(a, d))
    gx, gy = tx * 16 + 8 + rng.uniform(-1, 1, n) * reach, ty * 16 + 8 + rng.uniform(-1, 1, n) * reach
    return np.stack([gx, gy, 0.5 * d / det, -b / det, 0.5 * a / det, thr], 1).astype(np.float32)


def _bits(x):
    return int(np.unpackbits(x.view(np.uint8)).sum())


def test_block_masks_never_drop_a_block_the_pixel_test_accepts():
    lib = _lib()
    rng = np.random.default_rng(7)
    kept = exact = 0
    for rep in range(12):
        tx, ty = int(rng.integers(0, 18)), int(rng.integers(0, 16))
        sp = _splats(rng, 40000, tx, ty, needle=rep % 4 == 0)
        for half in (0.5, 0.0):
            mask8, rect8, exact8 = _masks(lib, sp, half, tx, ty, 288, 256)
            assert not (exact8 & ~mask8).any(), "ps_block_mask8 dropped a block with a passing pixel"
            assert not (exact8 & ~rect8).any(), "the bounding-rectangle mask dropped a block with a passing pixel"
            kept += _bits(mask8)
            exact += _bits(exact8)
    assert exact > 100000 and kept <= 1.03 * exact, (kept, exact)  # measured: 1.015


def test_block_masks_degenerate_records_count_as_hits():
    lib = _lib()
    nan, inf = float("nan"), float("inf")
    sp = np.array([[8.0, 8.0, nan, 0.0, 1.0, 3.0], [8.0, 8.0, 1.0, nan, 1.0, 3.0], [8.0, 8.0, 1.0, 0.0, 1.0, nan],
                   [nan, 8.0, 1.0, 0.0, 1.0, 3.0], [8.0, 8.0, 0.0, 0.0, 0.0, 3.0], [8.0, 8.0, -1.0, 0.0, -1.0, 3.0],
                   [8.0, 8.0, inf, 0.0, inf, 3.0]], np.float32)
    mask8, _, exact8 = _masks(lib, sp, 0.5, 0, 0, 288, 256)
    assert not (exact8 & ~mask8).any()
    assert (mask8[[0, 1, 2]] == 0xFF).all(), mask8  # NaN in the conic / threshold: every block is kept, the exact test decides
    # (a NaN mean -- row 3 -- never passes the pixel test, sigma is NaN: any mask is correct; the projection culls it anyway)
    assert mask8[4] == 0xFF and mask8[5] == 0xFF      # flat / concave "conics": sigma <= thr everywhere


def _splats_2d(rng, n, cx, cy, needle):
    """2D records around (cx, cy): u v L cos sin 1/ax 1/ay with ax = 2 sigma^2 + 1e-8, L = ln(o / tau)."""
    sx, sy = np.exp(rng.uniform(-3.0, 2.5, n)), np.exp(rng.uniform(-3.0, 2.5, n))
    if needle:
        sx, sy = needle * np.exp(rng.uniform(-0.3, 0.3, n)), np.exp(rng.uniform(0.0, 3.5, n))
    th = rng.uniform(-7.0, 7.0, n)
    L = np.log(np.exp(rng.uniform(np.log(2.0 ** -27), 0, n)) * 2.0 ** 28)
    reach = 10 + np.sqrt(2 * L) * np.maximum(sx, sy)
    u, v = cx + rng.uniform(-1, 1, n) * reach, cy + rng.uniform(-1, 1, n) * reach
    return np.stack([u, v, L, np.cos(th), np.sin(th), 1 / (2 * sx * sx + 1e-8), 1 / (2 * sy * sy + 1e-8)], 1).astype(np.float32)


def test_block_masks_2d_records_never_drop_a_block_the_footprint_test_accepts():
    """2D mode: the record holds (cos, sin, 1/ax, 1/ay) and the threshold L = ln(o / tau); the block split turns it into a
    conic (ps_conic2d) and the rasterizers test q <= L at integer pixel centres (ps_q2d, which rotates first).  Needles
    (sigma_x << sigma_y) make the conic form cancel catastrophically in fp32: the cull threshold carries a bound of that
    error (PS_CONIC_ERR), without which about 0.1 % (sigma_x = 0.01 px) to 9 % (3e-4 px) of such splats lost a block."""
    lib = _lib()
    rng = np.random.default_rng(11)
    fp, up = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)
    kept = exact = 0
    for rep, needle in enumerate([0, 0, 0, 0, 0, 0, 0.03, 0.01, 0.003, 0.001, 0.0003]):
        n = 40000
        tx, ty = int(rng.integers(0, 36)), int(rng.integers(0, 32))
        sp = _splats_2d(rng, n, tx * 16 + 8, ty * 16 + 8, needle)
        mask8, exact8 = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        lib.hc_block_masks_2d(sp.ctypes.data_as(fp), n, tx, ty, mask8.ctypes.data_as(up), exact8.ctypes.data_as(up))
        assert not (exact8 & ~mask8).any(), f"ps_block_mask8 (2D records, needle {needle}) dropped a block with a pixel inside the footprint"
        if not needle:  # tightness matters for ordinary splats only: it decides how many entries the rasterizers stage
            kept += _bits(mask8)
            exact += _bits(exact8)
    assert exact > 100000 and kept <= 1.10 * exact, (kept, exact)  # measured 1.06 (independent axes: aspect ratios up to 250)
    # the case that exposed the cancellation (tile (16, 21), sigma_x = 0.0015 px, sigma_y = 7.8 px)
    sp = np.array([[2.9264554e+02, 3.2695892e+02, 1.6998478e+01, 4.5190281e-01, 8.9206719e-01, 2.1190447e+05, 8.2178069e-03]], np.float32)
    mask8, exact8 = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
    lib.hc_block_masks_2d(sp.ctypes.data_as(fp), 1, 16, 21, mask8.ctypes.data_as(up), exact8.ctypes.data_as(up))
    assert exact8[0] == 0b10000 and not (exact8[0] & ~mask8[0])


def test_box_recull_never_drops_an_entry_with_a_passing_live_pixel():
    """ps_ellipse_hits_box: the rasterizers' re-cull of a staged entry against the box of the pixels that are still live."""
    lib = _lib()
    fp, ip, bp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8)
    rng = np.random.default_rng(3)
    for needle in (0, 0, 0.1, 0.01, 0.001):
        n = 100000
        bx, by = rng.integers(0, 72, n) * 8, rng.integers(0, 128, n) * 4
        x0 = bx + rng.integers(0, 8, n)
        x1 = np.minimum(bx + 7, x0 + rng.integers(0, 8, n))
        y0 = by + rng.integers(0, 4, n)
        y1 = np.minimum(by + 3, y0 + rng.integers(0, 4, n))
        sp = _splats_2d(rng, n, bx + 4, by + 2, needle)
        boxes = np.stack([x0, x1, y0, y1], 1).astype(np.int32)
        hit, ex = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        lib.hc_box_hits_2d(sp.ctypes.data_as(fp), boxes.ctypes.data_as(ip), n, hit.ctypes.data_as(bp), ex.ctypes.data_as(bp))
        assert ex.sum() > 20 and not (ex & ~hit & 1).any(), f"re-cull dropped a passing entry (needle {needle})"
        if not needle:
            assert hit.sum() <= 1.05 * ex.sum()


def test_2d_pixel_rectangle_contains_every_passing_pixel():
    """ps_project2d lists a row on the tiles of its pixel rectangle: every integer pixel of the image
    that passes q <= L in the contract's arithmetic must lie inside it, needles included; unlisted rows pass nowhere."""
    from helpers import hc_project, host_contract
    hc = host_contract()
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(4)
    W, H = 96, 80
    ys, xs = np.mgrid[0:H, 0:W]
    px, py = xs.ravel().astype(np.float32), ys.ravel().astype(np.float32)
    q = np.zeros(W * H, np.float32)
    total = 0
    for needle in (0, 0.03, 0.003, 0.0003):
        n = 500
        lsx, lsy = rng.uniform(-3, 2.5, n), rng.uniform(-3, 2.5, n)
        if needle:
            lsx, lsy = np.log(needle) + rng.uniform(-0.3, 0.3, n), rng.uniform(0, 3.0, n)
        rows = np.stack([rng.uniform(-10, W + 10, n), rng.uniform(-10, H + 10, n), lsx, lsy, rng.uniform(-7, 7, n),
                         rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.normal(0, 3, n)], 1).astype(np.float32)
        rec, tile, _ = hc_project("2d", rows, W, H)
        for i in range(n):
            o = 1 / (1 + np.exp(-np.float64(rows[i, 8])))
            if o <= 2.0 ** -28:
                continue
            listed = tile[i, 2] > tile[i, 0] and tile[i, 3] > tile[i, 1]
            if listed:
                g = np.array([rows[i, 0], rows[i, 1], rec[i, 4], rec[i, 5], rec[i, 6], rec[i, 7]], np.float32)
            else:  # no record was written: rebuild its pair constants from the row
                ax = np.float32(2) * np.exp(rows[i, 2]) ** 2 + np.float32(1e-8)
                ay = np.float32(2) * np.exp(rows[i, 3]) ** 2 + np.float32(1e-8)
                g = np.array([rows[i, 0], rows[i, 1], np.cos(rows[i, 4]), np.sin(rows[i, 4]), 1 / ax, 1 / ay], np.float32)
            hc.hc_pairs2d(g.ctypes.data_as(fp), px.ctypes.data_as(fp), py.ctypes.data_as(fp), W * H, q.ctypes.data_as(fp))
            passing = np.nonzero(q <= np.float32(np.log(o * 2.0 ** 28)))[0]
            total += len(passing)
            if not len(passing):
                continue
            assert listed, f"row {i} (needle {needle}) passes at {len(passing)} pixels but is listed nowhere"
            b0, b1 = rec[i, 2].view(np.uint32), rec[i, 3].view(np.uint32)
            x0, y0, x1, y1 = int(b0 & 0xFFFF), int(b0 >> 16), int(b1 & 0xFFFF), int(b1 >> 16)
            pxs, pys = passing % W, passing // W
            assert pxs.min() >= x0 and pxs.max() <= x1 and pys.min() >= y0 and pys.max() <= y1, (i, needle)
    assert total > 50000
