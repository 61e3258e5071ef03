"""Offline comparison of work mappings for the 3D rasterizers (no GPU): for a few synthetic c2 views, replays the
per-block lists chunk by chunk and counts, for every 32-entry chunk of every 8x4 pixel block,
  v4  : entries walked by the whole warp (entry reaches a still-live pixel)                      -> 32-lane iterations
  v6  : iterations of the per-pixel walk = max over pixel lanes of its candidate count (2 entries per iteration)
  bal : candidate pairs (balanced alpha evaluation, 32 pairs per iteration) + ordered pass = max candidates per pixel
and turns them into warp-instruction estimates with the per-iteration costs measured by ncu's source counters for the
existing kernels (v4: 64 per walked entry; v6: 110 per two-entry iteration + 290 per chunk set-up).
CPU only; lives under tests/ because it uses the oracle (test infrastructure) for projection and binning.  python tests/analysis/sim_raster_schemes.py [views]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as ora  # noqa: E402
from pose_splatter_b200 import synth  # noqa: E402

ALPHA_MIN, ALPHA_MAX, T_STOP, SLACK = 1.0 / 255.0, 0.999, 1e-4, 0.01


def view_stats(d, v):
    W, H = d["width"], d["height"]
    tab = ora.project("3d", d["params"][int(d["view_frame"][v])].numpy(), W, H, d["viewmats"][v].numpy(), d["Ks"][v].numpy())
    b = ora.bin_view(tab, W, H)
    g = tab["geom"]
    tw = (W + 15) // 16
    out = dict(chunks=0, staged=0, v4_walk=0, v6_iter=0, pairs=0, ord_iter=0, contrib=0, ent_iter=0)
    for tile in range(len(b["offsets"]) - 1):
        s, e = int(b["offsets"][tile]), int(b["offsets"][tile + 1])
        if e == s:
            continue
        ids = b["vals"][s:e]
        x, y, A, B, C, o = (g[ids, k].astype(np.float64) for k in range(6))
        thr = np.log(255.0 * o)
        tx, ty = tile % tw, tile // tw
        for blk in range(8):
            bx, by = tx * 16 + (blk & 1) * 8, ty * 16 + (blk >> 1) * 4
            if bx >= W or by >= H:
                continue
            px = bx + (np.arange(32) & 7) + 0.5
            py = by + (np.arange(32) >> 3) + 0.5
            inside = (px < W) & (py < H)
            dx, dy = x[:, None] - px[None, :], y[:, None] - py[None, :]
            sig = 0.5 * (A[:, None] * dx * dx + C[:, None] * dy * dy) + B[:, None] * dx * dy
            cand = (sig >= 0) & (sig <= thr[:, None] + SLACK) & inside[None, :]
            keep = cand.any(1)  # the block list (the kernel's box test is marginally wider)
            if not keep.any():
                continue
            cand, sigk, ok = cand[keep], sig[keep], o[keep]
            alpha = np.minimum(ALPHA_MAX, ok[:, None] * np.exp(-sigk))
            contrib = cand & (alpha >= ALPHA_MIN)
            # per-pixel transmittance before every entry; a pixel is done once T (1 - alpha) <= 1e-4
            fac = np.where(contrib, 1.0 - alpha, 1.0)
            T_after = np.cumprod(fac, 0)
            stop = contrib & (T_after <= T_STOP)
            first_stop = np.where(stop.any(0), stop.argmax(0), len(cand))  # entry index at which the pixel terminates
            E = len(cand)
            for c0 in range(0, E, 32):
                live = inside & (first_stop >= c0)  # pixels still live at the start of the chunk
                if not live.any():
                    break
                em = cand[c0:c0 + 32] & live[None, :]
                out["chunks"] += 1
                out["staged"] += len(em)
                # v4: the warp walks an entry if it reaches the bounding box of the live pixels
                lx, ly = (np.arange(32) & 7)[live], (np.arange(32) >> 3)[live]
                box = ((np.arange(32) & 7) >= lx.min()) & ((np.arange(32) & 7) <= lx.max()) & \
                      ((np.arange(32) >> 3) >= ly.min()) & ((np.arange(32) >> 3) <= ly.max())
                out["v4_walk"] += int((cand[c0:c0 + 32] & box[None, :]).any(1).sum())
                per_pixel = em.sum(0)
                out["v6_iter"] += int(np.ceil(per_pixel.max() / 2)) if per_pixel.max() else 0
                out["pairs"] += int(em.sum())
                out["ord_iter"] += int(per_pixel.max())
                out["ent_iter"] += int(em.sum(1).max())
                out["contrib"] += int((contrib[c0:c0 + 32] & live[None, :]).sum())
    return out


def main():
    n_views = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    d = synth.make_views("c2", n_frames=(n_views + 5) // 6, n_cams=6, seed=1000)
    tot = None
    for v in range(n_views):
        s = view_stats(d, v)
        tot = s if tot is None else {k: tot[k] + s[k] for k in tot}
    c = tot["chunks"]
    print(f"{n_views} views: {c} chunks, per chunk: staged {tot['staged'] / c:.1f}, v4 walked {tot['v4_walk'] / c:.1f}, "
          f"candidate pairs {tot['pairs'] / c:.1f} (contributing {tot['contrib'] / c:.1f}), v6 two-entry iterations {tot['v6_iter'] / c:.2f}, "
          f"max candidates per pixel {tot['ord_iter'] / c:.2f}, max candidates per entry {tot['ent_iter'] / c:.2f}")
    v4 = tot["v4_walk"] / c * 64 + 150
    v6 = tot["v6_iter"] / c * 110 + 290 + 60
    bal = 290 + 60 + tot["ent_iter"] / c * 6 + np.ceil(tot["pairs"] / c / 32) * 58 + tot["ord_iter"] / c * 22
    print(f"estimated warp instructions per chunk: v4 {v4:.0f}, v6 {v6:.0f}, balanced alpha + ordered pass {bal:.0f}")


if __name__ == "__main__":
    main()
