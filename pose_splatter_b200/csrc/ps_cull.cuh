// ps_cull.cuh -- conservative footprint tests shared by the binning (ps_bin.cu) and the rasterizers (ps_raster.cu).
// None of them decides a result: they only drop (pixel block, splat) pairs whose exact per-pixel test
// (3D: sigma <= log(255 * opacity) + slack, 2D: q <= L) is guaranteed to fail for every pixel of the block.
#pragma once

#include "ps_contract.cuh"

#define PS_THR_SLACK 0.01f      // slack on sigma <= log(255*opacity): covers the rounding of exp / log / sigma
#define PS_CONIC_ERR 1.0e-6f    // 16 ulp of the term magnitudes: rounding of the conic coefficients and of the form's evaluation
#define PS_SLOT_MASK_SHIFT 24   // list slot word = depth rank (3D) / row index (2D) in the low 24 bits | 8-bit block mask

// 2D: q = dxr^2 iax + dyr^2 iay with (dxr, dyr) = R (dx, dy) is the quadratic form hA dx^2 + B dx dy + hC dy^2
__device__ __forceinline__ void ps_conic2d(const float4 &r1, float &hA, float &B, float &hC)
{
    const float cc = r1.x * r1.x, ss = r1.y * r1.y;
    hA = cc * r1.z + ss * r1.w;
    hC = ss * r1.z + cc * r1.w;
    B = 2.0f * r1.x * r1.y * (r1.z - r1.w);
}

// Which of the eight 8x4 pixel blocks of tile (tx, ty) can the splat contribute to?  `half` = 0.5 (3D pixel centres) or
// 0 (2D: integer pixel centres).  sigma(u) = hA ux^2 + hC uy^2 + B ux uy (u = pixel - mean) is convex with its minimum
// at u = 0, so over a box that does not contain 0 the minimum lies on an edge facing the mean; along such an edge it is
// a 1-D parabola.  The per-column / per-row terms are shared by the blocks: their pixel-centre boxes are bounded by
// 4 vertical and 8 horizontal lines.  NaN / inf (degenerate conics) count as a hit.
__device__ __forceinline__ uint32_t ps_block_mask8(float gx, float gy, float hA, float B, float hC, float thr, float half, int tx, int ty)
{
    const float kx = __fdividef(-B, 2.0f * hC), ky = __fdividef(-B, 2.0f * hA);
    const float X0 = ((float)(tx * PS_TILE) + half) - gx, Y0 = ((float)(ty * PS_TILE) + half) - gy;
    // The quadratic form of a needle (2D rows with sigma_x << sigma_y, giant 3D splats) is evaluated here as a sum of
    // terms that nearly cancel, where the exact per-pixel test rotates first: the threshold is widened by a bound of that
    // rounding error over the tile (PS_CONIC_ERR * the sum of the term magnitudes), which is nothing (< 1e-3) for
    // ordinary splats and opens the test up for ill-conditioned ones instead of dropping their blocks.
    const float xm = fmaxf(fabsf(X0), fabsf(X0 + 15.0f)), ym = fmaxf(fabsf(Y0), fabsf(Y0 + 15.0f));
    const float lim = thr * 1.0001f + 2.0f * PS_THR_SLACK + PS_CONIC_ERR * ((hA * xm + fabsf(B) * ym) * xm + hC * ym * ym);
    float ux0[2], ux1[2], cx[2], tx_[2], ax[2], bx_[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        ux0[i] = X0 + 8.0f * i;
        ux1[i] = X0 + (8.0f * i + 7.0f);
        cx[i] = fminf(fmaxf(0.0f, ux0[i]), ux1[i]);
        tx_[i] = kx * cx[i];          // unconstrained minimiser along the vertical line ux = cx
        ax[i] = hA * cx[i] * cx[i];
        bx_[i] = B * cx[i];
    }
    float uy0[4], uy1[4], cy[4], ty_[4], ay[4], by_[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uy0[j] = Y0 + 4.0f * j;
        uy1[j] = Y0 + (4.0f * j + 3.0f);
        cy[j] = fminf(fmaxf(0.0f, uy0[j]), uy1[j]);
        ty_[j] = ky * cy[j];
        ay[j] = hC * cy[j] * cy[j];
        by_[j] = B * cy[j];
    }
    uint32_t m8 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = k & 1, j = k >> 1;
        const float t1 = fminf(fmaxf(tx_[i], uy0[j]), uy1[j]);
        const float s1 = ax[i] + (hC * t1 + bx_[i]) * t1;
        const float t2 = fminf(fmaxf(ty_[j], ux0[i]), ux1[i]);
        const float s2 = ay[j] + (hA * t2 + by_[j]) * t2;
        // s1 / s2 = minimum of sigma over the block's vertical / horizontal edge nearest to the mean.  No guards on
        // "mean inside the column / row range": there cx (cy) = 0 and s1 (s2) degenerates to sigma at the nearest point
        // of the other axis, which is >= the other minimum, so testing both is the same decision (a superset by at most
        // a rounding error); mean inside the block gives s1 = s2 = 0 <= lim.  NaN counts as a hit.
        const bool miss = (s1 > lim) && (s2 > lim);
        m8 |= miss ? 0u : (1u << k);
    }
    return m8;
}

// Can the splat pass sigma <= thr (3D) / q <= L (2D, conic from ps_conic2d) anywhere on the pixel centres [x0, x1] x [y0, y1]?
// Same argument as ps_block_mask8 for one box (the rasterizers' re-cull against the box of the pixels that are still live);
// the threshold carries the same bound of the form's rounding error.  NaN / inf (degenerate conics) count as a hit.
__device__ __forceinline__ bool ps_ellipse_hits_box(float gx, float gy, float hA, float B, float hC, float thr, float x0,
                                                    float x1, float y0, float y1)
{
    const float ux0 = x0 - gx, ux1 = x1 - gx;
    const float uy0 = y0 - gy, uy1 = y1 - gy;
    const float cx = fminf(fmaxf(0.0f, ux0), ux1), cy = fminf(fmaxf(0.0f, uy0), uy1);
    if (cx == 0.0f && cy == 0.0f) return true;
    float best = 3.0e38f;
    if (cx != 0.0f) {
        const float t = fminf(fmaxf(__fdividef(-B * cx, 2.0f * hC), uy0), uy1);
        best = hA * cx * cx + (hC * t + B * cx) * t;
    }
    if (cy != 0.0f) {
        const float t = fminf(fmaxf(__fdividef(-B * cy, 2.0f * hA), ux0), ux1);
        const float sv = hC * cy * cy + (hA * t + B * cy) * t;
        best = (sv < best || !(best == best)) ? sv : best;
    }
    const float xm = fmaxf(fabsf(ux0), fabsf(ux1)), ym = fmaxf(fabsf(uy0), fabsf(uy1));
    return !(best > thr * 1.0001f + 2.0f * PS_THR_SLACK + PS_CONIC_ERR * ((hA * xm + fabsf(B) * ym) * xm + hC * ym * ym));
}

// Bounding box of the footprint {u : hA ux^2 + B ux uy + hC uy^2 <= lim} around the mean, half-widths in pixels
// (slightly inflated).  A degenerate conic (not positive definite, NaN) gets an unbounded box.
__device__ __forceinline__ void ps_footprint_box(float hA, float B, float hC, float thr, float &ex, float &ey)
{
    const float lim = thr * 1.0001f + 2.0f * PS_THR_SLACK;
    const float det = hA * hC - 0.25f * B * B;
    ex = ey = 3.0e38f;
    if (det > 0.0f && lim >= 0.0f) {
        const float k = __fdividef(lim, det);
        ex = sqrtf(k * hC) * 1.0001f + 1e-3f;
        ey = sqrtf(k * hA) * 1.0001f + 1e-3f;
        if (!(ex == ex) || !(ey == ey)) ex = ey = 3.0e38f;
    }
}

// Footprint bounding box -> inclusive rectangle of 8x4 pixel blocks (global block coordinates: column = x / 8,
// row = y / 4) whose pixel centres it meets, clipped to the image; empty if c0 > c1 or r0 > r1.
struct PsBlockRect { int c0, c1, r0, r1; };
__device__ __forceinline__ PsBlockRect ps_block_rect(float gx, float gy, float ex, float ey, float half, int W, int H)
{
    // block column c holds the pixel centres [8c + half, 8c + 7 + half]; it meets [gx - ex, gx + ex] iff
    // 8c + half <= gx + ex and 8c + 7 + half >= gx - ex
    const float big = 1.0e9f;
    const float xl = fmaxf(gx - ex, -big), xh = fminf(gx + ex, big), yl = fmaxf(gy - ey, -big), yh = fminf(gy + ey, big);
    PsBlockRect r;
    r.c0 = max(0, (int)ceilf((xl - half - 7.0f) * 0.125f));
    r.c1 = min((W - 1) >> 3, (int)floorf((xh - half) * 0.125f));
    r.r0 = max(0, (int)ceilf((yl - half - 3.0f) * 0.25f));
    r.r1 = min((H - 1) >> 2, (int)floorf((yh - half) * 0.25f));
    return r;
}

// the blocks of tile (tx, ty) inside the block rectangle: a superset of the blocks the exact test (ps_block_mask8)
// keeps, for a handful of integer instructions
__device__ __forceinline__ uint32_t ps_block_mask8_rect(const PsBlockRect &r, int tx, int ty)
{
    const int c = 2 * tx, q = 4 * ty;
    const uint32_t cols = ((r.c0 <= c && c <= r.c1) ? 1u : 0u) | ((r.c0 <= c + 1 && c + 1 <= r.c1) ? 2u : 0u);
    uint32_t m8 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (r.r0 <= q + j && q + j <= r.r1) m8 |= cols << (2 * j);
    return m8;
}

// blocks of tile (tx, ty) that have at least one pixel inside the W x H image
__device__ __forceinline__ uint32_t ps_blocks_inside8(int tx, int ty, int W, int H)
{
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (tx * PS_TILE + (k & 1) * 8 < W && ty * PS_TILE + (k >> 1) * 4 < H) m |= 1u << k;
    return m;
}
