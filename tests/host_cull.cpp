// Host build of the product's conservative footprint tests (pose_splatter_b200/csrc/ps_cull.cuh) so that the CPU
// test-suite can check, without a GPU, the one property they must have: a (pixel block, splat) pair is never dropped
// if any pixel centre of the block passes the rasterizers' own per-pixel test.
// Build: g++ -O2 -shared -fPIC -ffp-contract=off -o tests/_host_cull.so tests/host_cull.cpp
#include <algorithm>
#include <math.h>
#include <stdint.h>
using std::max;
using std::min;
#define __device__
#define __forceinline__ inline
static inline float __fdividef(float a, float b) { return a / b; }
struct float4 { float x, y, z, w; };
#include "../pose_splatter_b200/csrc/ps_cull.cuh"

extern "C" {
// splats [n,6] = gx gy hA B hC thr (3D record layout: halved conic, thr = log(255 * opacity)); tile (tx, ty)
//   mask8  [n]: ps_block_mask8
//   rect8  [n]: ps_block_mask8_rect of the footprint box (the cheap superset used by PS_BIN_MODE=bytes|split)
//   exact8 [n]: brute force over the 32 pixel centres of each block with the rasterizers' candidate test
//               sigma >= 0 && sigma <= thr + PS_THR_SLACK (ps_sigma3d, the contract's arithmetic)
void hc_block_masks(const float *splats, int n, float half, int tx, int ty, int W, int H, uint32_t *mask8, uint32_t *rect8,
                    uint32_t *exact8)
{
    for (int i = 0; i < n; ++i) {
        const float *s = splats + 6 * (size_t)i;
        mask8[i] = ps_block_mask8(s[0], s[1], s[2], s[3], s[4], s[5], half, tx, ty);
        float ex, ey;
        ps_footprint_box(s[2], s[3], s[4], s[5], ex, ey);
        const PsBlockRect r = ps_block_rect(s[0], s[1], ex, ey, half, W, H);
        rect8[i] = ps_block_mask8_rect(r, tx, ty);
        uint32_t m = 0;
        for (int k = 0; k < 8; ++k)
            for (int p = 0; p < 32 && !((m >> k) & 1u); ++p) {
                const float px = (float)(tx * PS_TILE + (k & 1) * 8 + (p & 7)) + half;
                const float py = (float)(ty * PS_TILE + (k >> 1) * 4 + (p >> 3)) + half;
                float dx, dy;
                const float sg = ps_sigma3d(s[0], s[1], s[2], s[3], s[4], px, py, &dx, &dy);
                if (sg >= 0.0f && sg <= s[5] + PS_THR_SLACK) m |= 1u << k;
            }
        exact8[i] = m;
    }
}

// 2D records: splats [n,7] = u v L cos sin 1/ax 1/ay (rec0.xyz, rec1); integer pixel centres; exact test q <= L (ps_q2d)
void hc_block_masks_2d(const float *splats, int n, int tx, int ty, uint32_t *mask8, uint32_t *exact8)
{
    for (int i = 0; i < n; ++i) {
        const float *s = splats + 7 * (size_t)i;
        const float4 r1 = { s[3], s[4], s[5], s[6] };
        float hA, B, hC;
        ps_conic2d(r1, hA, B, hC);
        mask8[i] = ps_block_mask8(s[0], s[1], hA, B, hC, s[2], 0.0f, tx, ty);
        uint32_t m = 0;
        for (int k = 0; k < 8; ++k)
            for (int p = 0; p < 32 && !((m >> k) & 1u); ++p) {
                const float px = (float)(tx * PS_TILE + (k & 1) * 8 + (p & 7));
                const float py = (float)(ty * PS_TILE + (k >> 1) * 4 + (p >> 3));
                float a, b;
                if (ps_q2d(s[0], s[1], s[3], s[4], s[5], s[6], px, py, &a, &b) <= s[2]) m |= 1u << k;
            }
        exact8[i] = m;
    }
}

// the rasterizers' re-cull of one entry against a box of live pixels: splats as above (2D records), boxes [n,4] =
// x0 x1 y0 y1 (integer pixel coordinates, inclusive); hit [n] = ps_ellipse_hits_box, exact [n] = any pixel with q <= L
void hc_box_hits_2d(const float *splats, const int *boxes, int n, uint8_t *hit, uint8_t *exact)
{
    for (int i = 0; i < n; ++i) {
        const float *s = splats + 7 * (size_t)i;
        const int *b = boxes + 4 * (size_t)i;
        const float4 r1 = { s[3], s[4], s[5], s[6] };
        float hA, B, hC;
        ps_conic2d(r1, hA, B, hC);
        hit[i] = ps_ellipse_hits_box(s[0], s[1], hA, B, hC, s[2], (float)b[0], (float)b[1], (float)b[2], (float)b[3]) ? 1 : 0;
        uint8_t e = 0;
        for (int y = b[2]; y <= b[3] && !e; ++y)
            for (int x = b[0]; x <= b[1] && !e; ++x) {
                float a, c;
                if (ps_q2d(s[0], s[1], s[3], s[4], s[5], s[6], (float)x, (float)y, &a, &c) <= s[2]) e = 1;
            }
        exact[i] = e;
    }
}
}
