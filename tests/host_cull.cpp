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
}
