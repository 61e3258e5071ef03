// ps_loss.cu -- the per-view training loss that follows render() every step, fused with its own backward
// (SURVEY.md 8f-f1): soft IoU(alpha, mask) + img_lambda * L1(rgb, img) / sum(mask) + ssim_lambda * (1 - SSIM),
// producing the three loss terms and d_rgb / d_alpha -- the cotangents ps_backward takes -- in one call.
// Replaces scripts/training/train_script.py:30-36 (get_iou_loss), :129-133 (three torch / torchmetrics graphs and
// their autograd) of the reference.  SSIM = torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0) (:270),
// called as ssim(target_img, rgb): 11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03, windows entirely inside the
// image only (its reflection padding is cropped away again), variances clamped at 0.  torchmetrics is not in the
// reference tree: the algorithm is restated from its published source (oracle/loss_ref.py, parity unpinned).
//
// Launches per call, all views batched (default path; PS_LOSS_TILED=1 selects round 1's 32 x 32 tile kernels):
//   loss_reduce     per-view sums  I = sum a m, U = sum (a + m - a m), sum m                      [HBM, one read, 16-byte loads]
//   ssim_march      strip-marching SSIM: separable 11-tap windows of (p, q, pp, qq, pq), per window centre S and the
//                   three adjoints dS/d(mu_q), dS/d(E qq), dS/d(E pq) (scaled by -lambda / count); also sum |t - rgb|
//   loss_bwd_march  the adjoints filtered back with the same window (its transpose: the taps are symmetric), combined
//                   with the L1 sign term -> d_rgb [V,H,W,3]
//   iou_bwd         IoU quotient rule -> d_alpha [V,H,W]
//   loss_finalize   losses [V,3] = (iou, ssim, img)
#include "ps_internal.h"
#include <cstdlib>

namespace {

constexpr int LT = 32;          // tile edge (pixels)
constexpr int HALO = 5;         // (11 - 1) / 2
constexpr int LIN = LT + 2 * HALO; // 42
constexpr int LTHREADS = 256;
constexpr int NSTAT = 8;        // doubles per view: I, U, sum m, sum |t - rgb|, sum S

__constant__ float c_taps[11];

struct LossDims { int V, H, W; };

__device__ __forceinline__ double block_sum(double v, double *scratch)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < LTHREADS / 32; ++w) t += scratch[w];
    return t; // valid in thread 0
}

__global__ void __launch_bounds__(LTHREADS)
loss_reduce_kernel(LossDims d, const float *__restrict__ rgb, const float *__restrict__ alpha,
                   const float *__restrict__ timg, const float *__restrict__ mask, double *__restrict__ stats)
{
    __shared__ double scratch[LTHREADS / 32];
    const int v = blockIdx.y;
    const size_t npix = (size_t)d.H * d.W;
    const float *a = alpha + v * npix, *m = mask + v * npix, *q = rgb + 3 * v * npix, *t = timg + 3 * v * npix;
    float sI = 0.f, sU = 0.f, sM = 0.f, sL = 0.f;
    if (!rgb && (npix & 3) == 0) { // alpha and mask alone (the SSIM kernels take the L1 sum): 16-byte loads
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *m4 = reinterpret_cast<const float4 *>(m);
        for (size_t i = (size_t)blockIdx.x * LTHREADS + threadIdx.x; i < npix / 4; i += (size_t)gridDim.x * LTHREADS) {
            const float4 av = __ldg(a4 + i), mv = __ldg(m4 + i);
            sI += av.x * mv.x + av.y * mv.y + av.z * mv.z + av.w * mv.w;
            sU += (av.x + mv.x - av.x * mv.x) + (av.y + mv.y - av.y * mv.y) + (av.z + mv.z - av.z * mv.z) + (av.w + mv.w - av.w * mv.w);
            sM += mv.x + mv.y + mv.z + mv.w;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * LTHREADS + threadIdx.x; i < npix; i += (size_t)gridDim.x * LTHREADS) {
            const float av = a[i], mv = m[i];
            sI += av * mv;
            sU += av + mv - av * mv;
            sM += mv;
            if (rgb) sL += fabsf(t[i] - q[3 * i]) + fabsf(t[npix + i] - q[3 * i + 1]) + fabsf(t[2 * npix + i] - q[3 * i + 2]);
        }
    }
    double r;
    r = block_sum((double)sI, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 0, r);
    r = block_sum((double)sU, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 1, r);
    r = block_sum((double)sM, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 2, r);
    if (rgb) { r = block_sum((double)sL, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 3, r); }
}

// ---- packed fp32: Blackwell issues two FMAs per instruction (fma.rn.f32x2 -> FFMA2); a plain three-register FFMA
// only reaches half of the FP32 pipe's rate.  The window passes below are written on pairs: the horizontal pass filters
// two image rows at once (the tile is kept row-pair interleaved in shared memory), the vertical pass two columns.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}

constexpr int LP = LIN / 2;      // row pairs of the 42-row input tile
constexpr int HP_ITEMS = LP * (LT / 4); // horizontal-pass items: a row pair x four output columns
static_assert(LIN % 2 == 0 && LT % 4 == 0 && LTHREADS == 256, "tile geometry");

// 11-tap window over 14 consecutive pairs -> 4 consecutive output pairs
__device__ __forceinline__ void window4(const float2 (&in)[14], float2 (&out)[4])
{
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int tp = 0; tp < 11; ++tp) s = fma2(make_float2(c_taps[tp], c_taps[tp]), in[o + tp], s);
        out[o] = s;
    }
}
// 11-tap window over 12 consecutive pairs -> 2 consecutive output pairs
__device__ __forceinline__ void window2(const float2 (&in)[12], float2 (&out)[2])
{
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int tp = 0; tp < 11; ++tp) s = fma2(make_float2(c_taps[tp], c_taps[tp]), in[o + tp], s);
        out[o] = s;
    }
}

// SSIM forward of one view tile, all three channels, fused with the per-view sums of the other two loss terms
// (soft-IoU I and U, sum of the mask, L1 distance).  p = target image (planar), q = render (interleaved).
template <int MINB>
__global__ void __launch_bounds__(LTHREADS, MINB)
ssim_fwd_kernel(LossDims d, const float *__restrict__ rgb, const float *__restrict__ alpha, const float *__restrict__ timg,
                const float *__restrict__ mask, float coef_over_count, float c1, float c2,
                float *__restrict__ adj /* [V,3,3,H,W]: A1 | A2 | A3 per channel */, double *__restrict__ stats)
{
    __shared__ float2 sp[LP][LIN + 1], sq[LP][LIN + 1]; // [row pair][column]: .x = even row, .y = odd row of the pair
    __shared__ __align__(16) float sh[5][LIN][LT + 4];  // horizontally filtered maps, rows 16-byte aligned
    __shared__ double scratch[LTHREADS / 32];
    const int v = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t npix = (size_t)d.H * d.W;
    const int tid = threadIdx.x;
    float ssum = 0.f, sI = 0.f, sU = 0.f, sM = 0.f, sL = 0.f;
    {   // soft-IoU sums over this tile's own pixels (one 4-pixel row segment per thread)
        const int r = tid >> 3, cg = (tid & 7) * 4;
        const int y = y0 + r;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int x = x0 + cg + o;
            if (y < d.H && x < d.W) {
                const float av = alpha[v * npix + (size_t)y * d.W + x], mv = mask[v * npix + (size_t)y * d.W + x];
                sI += av * mv;
                sU += av + mv - av * mv;
                sM += mv;
            }
        }
    }
    // The (p, q) tile of a channel is fetched into registers while the previous channel is being filtered (the passes are
    // separated by CTA barriers, so a plain load -> store -> barrier sequence would expose the whole DRAM latency per
    // channel) and parked in shared memory once the horizontal pass, the only reader of sp / sq, is done.
    constexpr int NLD = (LIN * LIN + LTHREADS - 1) / LTHREADS; // tile elements per thread
    const float *qv = rgb + 3 * (size_t)v * npix;
    float pre_p[NLD], pre_q[NLD];
    auto fetch = [&](int ch) {
        const float *pch = timg + (3 * (size_t)v + ch) * npix;
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * LTHREADS;
            const int r = i / LIN, c = i - r * LIN;
            const int y = y0 + r - HALO, x = x0 + c - HALO;
            const bool in = i < LIN * LIN && y >= 0 && y < d.H && x >= 0 && x < d.W;
            const size_t pix = (size_t)y * d.W + x;
            pre_p[j] = in ? __ldg(pch + pix) : 0.f;
            pre_q[j] = in ? __ldg(qv + 3 * pix + ch) : 0.f;
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * LTHREADS;
            if (i >= LIN * LIN) continue;
            const int r = i / LIN, c = i - r * LIN;
            *(reinterpret_cast<float *>(&sp[r >> 1][c]) + (r & 1)) = pre_p[j];
            *(reinterpret_cast<float *>(&sq[r >> 1][c]) + (r & 1)) = pre_q[j];
            // L1 term on the tile's own pixels (the halo belongs to the neighbours; outside the image both are 0)
            if (r >= HALO && r < HALO + LT && c >= HALO && c < HALO + LT) sL += fabsf(pre_p[j] - pre_q[j]);
        }
    };
    fetch(0);
    park();
#pragma unroll 1 // three unrolled channels would be 4 k instructions (64 KB of code): keep the body in the instruction cache
    for (int ch = 0; ch < 3; ++ch) {
        __syncthreads(); // sp / sq hold channel ch; the previous channel's vertical pass is done with sh
        if (ch + 1 < 3) fetch(ch + 1);
        for (int item = tid; item < HP_ITEMS; item += LTHREADS) {
            // consecutive lanes take consecutive row pairs (their shared-memory rows fall into different banks)
            const int cgi = item / LP, rp = item - cgi * LP, cg = cgi * 4;
            float2 in2[14], t2[14], out[4];
            auto put = [&](int k, const float2 (&o)[4]) {
                *reinterpret_cast<float4 *>(&sh[k][2 * rp][cg]) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
                *reinterpret_cast<float4 *>(&sh[k][2 * rp + 1][cg]) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
            };
            // three short-lived register windows (p | q | p q) instead of p and q held together: fewer live registers
#pragma unroll
            for (int i = 0; i < 14; ++i) in2[i] = sp[rp][cg + i];
            window4(in2, out); put(0, out);
#pragma unroll
            for (int i = 0; i < 14; ++i) t2[i] = mul2(in2[i], in2[i]);
            window4(t2, out); put(2, out);
#pragma unroll
            for (int i = 0; i < 14; ++i) in2[i] = sq[rp][cg + i];
            window4(in2, out); put(1, out);
#pragma unroll
            for (int i = 0; i < 14; ++i) t2[i] = mul2(in2[i], in2[i]);
            window4(t2, out); put(3, out);
#pragma unroll
            for (int i = 0; i < 14; ++i) t2[i] = mul2(in2[i], sp[rp][cg + i]);
            window4(t2, out); put(4, out);
        }
        __syncthreads();
        if (ch + 1 < 3) park(); // nobody reads sp / sq any more: the next channel moves in under the vertical pass
        {   // vertical pass: thread = (column pair, two output rows)
            const int c = (tid & 15) * 2, r0 = (tid >> 4) * 2;
            float2 res[5][2];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                float2 in[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) in[i] = *reinterpret_cast<const float2 *>(&sh[k][r0 + i][c]);
                window2(in, res[k]);
            }
            float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix, *a2 = a1 + npix, *a3 = a2 + npix;
#pragma unroll
            for (int o = 0; o < 2; ++o) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int y = y0 + r0 + o, x = x0 + c + h;
                    if (y >= d.H || x >= d.W) continue;
                    const bool centre = y >= HALO && y < d.H - HALO && x >= HALO && x < d.W - HALO;
                    float g1 = 0.f, g2 = 0.f, g3 = 0.f;
                    if (centre) {
                        const float mp = h ? res[0][o].y : res[0][o].x, mq = h ? res[1][o].y : res[1][o].x;
                        const float epp = h ? res[2][o].y : res[2][o].x, eqq = h ? res[3][o].y : res[3][o].x;
                        const float epq = h ? res[4][o].y : res[4][o].x;
                        const float vpp = epp - mp * mp, vqq = eqq - mq * mq, vpq = epq - mp * mq;
                        const bool qfree = vqq > 0.f; // clamp(., min = 0) passes the gradient only when not clamped
                        const float N1 = 2.f * mp * mq + c1, N2 = 2.f * vpq + c2;
                        const float D1 = mp * mp + mq * mq + c1, D2 = fmaxf(vpp, 0.f) + fmaxf(vqq, 0.f) + c2;
                        // D1 >= c1, D2 >= c2 > 0: approximate reciprocals (2 ulp) are far inside the 1e-3 gradient tolerance
                        const float i1 = __fdividef(1.0f, D1), i2 = __fdividef(1.0f, D2);
                        const float inv = i1 * i2;
                        const float S = N1 * N2 * inv;
                        ssum += S;
                        const float dD2 = qfree ? -2.f * mq : 0.f;
                        const float dmu = (2.f * mp * N2 - 2.f * mp * N1) * inv - S * (2.f * mq * i1 + dD2 * i2);
                        g1 = coef_over_count * dmu;
                        g2 = qfree ? coef_over_count * (-S * i2) : 0.f;
                        g3 = coef_over_count * 2.f * N1 * inv;
                    }
                    const size_t pix = (size_t)y * d.W + x;
                    a1[pix] = g1; a2[pix] = g2; a3[pix] = g3;
                }
            }
        }
    }
    double r;
    r = block_sum((double)ssum, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 4, r);
    r = block_sum((double)sI, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 0, r);
    r = block_sum((double)sU, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 1, r);
    r = block_sum((double)sM, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 2, r);
    r = block_sum((double)sL, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 3, r);
}

template <int MINB>
__global__ void __launch_bounds__(LTHREADS, MINB)
loss_bwd_kernel(LossDims d, const float *__restrict__ rgb, const float *__restrict__ alpha,
                const float *__restrict__ timg, const float *__restrict__ mask, const float *__restrict__ adj,
                const double *__restrict__ stats, float img_lambda, float *__restrict__ d_rgb, float *__restrict__ d_alpha)
{
    __shared__ float2 sa[3][LP][LIN + 1];              // adjoint maps, row-pair interleaved
    __shared__ __align__(16) float sh[3][LIN][LT + 4];
    const int v = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t npix = (size_t)d.H * d.W;
    const int tid = threadIdx.x;
    const double I = stats[v * NSTAT + 0] + 1e-6, U = stats[v * NSTAT + 1] + 1e-6, msum = stats[v * NSTAT + 2];
    const float l1 = img_lambda == 0.0f ? 0.0f : (float)((double)img_lambda / msum); // lambda 0: no 0 * 0 / 0 for an empty mask
    const float iou_m = (float)(-1.0 / U), iou_c = (float)(I / (U * U)); // d(1 - I/U)/da = -m/U + I (1 - m) / U^2
    const int c = (tid & 15) * 2, r0 = (tid >> 4) * 2; // this thread's 2 x 2 output pixels
    float g[2][2][3];
    constexpr int NLD = (LIN * LIN + LTHREADS - 1) / LTHREADS;
    float pre[3][NLD]; // the next channel's adjoint tiles, in flight while this channel is filtered
    auto fetch = [&](int ch) {
        const float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix;
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * LTHREADS;
            const int r = i / LIN, cc = i - r * LIN;
            const int y = y0 + r - HALO, x = x0 + cc - HALO;
            const bool in = i < LIN * LIN && y >= 0 && y < d.H && x >= 0 && x < d.W;
            const size_t pix = (size_t)y * d.W + x;
#pragma unroll
            for (int k = 0; k < 3; ++k) pre[k][j] = in ? __ldg(a1 + k * npix + pix) : 0.f;
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = tid + j * LTHREADS;
            if (i >= LIN * LIN) continue;
            const int r = i / LIN, cc = i - r * LIN;
#pragma unroll
            for (int k = 0; k < 3; ++k) *(reinterpret_cast<float *>(&sa[k][r >> 1][cc]) + (r & 1)) = pre[k][j];
        }
    };
    fetch(0);
    park();
#pragma unroll 1
    for (int ch = 0; ch < 3; ++ch) {
        __syncthreads();
        if (ch + 1 < 3) fetch(ch + 1);
        for (int item = tid; item < HP_ITEMS; item += LTHREADS) {
            const int cgi = item / LP, rp = item - cgi * LP, cg = cgi * 4;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float2 in[14], out[4];
#pragma unroll
                for (int i = 0; i < 14; ++i) in[i] = sa[k][rp][cg + i];
                window4(in, out);
                *reinterpret_cast<float4 *>(&sh[k][2 * rp][cg]) = make_float4(out[0].x, out[1].x, out[2].x, out[3].x);
                *reinterpret_cast<float4 *>(&sh[k][2 * rp + 1][cg]) = make_float4(out[0].y, out[1].y, out[2].y, out[3].y);
            }
        }
        __syncthreads();
        if (ch + 1 < 3) park();
        float2 res[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float2 in[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) in[i] = *reinterpret_cast<const float2 *>(&sh[k][r0 + i][c]);
            window2(in, res[k]);
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int y = y0 + r0 + o, x = x0 + c + h;
                g[o][h][ch] = 0.f;
                if (y >= d.H || x >= d.W) continue;
                const size_t pix = (size_t)y * d.W + x;
                const float q = rgb[3 * ((size_t)v * npix + pix) + ch], p = timg[(3 * (size_t)v + ch) * npix + pix];
                const float diff = p - q;
                const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                const float w1 = h ? res[0][o].y : res[0][o].x, w2 = h ? res[1][o].y : res[1][o].x, w3 = h ? res[2][o].y : res[2][o].x;
                g[o][h][ch] = w1 + 2.f * q * w2 + p * w3 - l1 * sgn;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int y = y0 + r0 + o, x = x0 + c + h;
            if (y >= d.H || x >= d.W) continue;
            const size_t pix = (size_t)v * npix + (size_t)y * d.W + x;
            d_rgb[3 * pix] = g[o][h][0]; d_rgb[3 * pix + 1] = g[o][h][1]; d_rgb[3 * pix + 2] = g[o][h][2];
            const float m = mask[pix];
            d_alpha[pix] = iou_m * m + iou_c * (1.f - m);
        }
    }
    (void)alpha;
}

// ---- strip-marching SSIM kernels (default) -----------------------------------------------------------------------------
// The tiled kernels above stage a 42 x 42 input tile per 32 x 32 outputs behind CTA barriers and keep every intermediate in
// shared memory: they run at a quarter of the FP32 pipe's rate, stalled on barriers and on the tile loads.  The marching
// kernels turn the work around.  A CTA owns a strip of SW = NT - 16 output columns of one (view, channel) and walks down a
// band of rows, eight rows per step:
//   ring      thread = column (SW + 10 of them): the next eight input rows arrive by 4-byte cp.async in a 24-row ring in
//             shared memory that only the owning thread ever reads -- no barrier, one step of compute between issue and use
//   vertical  thread = column: 18 rows of its own column from the ring -> 8 vertically filtered rows of every map, in
//             registers (packed f32x2: two maps per FFMA2); written once to an exchange buffer
//   horizontal thread = (row, 8 columns): 18-column windows fetched as 16-byte shared loads (rows strided so that the eight
//             rows of a load phase fall into different banks) -> 8 outputs per map, then the pointwise part
//   epilogue  thread = column again (through a small shared transpose): every global access is row-contiguous
// Two barriers per eight rows; global loads and stores are coalesced; no halo is re-read along a strip (1.07x across strips).
// NT is a template parameter: every shared-memory stride is a compile-time constant (the first version, with run-time
// strides, issued more integer than FP32 instructions).
constexpr int RB = 8;              // rows per marching step
constexpr int WIN8 = RB + 2 * HALO; // 18 inputs -> 8 outputs

template <int NT>
struct March {
    static constexpr int SW = NT - 16;   // output columns per strip (multiple of 8, SW + 10 <= NT)
    static constexpr int RS2 = NT + 2;   // float2 per exchange row: == 2 (mod 4): 16-byte rows, phase lanes spread over the banks
    static constexpr int RS1 = NT + 4;   // floats per exchange row: >= SW + 12, == 4 (mod 8)
    static constexpr int OS = SW + 4;    // floats per output-transpose row: == 4 (mod 8)
    static constexpr size_t OB = 3 * (size_t)RB * OS * sizeof(float);
    static constexpr size_t SMEM_FWD = 3 * (size_t)RB * NT * sizeof(float2) + RB * (2 * RS2 * sizeof(float2) + RS1 * sizeof(float)) + OB;
    static constexpr size_t SMEM_BWD = 3 * (size_t)RB * NT * (sizeof(float2) + sizeof(float)) + RB * (RS2 * sizeof(float2) + RS1 * sizeof(float)) + OB;
};

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_l() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_l() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// the taps are symmetric: six distinct values, each duplicated into both halves of a packed operand
__device__ __forceinline__ void load_taps(float2 (&c)[6])
{
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = make_float2(c_taps[k], c_taps[k]);
}
template <int N>
__device__ __forceinline__ void window8(const float2 (&in)[N], float2 (&out)[RB], const float2 (&c)[6])
{
    static_assert(N >= WIN8, "window");
#pragma unroll
    for (int o = 0; o < RB; ++o) {
        float2 s = mul2(c[0], in[o]);
#pragma unroll
        for (int tp = 1; tp < 11; ++tp) s = fma2(c[tp < 6 ? tp : 10 - tp], in[o + tp], s);
        out[o] = s;
    }
}
// one map alone: scalar FMAs.  (Packing output pairs (o, o + 1) into FFMA2 needs a shifted copy of the inputs for the odd
// taps: measured slower -- the extra registers spill in the backward kernel, no gain in the forward.)
template <int N>
__device__ __forceinline__ void window8s(const float (&in)[N], float (&out)[RB], const float2 (&c)[6])
{
    static_assert(N >= WIN8, "window");
#pragma unroll
    for (int o = 0; o < RB; ++o) {
        float s = c[0].x * in[o];
#pragma unroll
        for (int tp = 1; tp < 11; ++tp) s = fmaf(c[tp < 6 ? tp : 10 - tp].x, in[o + tp], s);
        out[o] = s;
    }
}
// 18 consecutive float2 / 20 consecutive floats of an exchange row as 16-byte shared loads
__device__ __forceinline__ void load_row18(const float2 *src2, float2 (&in)[WIN8])
{
    const float4 *src = reinterpret_cast<const float4 *>(src2);
#pragma unroll
    for (int k = 0; k < WIN8 / 2; ++k) {
        const float4 w = src[k];
        in[2 * k] = make_float2(w.x, w.y); in[2 * k + 1] = make_float2(w.z, w.w);
    }
}
__device__ __forceinline__ void load_row20(const float *src1, float (&f)[20])
{
    const float4 *src = reinterpret_cast<const float4 *>(src1);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float4 w = src[k];
        f[4 * k] = w.x; f[4 * k + 1] = w.y; f[4 * k + 2] = w.z; f[4 * k + 3] = w.w;
    }
}
__device__ __forceinline__ void store_row8(float *dst, const float (&g)[RB])
{
    float4 *o = reinterpret_cast<float4 *>(dst);
    o[0] = make_float4(g[0], g[1], g[2], g[3]); o[1] = make_float4(g[4], g[5], g[6], g[7]);
}

__device__ __forceinline__ double block_sum_n(double v, double *scratch)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t; // valid in thread 0
}

// SSIM forward of one (strip, band, view, channel) + the L1 sum of its pixels (the soft-IoU sums: loss_reduce_kernel).
template <int NT>
__global__ void __launch_bounds__(NT, NT <= 160 ? 3 : 2)
ssim_march_kernel(LossDims d, int band, const float *__restrict__ rgb, const float *__restrict__ timg, float coef_over_count,
                  float c1, float c2, float *__restrict__ adj, double *__restrict__ stats)
{
    using G = March<NT>;
    extern __shared__ __align__(16) unsigned char march_smem[];
    __shared__ double scratch[8];
    float2 *ring = reinterpret_cast<float2 *>(march_smem);      // [3 * RB][NT]  (p, q) input rows, thread-private columns
    float2 *ex01 = ring + 3 * RB * NT;                           // [RB][RS2]     vertically filtered (p, q)
    float2 *ex23 = ex01 + RB * G::RS2;                           // [RB][RS2]     ... (pp, qq)
    float *ex4 = reinterpret_cast<float *>(ex23 + RB * G::RS2);  // [RB][RS1]     ... pq
    float *ob = ex4 + RB * G::RS1;                               // [3][RB][OS]   the three adjoints on their way out
    const int t = threadIdx.x;
    const int ch = blockIdx.x % 3, strip = blockIdx.x / 3, v = blockIdx.z;
    const int x0 = strip * G::SW, yb = blockIdx.y * band, yend = min(yb + band, d.H);
    const int nb = (yend - yb + RB - 1) / RB;
    const size_t npix = (size_t)d.H * d.W;
    const int xc = x0 - HALO + t; // vertical pass: this thread's column
    const bool col_on = t < G::SW + 2 * HALO;
    const bool col_in = col_on && xc >= 0 && xc < d.W;
    const bool col_own = t >= HALO && t < HALO + G::SW && xc < d.W;
    const float *pch = timg + (3 * (size_t)v + ch) * npix + xc; // target image, planar: this thread's column
    const float *qch = rgb + 3 * ((size_t)v * npix + xc) + ch;  // render, interleaved
    float2 c[6];
    load_taps(c);
    if (col_on && !col_in) { // a column outside the image stays zero for the whole walk
#pragma unroll
        for (int k = 0; k < 3 * RB; ++k) ring[k * NT + t] = make_float2(0.f, 0.f);
    }
    // unit u = input rows yb + RB (u - 1) .. + RB - 1; step i filters units i, i + 1, i + 2
    auto stage = [&](int u, int slot) {
        if (col_in && u <= nb + 1) {
            float2 *dst = ring + slot * RB * NT + t;
            const int y0 = yb + RB * (u - 1);
            if (y0 >= 0 && y0 + RB <= d.H) {
                const float *pp = pch + (size_t)y0 * d.W, *qq = qch + 3 * (size_t)y0 * d.W;
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    cp_async4(&dst[k * NT].x, pp);
                    cp_async4(&dst[k * NT].y, qq);
                    pp += d.W; qq += 3 * d.W;
                }
            } else { // a unit across the top / bottom edge of the image: rows outside are zero
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const int y = y0 + k;
                    if (y >= 0 && y < d.H) {
                        cp_async4(&dst[k * NT].x, pch + (size_t)y * d.W);
                        cp_async4(&dst[k * NT].y, qch + 3 * (size_t)y * d.W);
                    } else {
                        dst[k * NT] = make_float2(0.f, 0.f);
                    }
                }
            }
        }
        cp_async_commit_l();
    };
    stage(0, 0); stage(1, 1); stage(2, 2);
    float ssum = 0.f, sL = 0.f;
    int s0 = 0, s1 = 1, s2 = 2; // ring slots of units i, i + 1, i + 2
    for (int i = 0; i < nb; ++i) {
        const int yo = yb + RB * i;
        cp_async_wait_all_l();
        float2 win[WIN8]; // rows yo - 5 .. yo + 12 of this thread's column
        if (col_on) {
            const float2 *u0 = ring + s0 * RB * NT + t, *u1 = ring + s1 * RB * NT + t, *u2 = ring + s2 * RB * NT + t;
#pragma unroll
            for (int k = 0; k < HALO; ++k) win[k] = u0[(RB - HALO + k) * NT];
#pragma unroll
            for (int k = 0; k < RB; ++k) win[HALO + k] = u1[k * NT];
#pragma unroll
            for (int k = 0; k < HALO; ++k) win[HALO + RB + k] = u2[k * NT];
        }
        stage(i + 3, s0); // into the slot of unit i, whose rows are in registers now
        { const int sx = s0; s0 = s1; s1 = s2; s2 = sx; }
        if (col_on) {
            if (col_own) {
#pragma unroll
                for (int k = 0; k < RB; ++k)
                    if (yo + k < yend) sL += fabsf(win[HALO + k].x - win[HALO + k].y);
            }
            float2 out[RB];
            window8(win, out, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex01[r * G::RS2 + t] = out[r];
            {
                float2 sq[WIN8];
#pragma unroll
                for (int k = 0; k < WIN8; ++k) sq[k] = mul2(win[k], win[k]);
                window8(sq, out, c);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) ex23[r * G::RS2 + t] = out[r];
            float pq[WIN8], o4[RB];
#pragma unroll
            for (int k = 0; k < WIN8; ++k) pq[k] = win[k].x * win[k].y;
            window8s(pq, o4, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex4[r * G::RS1 + t] = o4[r];
        }
        __syncthreads();
        if (t < G::SW) { // horizontal pass + SSIM: thread = (row, eight columns)
            const int row = t & 7, cg = t >> 3;
            const int y = yo + row;
            float2 in[WIN8], m[RB], e[RB];
            float epq[RB];
            load_row18(ex01 + row * G::RS2 + 8 * cg, in);
            window8(in, m, c);
            load_row18(ex23 + row * G::RS2 + 8 * cg, in);
            window8(in, e, c);
            {
                float f[20];
                load_row20(ex4 + row * G::RS1 + 8 * cg, f);
                window8s(f, epq, c);
            }
            float g1[RB], g2[RB], g3[RB];
            const bool row_ok = y >= HALO && y < d.H - HALO && y < yend;
#pragma unroll
            for (int o = 0; o < RB; ++o) {
                const int x = x0 + 8 * cg + o;
                g1[o] = 0.f; g2[o] = 0.f; g3[o] = 0.f;
                if (row_ok && x >= HALO && x < d.W - HALO) { // windows entirely inside the image only
                    const float mp = m[o].x, mq = m[o].y, epp = e[o].x, eqq = e[o].y;
                    const float vpp = epp - mp * mp, vqq = eqq - mq * mq, vpq = epq[o] - mp * mq;
                    const bool qfree = vqq > 0.f; // clamp(., min = 0) passes the gradient only when not clamped
                    const float N1 = 2.f * mp * mq + c1, N2 = 2.f * vpq + c2;
                    const float D1 = mp * mp + mq * mq + c1, D2 = fmaxf(vpp, 0.f) + fmaxf(vqq, 0.f) + c2;
                    // D1 >= c1, D2 >= c2 > 0: approximate reciprocals (2 ulp) are far inside the 1e-3 gradient tolerance
                    const float i1 = __fdividef(1.0f, D1), i2 = __fdividef(1.0f, D2);
                    const float inv = i1 * i2;
                    const float S = N1 * N2 * inv;
                    ssum += S;
                    const float dD2 = qfree ? -2.f * mq : 0.f;
                    const float dmu = (2.f * mp * N2 - 2.f * mp * N1) * inv - S * (2.f * mq * i1 + dD2 * i2);
                    g1[o] = coef_over_count * dmu;
                    g2[o] = qfree ? coef_over_count * (-S * i2) : 0.f;
                    g3[o] = coef_over_count * 2.f * N1 * inv;
                }
            }
            store_row8(ob + (0 * RB + row) * G::OS + 8 * cg, g1);
            store_row8(ob + (1 * RB + row) * G::OS + 8 * cg, g2);
            store_row8(ob + (2 * RB + row) * G::OS + 8 * cg, g3);
        }
        __syncthreads();
        if (t < G::SW && x0 + t < d.W) { // epilogue: thread = column, row-contiguous stores
            float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix + (size_t)yo * d.W + x0 + t;
            const int nr = min(RB, yend - yo);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r < nr) {
                    a1[0] = ob[(0 * RB + r) * G::OS + t];
                    a1[npix] = ob[(1 * RB + r) * G::OS + t];
                    a1[2 * npix] = ob[(2 * RB + r) * G::OS + t];
                }
                a1 += d.W;
            }
        }
    }
    double r;
    r = block_sum_n((double)ssum, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 4, r);
    r = block_sum_n((double)sL, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 3, r);
}

// The adjoints filtered back (same window: the taps are symmetric) and combined with the L1 sign term -> d_rgb; same
// marching structure, maps (A1, A2) packed + A3.  (d_alpha: iou_bwd_kernel.)
template <int NT>
__global__ void __launch_bounds__(NT, NT <= 160 ? 3 : 2)
loss_bwd_march_kernel(LossDims d, int band, const float *__restrict__ rgb, const float *__restrict__ timg,
                      const float *__restrict__ adj, const double *__restrict__ stats, float img_lambda, float *__restrict__ d_rgb)
{
    using G = March<NT>;
    extern __shared__ __align__(16) unsigned char march_smem[];
    float2 *ring = reinterpret_cast<float2 *>(march_smem);             // [3 * RB][NT]  (A1, A2)
    float *ring3 = reinterpret_cast<float *>(ring + 3 * RB * NT);      // [3 * RB][NT]  A3
    float2 *ex01 = reinterpret_cast<float2 *>(ring3 + 3 * RB * NT);    // [RB][RS2]
    float *ex2 = reinterpret_cast<float *>(ex01 + RB * G::RS2);        // [RB][RS1]
    float *ob = ex2 + RB * G::RS1;                                     // [3][RB][OS]
    const int t = threadIdx.x;
    const int ch = blockIdx.x % 3, strip = blockIdx.x / 3, v = blockIdx.z;
    const int x0 = strip * G::SW, yb = blockIdx.y * band, yend = min(yb + band, d.H);
    const int nb = (yend - yb + RB - 1) / RB;
    const size_t npix = (size_t)d.H * d.W;
    const int xc = x0 - HALO + t;
    const bool col_on = t < G::SW + 2 * HALO;
    const bool col_in = col_on && xc >= 0 && xc < d.W;
    const bool epi_on = t < G::SW && x0 + t < d.W;
    const float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix + xc;
    const double msum = stats[v * NSTAT + 2];
    const float l1 = img_lambda == 0.0f ? 0.0f : (float)((double)img_lambda / msum); // lambda 0: no 0 * 0 / 0 for an empty mask
    float2 c[6];
    load_taps(c);
    if (col_on && !col_in) {
#pragma unroll
        for (int k = 0; k < 3 * RB; ++k) { ring[k * NT + t] = make_float2(0.f, 0.f); ring3[k * NT + t] = 0.f; }
    }
    auto stage = [&](int u, int slot) {
        if (col_in && u <= nb + 1) {
            float2 *dst = ring + slot * RB * NT + t;
            float *dst3 = ring3 + slot * RB * NT + t;
            const int y0 = yb + RB * (u - 1);
            if (y0 >= 0 && y0 + RB <= d.H) {
                const float *pa = a1 + (size_t)y0 * d.W;
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    cp_async4(&dst[k * NT].x, pa);
                    cp_async4(&dst[k * NT].y, pa + npix);
                    cp_async4(&dst3[k * NT], pa + 2 * npix);
                    pa += d.W;
                }
            } else {
#pragma unroll
                for (int k = 0; k < RB; ++k) {
                    const int y = y0 + k;
                    if (y >= 0 && y < d.H) {
                        const float *pa = a1 + (size_t)y * d.W;
                        cp_async4(&dst[k * NT].x, pa);
                        cp_async4(&dst[k * NT].y, pa + npix);
                        cp_async4(&dst3[k * NT], pa + 2 * npix);
                    } else {
                        dst[k * NT] = make_float2(0.f, 0.f);
                        dst3[k * NT] = 0.f;
                    }
                }
            }
        }
        cp_async_commit_l();
    };
    stage(0, 0); stage(1, 1); stage(2, 2);
    int s0 = 0, s1 = 1, s2 = 2;
    for (int i = 0; i < nb; ++i) {
        const int yo = yb + RB * i;
        cp_async_wait_all_l();
        float2 win[WIN8];
        float win3[WIN8];
        if (col_on) {
            const int b0 = s0 * RB * NT + t, b1 = s1 * RB * NT + t, b2 = s2 * RB * NT + t;
#pragma unroll
            for (int k = 0; k < HALO; ++k) { win[k] = ring[b0 + (RB - HALO + k) * NT]; win3[k] = ring3[b0 + (RB - HALO + k) * NT]; }
#pragma unroll
            for (int k = 0; k < RB; ++k) { win[HALO + k] = ring[b1 + k * NT]; win3[HALO + k] = ring3[b1 + k * NT]; }
#pragma unroll
            for (int k = 0; k < HALO; ++k) { win[HALO + RB + k] = ring[b2 + k * NT]; win3[HALO + RB + k] = ring3[b2 + k * NT]; }
        }
        stage(i + 3, s0);
        { const int sx = s0; s0 = s1; s1 = s2; s2 = sx; }
        if (col_on) {
            float2 out[RB];
            float o3[RB];
            window8(win, out, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex01[r * G::RS2 + t] = out[r];
            window8s(win3, o3, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex2[r * G::RS1 + t] = o3[r];
        }
        __syncthreads();
        // the epilogue's own inputs (render and target of this thread's column, eight rows) are fetched here: their
        // latency hides behind the horizontal pass instead of stalling the epilogue
        float eq[RB], ep[RB];
        const int nr = min(RB, yend - yo);
        if (epi_on) {
            const float *qg = rgb + 3 * ((size_t)v * npix + (size_t)yo * d.W + x0 + t) + ch;
            const float *pg = timg + (3 * (size_t)v + ch) * npix + (size_t)yo * d.W + x0 + t;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                eq[r] = r < nr ? __ldg(qg) : 0.f;
                ep[r] = r < nr ? __ldg(pg) : 0.f;
                qg += 3 * d.W; pg += d.W;
            }
        }
        if (t < G::SW) {
            const int row = t & 7, cg = t >> 3;
            float2 in[WIN8], w12[RB];
            float w3[RB];
            load_row18(ex01 + row * G::RS2 + 8 * cg, in);
            window8(in, w12, c);
            {
                float f[20];
                load_row20(ex2 + row * G::RS1 + 8 * cg, f);
                window8s(f, w3, c);
            }
            float w1[RB], w2[RB];
#pragma unroll
            for (int o = 0; o < RB; ++o) { w1[o] = w12[o].x; w2[o] = w12[o].y; }
            store_row8(ob + (0 * RB + row) * G::OS + 8 * cg, w1);
            store_row8(ob + (1 * RB + row) * G::OS + 8 * cg, w2);
            store_row8(ob + (2 * RB + row) * G::OS + 8 * cg, w3);
        }
        __syncthreads();
        if (epi_on) {
            float *dg = d_rgb + 3 * ((size_t)v * npix + (size_t)yo * d.W + x0 + t) + ch;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r < nr) {
                    const float q = eq[r], p = ep[r];
                    const float diff = p - q;
                    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                    const float w1 = ob[(0 * RB + r) * G::OS + t], w2 = ob[(1 * RB + r) * G::OS + t], w3 = ob[(2 * RB + r) * G::OS + t];
                    *dg = w1 + 2.f * q * w2 + p * w3 - l1 * sgn;
                }
                dg += 3 * d.W;
            }
        }
    }
}

// strip width of the marching kernels for an image width: the instantiation with the fewest threads over all strips
int march_nt(int W)
{
    const int cand[4] = { 64, 96, 160, 224 };
    int best = 160;
    long best_cost = -1;
    for (int k = 0; k < 4; ++k) {
        const int sw = cand[k] - 16;
        const long cost = (long)((W + sw - 1) / sw) * cand[k];
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = cand[k]; }
    }
    return best;
}

template <int NT>
int launch_march(const LossDims &d, int band, dim3 grid, const float *rgb, const float *timg, float coef, float c1, float c2,
                 float *adj, double *stats, float img_lambda, float *d_rgb, cudaStream_t s)
{
    using G = March<NT>;
    static bool attr_ready[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = dev < 64 ? dev : 63;
    if (!attr_ready[di] || dev >= 64) { // above 48 KB of dynamic shared memory is opt-in, per device
        if (cudaFuncSetAttribute(ssim_march_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_FWD) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(loss_bwd_march_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BWD) != cudaSuccess) return -1;
        attr_ready[di] = true;
    }
    ssim_march_kernel<NT><<<grid, NT, G::SMEM_FWD, s>>>(d, band, rgb, timg, coef, c1, c2, adj, stats);
    if (!d_rgb) return 1;
    loss_bwd_march_kernel<NT><<<grid, NT, G::SMEM_BWD, s>>>(d, band, rgb, timg, adj, stats, img_lambda, d_rgb);
    return 2;
}

__global__ void loss_finalize_kernel(LossDims d, const double *__restrict__ stats, float ssim_lambda, float img_lambda,
                                     float *__restrict__ losses)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= d.V) return;
    const double *s = stats + v * NSTAT;
    const double count = 3.0 * (double)(d.H - 2 * HALO) * (double)(d.W - 2 * HALO);
    losses[3 * v + 0] = (float)(1.0 - (s[0] + 1e-6) / (s[1] + 1e-6));
    // a term whose weight is zero is exactly zero (an empty target mask would otherwise give 0 * 0 / 0 = NaN)
    losses[3 * v + 1] = ssim_lambda == 0.0f ? 0.0f : (float)((double)ssim_lambda * (1.0 - s[4] / count));
    losses[3 * v + 2] = img_lambda == 0.0f ? 0.0f : (float)((double)img_lambda * s[3] / s[2]);
}

// soft IoU alone (scripts/training/train_script.py:30-36): losses [V], d_alpha [V,H,W] from the sums of loss_reduce
__global__ void __launch_bounds__(LTHREADS)
iou_bwd_kernel(LossDims d, const float *__restrict__ mask, const double *__restrict__ stats, float *__restrict__ losses,
               float *__restrict__ d_alpha)
{
    const int v = blockIdx.y;
    const size_t npix = (size_t)d.H * d.W;
    const double I = stats[v * NSTAT + 0] + 1e-6, U = stats[v * NSTAT + 1] + 1e-6;
    if (losses && blockIdx.x == 0 && threadIdx.x == 0) losses[v] = (float)(1.0 - I / U);
    if (!d_alpha) return;
    const float iou_m = (float)(-1.0 / U), iou_c = (float)(I / (U * U));
    if ((npix & 3) == 0) {
        const float4 *m4 = reinterpret_cast<const float4 *>(mask + v * npix);
        float4 *o4 = reinterpret_cast<float4 *>(d_alpha + v * npix);
        for (size_t i = (size_t)blockIdx.x * LTHREADS + threadIdx.x; i < npix / 4; i += (size_t)gridDim.x * LTHREADS) {
            const float4 m = __ldg(m4 + i);
            o4[i] = make_float4(iou_m * m.x + iou_c * (1.f - m.x), iou_m * m.y + iou_c * (1.f - m.y),
                                iou_m * m.z + iou_c * (1.f - m.z), iou_m * m.w + iou_c * (1.f - m.w));
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * LTHREADS + threadIdx.x; i < npix; i += (size_t)gridDim.x * LTHREADS) {
        const float m = mask[v * npix + i];
        d_alpha[v * npix + i] = iou_m * m + iou_c * (1.f - m);
    }
}

} // namespace

// stats: [V * 8] doubles, adj: [V * 9 * H * W] floats of scratch.  Returns launches or -1.
int ps_launch_view_loss(int V, int H, int W, const float *rgb, const float *alpha, const float *timg, const float *mask,
                        float ssim_lambda, float img_lambda, double *stats, float *adj, float *losses, float *d_rgb,
                        float *d_alpha, cudaStream_t s)
{
    static bool taps_ready[64] = {}; // __constant__ memory is per device
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = dev < 64 ? dev : 63;
    if (!taps_ready[di] || dev >= 64) { // exp(-(d / 1.5)^2 / 2), d = -5 .. 5, normalised (torchmetrics _gaussian)
        double g[11], sum = 0.0;
        for (int i = 0; i < 11; ++i) { const double dd = (double)(i - 5) / 1.5; g[i] = exp(-dd * dd / 2.0); sum += g[i]; }
        float gf[11];
        for (int i = 0; i < 11; ++i) gf[i] = (float)(g[i] / sum);
        if (cudaMemcpyToSymbolAsync(c_taps, gf, sizeof(gf), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(s) != cudaSuccess) return -1; // gf lives on this stack frame
        taps_ready[di] = true;
    }
    const LossDims d = { V, H, W };
    if (cudaMemsetAsync(stats, 0, (size_t)V * NSTAT * sizeof(double), s) != cudaSuccess) return -1;
    const size_t npix = (size_t)H * W;
    int bx = (int)((npix + LTHREADS * 8 - 1) / (LTHREADS * 8));
    bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
    const dim3 tiles((W + LT - 1) / LT, (H + LT - 1) / LT, V);
    const double count = 3.0 * (double)(H - 2 * HALO) * (double)(W - 2 * HALO);
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f; // (k * data_range)^2, data_range = 1.0
    int n = 1;
    static const int minb = getenv("PS_LOSS_MINB") ? atoi(getenv("PS_LOSS_MINB")) : 2; // A/B switch: registers bounded for 2 / 3 CTAs per SM
    const float coef = d_rgb ? (float)(-(double)ssim_lambda / count) : 0.0f;
    static const bool tiled = getenv("PS_LOSS_TILED") != nullptr; // A/B switch: the 32 x 32 tile kernels
    if (!tiled) {
        static const int band_env = getenv("PS_LOSS_BAND") ? atoi(getenv("PS_LOSS_BAND")) : 128;
        const int Hr = (H + RB - 1) / RB * RB;
        int band = band_env >= RB ? band_env / RB * RB : 128;
        band = band < Hr ? band : Hr;
        static const int nt_env = getenv("PS_LOSS_NT") ? atoi(getenv("PS_LOSS_NT")) : 0; // A/B switch: strip width
        const int nt = (nt_env == 64 || nt_env == 96 || nt_env == 160 || nt_env == 224) ? nt_env : march_nt(W);
        const int sw = nt - 16;
        const dim3 grid(3 * ((W + sw - 1) / sw), (H + band - 1) / band, V);
        // soft-IoU sums of alpha and mask: one streaming pass of their own (0.9 GB at c2)
        loss_reduce_kernel<<<dim3(bx, V), LTHREADS, 0, s>>>(d, nullptr, alpha, nullptr, mask, stats);
        n += 1;
        int m = -1;
        switch (nt) {
        case 64: m = launch_march<64>(d, band, grid, rgb, timg, coef, c1, c2, adj, stats, img_lambda, d_rgb, s); break;
        case 96: m = launch_march<96>(d, band, grid, rgb, timg, coef, c1, c2, adj, stats, img_lambda, d_rgb, s); break;
        case 224: m = launch_march<224>(d, band, grid, rgb, timg, coef, c1, c2, adj, stats, img_lambda, d_rgb, s); break;
        default: m = launch_march<160>(d, band, grid, rgb, timg, coef, c1, c2, adj, stats, img_lambda, d_rgb, s); break;
        }
        if (m < 0) return -1;
        n += m;
        if (d_rgb) {
            iou_bwd_kernel<<<dim3(bx, V), LTHREADS, 0, s>>>(d, mask, stats, nullptr, d_alpha);
            n += 1;
        }
        loss_finalize_kernel<<<(V + 127) / 128, 128, 0, s>>>(d, stats, ssim_lambda, img_lambda, losses);
        return cudaGetLastError() == cudaSuccess ? n : -1;
    }
    if (minb == 3) ssim_fwd_kernel<3><<<tiles, LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, coef, c1, c2, adj, stats);
    else ssim_fwd_kernel<2><<<tiles, LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, coef, c1, c2, adj, stats);
    n += 1;
    if (d_rgb) {
        if (minb == 3) loss_bwd_kernel<3><<<tiles, LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, adj, stats, img_lambda, d_rgb, d_alpha);
        else loss_bwd_kernel<2><<<tiles, LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, adj, stats, img_lambda, d_rgb, d_alpha);
        n += 1;
    }
    loss_finalize_kernel<<<(V + 127) / 128, 128, 0, s>>>(d, stats, ssim_lambda, img_lambda, losses);
    return cudaGetLastError() == cudaSuccess ? n : -1;
}

// stats: [V * 8] doubles of scratch.  Any image size (no SSIM window).
int ps_launch_iou_loss(int V, int H, int W, const float *alpha, const float *mask, double *stats, float *losses, float *d_alpha,
                       cudaStream_t s)
{
    const LossDims d = { V, H, W };
    if (cudaMemsetAsync(stats, 0, (size_t)V * NSTAT * sizeof(double), s) != cudaSuccess) return -1;
    const size_t npix = (size_t)H * W;
    int bx = (int)((npix + LTHREADS * 8 - 1) / (LTHREADS * 8));
    bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
    loss_reduce_kernel<<<dim3(bx, V), LTHREADS, 0, s>>>(d, nullptr, alpha, nullptr, mask, stats);
    iou_bwd_kernel<<<dim3(bx, V), LTHREADS, 0, s>>>(d, mask, stats, losses, d_alpha);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}
