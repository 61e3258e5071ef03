// ps_loss.cu -- the per-view training loss that follows render() every step, fused with its own backward
// (SURVEY.md 8f-f1): soft IoU(alpha, mask) + img_lambda * L1(rgb, img) / sum(mask) + ssim_lambda * (1 - SSIM),
// producing the three loss terms and d_rgb / d_alpha -- the cotangents ps_backward takes -- in one call.
// Replaces scripts/training/train_script.py:30-36 (get_iou_loss), :129-133 (three torch / torchmetrics graphs and
// their autograd) of the reference.  SSIM = torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0) (:270),
// called as ssim(target_img, rgb): 11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03, windows entirely inside the
// image only (its reflection padding is cropped away again), variances clamped at 0.  torchmetrics is not in the
// reference tree: the algorithm is restated from its published source (oracle/loss_ref.py, parity unpinned).
//
// Four launches per call, all views batched:
//   loss_reduce   per-view sums  I = sum a m, U = sum (a + m - a m), sum m, sum |t - rgb|       [HBM, one read]
//   ssim_fwd      32x32-pixel tiles, separable 11-tap windows of (p, q, pp, qq, pq) in shared memory; per window
//                 centre S and the three adjoints dS/d(mu_q), dS/d(E qq), dS/d(E pq) (scaled by -lambda / count)
//   loss_bwd      the adjoints filtered back with the same window (its transpose: the taps are symmetric), combined
//                 with the L1 sign term -> d_rgb [V,H,W,3]; IoU quotient rule -> d_alpha [V,H,W]
//   loss_finalize losses [V,3] = (iou, ssim, img)
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
    for (size_t i = (size_t)blockIdx.x * LTHREADS + threadIdx.x; i < npix; i += (size_t)gridDim.x * LTHREADS) {
        const float av = a[i], mv = m[i];
        sI += av * mv;
        sU += av + mv - av * mv;
        sM += mv;
        if (rgb) sL += fabsf(t[i] - q[3 * i]) + fabsf(t[npix + i] - q[3 * i + 1]) + fabsf(t[2 * npix + i] - q[3 * i + 2]);
    }
    double r;
    r = block_sum((double)sI, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 0, r);
    r = block_sum((double)sU, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 1, r);
    r = block_sum((double)sM, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 2, r);
    r = block_sum((double)sL, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 3, r);
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
// kernels turn the work around.  A CTA owns a strip of SW output columns of one (view, channel) and walks down a band of
// rows, eight rows per step:
//   ring      thread = column (SW + 10 of them): the next eight input rows arrive by 4-byte cp.async (zero-filled outside
//             the image) in a 24-row ring in shared memory that only the owning thread ever reads -- no barrier, one step of
//             compute between issue and use
//   vertical  thread = column: 18 rows of its own column from the ring -> 8 vertically filtered rows of every map, in
//             registers (packed f32x2: two maps per FFMA2); written once to an exchange buffer
//   horizontal thread = (row, 8 columns): 18-column windows fetched as 16-byte shared loads (rows strided so that the eight
//             rows of a load phase fall into different banks) -> 8 outputs per map, then the pointwise part
//   epilogue  thread = column again (through a small shared transpose): every global access is row-contiguous
// Two barriers per eight rows; global loads and stores are coalesced; no halo is re-read along a strip (1.07x across strips).
constexpr int RB = 8;              // rows per marching step
constexpr int WIN8 = RB + 2 * HALO; // 18 inputs -> 8 outputs

struct MarchGeom { int SW, NT, RS2, RS1, OS, B; };

__device__ __forceinline__ void cp_async4_z(void *smem, const void *gmem, bool ok)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(ok ? 4 : 0) : "memory");
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

// SSIM forward of one (strip, band, view, channel); channel-0 CTAs also take the soft-IoU sums of their pixels.
__global__ void __launch_bounds__(256, 2)
ssim_march_kernel(LossDims d, MarchGeom mg, const float *__restrict__ rgb, const float *__restrict__ alpha,
                  const float *__restrict__ timg, const float *__restrict__ mask, float coef_over_count, float c1, float c2,
                  float *__restrict__ adj, double *__restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char march_smem[];
    __shared__ double scratch[8];
    float2 *ring = reinterpret_cast<float2 *>(march_smem);      // [3 * RB][NT]  (p, q) input rows, thread-private columns
    float2 *ex01 = ring + 3 * RB * mg.NT;                        // [RB][RS2]     vertically filtered (p, q)
    float2 *ex23 = ex01 + RB * mg.RS2;                           // [RB][RS2]     ... (pp, qq)
    float *ex4 = reinterpret_cast<float *>(ex23 + RB * mg.RS2);  // [RB][RS1]     ... pq
    float *ob = ex4 + RB * mg.RS1;                               // [3][RB][OS]   the three adjoints on their way out
    const int t = threadIdx.x, NT = mg.NT;
    const int ch = blockIdx.x % 3, strip = blockIdx.x / 3, v = blockIdx.z;
    const int x0 = strip * mg.SW, yb = blockIdx.y * mg.B, yend = min(yb + mg.B, d.H);
    const int nb = (yend - yb + RB - 1) / RB;
    const size_t npix = (size_t)d.H * d.W;
    const int xc = x0 - HALO + t; // vertical pass: this thread's column
    const bool col_on = t < mg.SW + 2 * HALO;
    const bool col_in = col_on && xc >= 0 && xc < d.W;
    const bool col_own = t >= HALO && t < HALO + mg.SW && xc < d.W;
    const float *pch = timg + (3 * (size_t)v + ch) * npix; // target image, planar
    const float *qch = rgb + 3 * (size_t)v * npix + ch;    // render, interleaved
    float2 c[6];
    load_taps(c);
    // unit u = input rows yb + RB (u - 1) .. + RB - 1; step i filters units i, i + 1, i + 2
    auto stage = [&](int u) {
        if (col_on && u <= nb + 1) {
            float2 *dst = ring + (size_t)((u % 3) * RB) * NT + t;
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int y = yb + RB * (u - 1) + k;
                const bool ok = col_in && y >= 0 && y < d.H;
                const size_t pix = ok ? (size_t)y * d.W + xc : 0;
                cp_async4_z(&dst[k * NT].x, pch + pix, ok);
                cp_async4_z(&dst[k * NT].y, qch + 3 * pix, ok);
            }
        }
        cp_async_commit_l();
    };
    stage(0); stage(1); stage(2);
    float ssum = 0.f, sL = 0.f, sI = 0.f, sU = 0.f, sM = 0.f;
    for (int i = 0; i < nb; ++i) {
        const int yo = yb + RB * i;
        cp_async_wait_all_l();
        float2 win[WIN8]; // rows yo - 5 .. yo + 12 of this thread's column
        if (col_on) {
            const float2 *u0 = ring + (size_t)((i % 3) * RB) * NT + t, *u1 = ring + (size_t)(((i + 1) % 3) * RB) * NT + t,
                         *u2 = ring + (size_t)(((i + 2) % 3) * RB) * NT + t;
#pragma unroll
            for (int k = 0; k < HALO; ++k) win[k] = u0[(RB - HALO + k) * NT];
#pragma unroll
            for (int k = 0; k < RB; ++k) win[HALO + k] = u1[k * NT];
#pragma unroll
            for (int k = 0; k < HALO; ++k) win[HALO + RB + k] = u2[k * NT];
        }
        stage(i + 3); // into the slots of unit i, whose rows are in registers now
        if (col_on) {
            if (col_own) {
#pragma unroll
                for (int k = 0; k < RB; ++k)
                    if (yo + k < yend) sL += fabsf(win[HALO + k].x - win[HALO + k].y);
            }
            float2 out[RB];
            window8(win, out, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex01[r * mg.RS2 + t] = out[r];
            {
                float2 sq[WIN8];
#pragma unroll
                for (int k = 0; k < WIN8; ++k) sq[k] = mul2(win[k], win[k]);
                window8(sq, out, c);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) ex23[r * mg.RS2 + t] = out[r];
            float pq[WIN8], o4[RB];
#pragma unroll
            for (int k = 0; k < WIN8; ++k) pq[k] = win[k].x * win[k].y;
            window8s(pq, o4, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex4[r * mg.RS1 + t] = o4[r];
        }
        __syncthreads();
        if (t < mg.SW) { // horizontal pass + SSIM: thread = (row, eight columns)
            const int row = t & 7, cg = t >> 3;
            const int y = yo + row;
            float2 in[WIN8], m[RB], e[RB];
            float epq[RB];
            {
                const float4 *src = reinterpret_cast<const float4 *>(ex01 + row * mg.RS2 + 8 * cg);
#pragma unroll
                for (int k = 0; k < WIN8 / 2; ++k) {
                    const float4 w = src[k];
                    in[2 * k] = make_float2(w.x, w.y); in[2 * k + 1] = make_float2(w.z, w.w);
                }
                window8(in, m, c);
            }
            {
                const float4 *src = reinterpret_cast<const float4 *>(ex23 + row * mg.RS2 + 8 * cg);
#pragma unroll
                for (int k = 0; k < WIN8 / 2; ++k) {
                    const float4 w = src[k];
                    in[2 * k] = make_float2(w.x, w.y); in[2 * k + 1] = make_float2(w.z, w.w);
                }
                window8(in, e, c);
            }
            {
                float f[20];
                const float4 *src = reinterpret_cast<const float4 *>(ex4 + row * mg.RS1 + 8 * cg);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 w = src[k];
                    f[4 * k] = w.x; f[4 * k + 1] = w.y; f[4 * k + 2] = w.z; f[4 * k + 3] = w.w;
                }
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
            float4 *o1 = reinterpret_cast<float4 *>(ob + (0 * RB + row) * mg.OS + 8 * cg);
            float4 *o2 = reinterpret_cast<float4 *>(ob + (1 * RB + row) * mg.OS + 8 * cg);
            float4 *o3 = reinterpret_cast<float4 *>(ob + (2 * RB + row) * mg.OS + 8 * cg);
            o1[0] = make_float4(g1[0], g1[1], g1[2], g1[3]); o1[1] = make_float4(g1[4], g1[5], g1[6], g1[7]);
            o2[0] = make_float4(g2[0], g2[1], g2[2], g2[3]); o2[1] = make_float4(g2[4], g2[5], g2[6], g2[7]);
            o3[0] = make_float4(g3[0], g3[1], g3[2], g3[3]); o3[1] = make_float4(g3[4], g3[5], g3[6], g3[7]);
        }
        __syncthreads();
        if (t < mg.SW && x0 + t < d.W) { // epilogue: thread = column, row-contiguous stores
            const int x = x0 + t;
            float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix, *a2 = a1 + npix, *a3 = a2 + npix;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int y = yo + r;
                if (y >= yend) break;
                const size_t pix = (size_t)y * d.W + x;
                a1[pix] = ob[(0 * RB + r) * mg.OS + t];
                a2[pix] = ob[(1 * RB + r) * mg.OS + t];
                a3[pix] = ob[(2 * RB + r) * mg.OS + t];
                if (ch == 0) {
                    const float av = alpha[v * npix + pix], mv = mask[v * npix + pix];
                    sI += av * mv;
                    sU += av + mv - av * mv;
                    sM += mv;
                }
            }
        }
    }
    double r;
    r = block_sum_n((double)ssum, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 4, r);
    r = block_sum_n((double)sL, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 3, r);
    if (ch == 0) {
        r = block_sum_n((double)sI, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 0, r);
        r = block_sum_n((double)sU, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 1, r);
        r = block_sum_n((double)sM, scratch); if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 2, r);
    }
}

// The adjoints filtered back (same window: the taps are symmetric) and combined with the L1 sign term; same marching
// structure, maps (A1, A2) packed + A3.  Channel-0 CTAs also write d_alpha (IoU quotient rule).
__global__ void __launch_bounds__(256, 2)
loss_bwd_march_kernel(LossDims d, MarchGeom mg, const float *__restrict__ rgb, const float *__restrict__ timg,
                      const float *__restrict__ mask, const float *__restrict__ adj, const double *__restrict__ stats,
                      float img_lambda, float *__restrict__ d_rgb, float *__restrict__ d_alpha)
{
    extern __shared__ __align__(16) unsigned char march_smem[];
    float2 *ring = reinterpret_cast<float2 *>(march_smem);      // [3 * RB][NT]  (A1, A2)
    float *ring3 = reinterpret_cast<float *>(ring + 3 * RB * mg.NT); // [3 * RB][NT]  A3
    float2 *ex01 = reinterpret_cast<float2 *>(ring3 + 3 * RB * mg.NT); // [RB][RS2]
    float *ex2 = reinterpret_cast<float *>(ex01 + RB * mg.RS2);  // [RB][RS1]
    float *ob = ex2 + RB * mg.RS1;                               // [3][RB][OS]
    const int t = threadIdx.x, NT = mg.NT;
    const int ch = blockIdx.x % 3, strip = blockIdx.x / 3, v = blockIdx.z;
    const int x0 = strip * mg.SW, yb = blockIdx.y * mg.B, yend = min(yb + mg.B, d.H);
    const int nb = (yend - yb + RB - 1) / RB;
    const size_t npix = (size_t)d.H * d.W;
    const int xc = x0 - HALO + t;
    const bool col_on = t < mg.SW + 2 * HALO;
    const bool col_in = col_on && xc >= 0 && xc < d.W;
    const float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix, *a2 = a1 + npix, *a3 = a2 + npix;
    const double I = stats[v * NSTAT + 0] + 1e-6, U = stats[v * NSTAT + 1] + 1e-6, msum = stats[v * NSTAT + 2];
    const float l1 = img_lambda == 0.0f ? 0.0f : (float)((double)img_lambda / msum); // lambda 0: no 0 * 0 / 0 for an empty mask
    const float iou_m = (float)(-1.0 / U), iou_c = (float)(I / (U * U)); // d(1 - I/U)/da = -m/U + I (1 - m) / U^2
    float2 c[6];
    load_taps(c);
    auto stage = [&](int u) {
        if (col_on && u <= nb + 1) {
            float2 *dst = ring + (size_t)((u % 3) * RB) * NT + t;
            float *dst3 = ring3 + (size_t)((u % 3) * RB) * NT + t;
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int y = yb + RB * (u - 1) + k;
                const bool ok = col_in && y >= 0 && y < d.H;
                const size_t pix = ok ? (size_t)y * d.W + xc : 0;
                cp_async4_z(&dst[k * NT].x, a1 + pix, ok);
                cp_async4_z(&dst[k * NT].y, a2 + pix, ok);
                cp_async4_z(&dst3[k * NT], a3 + pix, ok);
            }
        }
        cp_async_commit_l();
    };
    stage(0); stage(1); stage(2);
    for (int i = 0; i < nb; ++i) {
        const int yo = yb + RB * i;
        cp_async_wait_all_l();
        float2 win[WIN8];
        float win3[WIN8];
        if (col_on) {
            const size_t b0 = (size_t)((i % 3) * RB) * NT + t, b1 = (size_t)(((i + 1) % 3) * RB) * NT + t,
                         b2 = (size_t)(((i + 2) % 3) * RB) * NT + t;
#pragma unroll
            for (int k = 0; k < HALO; ++k) { win[k] = ring[b0 + (RB - HALO + k) * NT]; win3[k] = ring3[b0 + (RB - HALO + k) * NT]; }
#pragma unroll
            for (int k = 0; k < RB; ++k) { win[HALO + k] = ring[b1 + k * NT]; win3[HALO + k] = ring3[b1 + k * NT]; }
#pragma unroll
            for (int k = 0; k < HALO; ++k) { win[HALO + RB + k] = ring[b2 + k * NT]; win3[HALO + RB + k] = ring3[b2 + k * NT]; }
        }
        stage(i + 3);
        if (col_on) {
            float2 out[RB];
            float o3[RB];
            window8(win, out, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex01[r * mg.RS2 + t] = out[r];
            window8s(win3, o3, c);
#pragma unroll
            for (int r = 0; r < RB; ++r) ex2[r * mg.RS1 + t] = o3[r];
        }
        __syncthreads();
        if (t < mg.SW) {
            const int row = t & 7, cg = t >> 3;
            float2 in[WIN8], w12[RB];
            float w3[RB];
            {
                const float4 *src = reinterpret_cast<const float4 *>(ex01 + row * mg.RS2 + 8 * cg);
#pragma unroll
                for (int k = 0; k < WIN8 / 2; ++k) {
                    const float4 w = src[k];
                    in[2 * k] = make_float2(w.x, w.y); in[2 * k + 1] = make_float2(w.z, w.w);
                }
                window8(in, w12, c);
            }
            {
                float f[20];
                const float4 *src = reinterpret_cast<const float4 *>(ex2 + row * mg.RS1 + 8 * cg);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 w = src[k];
                    f[4 * k] = w.x; f[4 * k + 1] = w.y; f[4 * k + 2] = w.z; f[4 * k + 3] = w.w;
                }
                window8s(f, w3, c);
            }
            float4 *o1 = reinterpret_cast<float4 *>(ob + (0 * RB + row) * mg.OS + 8 * cg);
            float4 *o2 = reinterpret_cast<float4 *>(ob + (1 * RB + row) * mg.OS + 8 * cg);
            float4 *o3 = reinterpret_cast<float4 *>(ob + (2 * RB + row) * mg.OS + 8 * cg);
            o1[0] = make_float4(w12[0].x, w12[1].x, w12[2].x, w12[3].x); o1[1] = make_float4(w12[4].x, w12[5].x, w12[6].x, w12[7].x);
            o2[0] = make_float4(w12[0].y, w12[1].y, w12[2].y, w12[3].y); o2[1] = make_float4(w12[4].y, w12[5].y, w12[6].y, w12[7].y);
            o3[0] = make_float4(w3[0], w3[1], w3[2], w3[3]); o3[1] = make_float4(w3[4], w3[5], w3[6], w3[7]);
        }
        __syncthreads();
        if (t < mg.SW && x0 + t < d.W) {
            const int x = x0 + t;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int y = yo + r;
                if (y >= yend) break;
                const size_t pix = (size_t)v * npix + (size_t)y * d.W + x;
                const float q = rgb[3 * pix + ch], p = timg[(3 * (size_t)v + ch) * npix + (size_t)y * d.W + x];
                const float diff = p - q;
                const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                const float w1 = ob[(0 * RB + r) * mg.OS + t], w2 = ob[(1 * RB + r) * mg.OS + t], w3 = ob[(2 * RB + r) * mg.OS + t];
                d_rgb[3 * pix + ch] = w1 + 2.f * q * w2 + p * w3 - l1 * sgn;
                if (ch == 0) {
                    const float m = mask[pix];
                    d_alpha[pix] = iou_m * m + iou_c * (1.f - m);
                }
            }
        }
    }
}

// strip / band geometry of the marching kernels for an image size
MarchGeom march_geom(int H, int W, int band)
{
    MarchGeom g;
    const int SW_MAX = 192; // output columns per strip: SW + 10 threads, at most 224
    const int n_strips = (W + SW_MAX - 1) / SW_MAX;
    g.SW = (((W + n_strips - 1) / n_strips) + 7) / 8 * 8;
    g.NT = (g.SW + 2 * HALO + 31) / 32 * 32;
    g.RS2 = g.NT + 2;                 // float2 per exchange row: == 2 (mod 4): 16-byte rows whose phase lanes spread over the banks
    g.RS1 = (g.SW + 12 + 7) / 8 * 8 + 4; // floats per exchange row: >= SW + 12, == 4 (mod 8)
    g.OS = (g.SW + 7) / 8 * 8 + 4;    // floats per output-transpose row: == 4 (mod 8)
    const int Hr = (H + RB - 1) / RB * RB;
    g.B = band < Hr ? band : Hr;
    return g;
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
    if (blockIdx.x == 0 && threadIdx.x == 0) losses[v] = (float)(1.0 - I / U);
    if (!d_alpha) return;
    const float iou_m = (float)(-1.0 / U), iou_c = (float)(I / (U * U));
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
    (void)bx;
    static const int minb = getenv("PS_LOSS_MINB") ? atoi(getenv("PS_LOSS_MINB")) : 2; // A/B switch: registers bounded for 2 / 3 CTAs per SM
    const float coef = d_rgb ? (float)(-(double)ssim_lambda / count) : 0.0f;
    static const bool tiled = getenv("PS_LOSS_TILED") != nullptr; // A/B switch: the 32 x 32 tile kernels
    if (!tiled) {
        static const int band = getenv("PS_LOSS_BAND") ? atoi(getenv("PS_LOSS_BAND")) : 128;
        const MarchGeom mg = march_geom(H, W, band >= RB ? band / RB * RB : 128);
        const size_t ex_f = (size_t)RB * (2 * mg.RS2 * sizeof(float2) + mg.RS1 * sizeof(float)) + 3 * (size_t)RB * mg.OS * sizeof(float);
        const size_t ex_b = (size_t)RB * (mg.RS2 * sizeof(float2) + mg.RS1 * sizeof(float)) + 3 * (size_t)RB * mg.OS * sizeof(float);
        const size_t smem_f = 3 * (size_t)RB * mg.NT * sizeof(float2) + ex_f;
        const size_t smem_b = 3 * (size_t)RB * mg.NT * (sizeof(float2) + sizeof(float)) + ex_b;
        static bool attr_ready[64] = {};
        if (!attr_ready[di] || dev >= 64) { // above 48 KB of dynamic shared memory is opt-in, per device
            if (cudaFuncSetAttribute(ssim_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess) return -1;
            if (cudaFuncSetAttribute(loss_bwd_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess) return -1;
            attr_ready[di] = true;
        }
        const int n_strips = (W + mg.SW - 1) / mg.SW, n_bands = (H + mg.B - 1) / mg.B;
        const dim3 grid(3 * n_strips, n_bands, V);
        ssim_march_kernel<<<grid, mg.NT, smem_f, s>>>(d, mg, rgb, alpha, timg, mask, coef, c1, c2, adj, stats);
        n += 1;
        if (d_rgb) {
            loss_bwd_march_kernel<<<grid, mg.NT, smem_b, s>>>(d, mg, rgb, timg, mask, adj, stats, img_lambda, d_rgb, d_alpha);
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
