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

// horizontal 11-tap pass over a LIN-wide row segment: 4 consecutive outputs per thread
template <int NMAP, typename F>
__device__ __forceinline__ void hpass4(F value_at /* (k, i) -> input i of map k, i in [0, 14) */, float (&out)[NMAP][4])
{
#pragma unroll
    for (int k = 0; k < NMAP; ++k) {
        float in[14];
#pragma unroll
        for (int i = 0; i < 14; ++i) in[i] = value_at(k, i);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float s = 0.f;
#pragma unroll
            for (int tp = 0; tp < 11; ++tp) s = fmaf(c_taps[tp], in[o + tp], s);
            out[k][o] = s;
        }
    }
}

// SSIM forward of one view tile, all three channels.  p = target image (planar), q = render (interleaved).
__global__ void __launch_bounds__(LTHREADS)
ssim_fwd_kernel(LossDims d, const float *__restrict__ rgb, const float *__restrict__ timg, float coef_over_count,
                float c1, float c2, float *__restrict__ adj /* [V,3,3,H,W]: A1 | A2 | A3 per channel */,
                double *__restrict__ stats)
{
    __shared__ float sp[LIN][LIN + 1], sq[LIN][LIN + 1];
    __shared__ float sh[5][LIN][LT + 1]; // + 1: the four rows a warp writes per store fall into different banks
    __shared__ double scratch[LTHREADS / 32];
    const int v = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t npix = (size_t)d.H * d.W;
    const int tid = threadIdx.x;
    float ssum = 0.f;
    for (int ch = 0; ch < 3; ++ch) {
        const float *pch = timg + (3 * (size_t)v + ch) * npix;
        const float *qv = rgb + 3 * (size_t)v * npix;
        __syncthreads(); // the previous channel's vertical pass is done with sh / sp / sq
        for (int i = tid; i < LIN * LIN; i += LTHREADS) {
            const int r = i / LIN, c = i - r * LIN;
            const int y = y0 + r - HALO, x = x0 + c - HALO;
            const bool in = y >= 0 && y < d.H && x >= 0 && x < d.W;
            const size_t pix = (size_t)y * d.W + x;
            sp[r][c] = in ? pch[pix] : 0.f;
            sq[r][c] = in ? qv[3 * pix + ch] : 0.f;
        }
        __syncthreads();
        for (int item = tid; item < LIN * (LT / 4); item += LTHREADS) {
            const int r = item / (LT / 4), cg = (item - r * (LT / 4)) * 4;
            float pv[14], qv14[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) { pv[i] = sp[r][cg + i]; qv14[i] = sq[r][cg + i]; }
            float out[5][4];
            hpass4<5>([&](int k, int i) -> float {
                return k == 0 ? pv[i] : k == 1 ? qv14[i] : k == 2 ? pv[i] * pv[i] : k == 3 ? qv14[i] * qv14[i] : pv[i] * qv14[i];
            }, out);
#pragma unroll
            for (int k = 0; k < 5; ++k)
#pragma unroll
                for (int o = 0; o < 4; ++o) sh[k][r][cg + o] = out[k][o];
        }
        __syncthreads();
        {
            const int c = tid & 31, r0 = (tid >> 5) * 4;
            float res[5][4];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                float in[14];
#pragma unroll
                for (int i = 0; i < 14; ++i) in[i] = sh[k][r0 + i][c];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float s = 0.f;
#pragma unroll
                    for (int tp = 0; tp < 11; ++tp) s = fmaf(c_taps[tp], in[o + tp], s);
                    res[k][o] = s;
                }
            }
            float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix, *a2 = a1 + npix, *a3 = a2 + npix;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int y = y0 + r0 + o, x = x0 + c;
                if (y >= d.H || x >= d.W) continue;
                const bool centre = y >= HALO && y < d.H - HALO && x >= HALO && x < d.W - HALO;
                float g1 = 0.f, g2 = 0.f, g3 = 0.f;
                if (centre) {
                    const float mp = res[0][o], mq = res[1][o];
                    const float vpp = res[2][o] - mp * mp, vqq = res[3][o] - mq * mq, vpq = res[4][o] - mp * mq;
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
    const double r = block_sum((double)ssum, scratch);
    if (threadIdx.x == 0) atomicAdd(stats + v * NSTAT + 4, r);
}

__global__ void __launch_bounds__(LTHREADS)
loss_bwd_kernel(LossDims d, const float *__restrict__ rgb, const float *__restrict__ alpha,
                const float *__restrict__ timg, const float *__restrict__ mask, const float *__restrict__ adj,
                const double *__restrict__ stats, float img_lambda, float *__restrict__ d_rgb, float *__restrict__ d_alpha)
{
    __shared__ float sa[3][LIN][LIN + 1];
    __shared__ float sh[3][LIN][LT + 1];
    const int v = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t npix = (size_t)d.H * d.W;
    const int tid = threadIdx.x;
    const double I = stats[v * NSTAT + 0] + 1e-6, U = stats[v * NSTAT + 1] + 1e-6, msum = stats[v * NSTAT + 2];
    const float l1 = img_lambda == 0.0f ? 0.0f : (float)((double)img_lambda / msum); // lambda 0: no 0 * 0 / 0 for an empty mask
    const float iou_m = (float)(-1.0 / U), iou_c = (float)(I / (U * U)); // d(1 - I/U)/da = -m/U + I (1 - m) / U^2
    const int c = tid & 31, r0 = (tid >> 5) * 4;
    float g[4][3];
    for (int ch = 0; ch < 3; ++ch) {
        const float *a1 = adj + ((3 * (size_t)v + ch) * 3 + 0) * npix;
        __syncthreads();
        for (int i = tid; i < LIN * LIN; i += LTHREADS) {
            const int r = i / LIN, cc = i - r * LIN;
            const int y = y0 + r - HALO, x = x0 + cc - HALO;
            const bool in = y >= 0 && y < d.H && x >= 0 && x < d.W;
            const size_t pix = (size_t)y * d.W + x;
#pragma unroll
            for (int k = 0; k < 3; ++k) sa[k][r][cc] = in ? a1[k * npix + pix] : 0.f;
        }
        __syncthreads();
        for (int item = tid; item < LIN * (LT / 4); item += LTHREADS) {
            const int r = item / (LT / 4), cg = (item - r * (LT / 4)) * 4;
            float out[3][4];
            hpass4<3>([&](int k, int i) -> float { return sa[k][r][cg + i]; }, out);
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int o = 0; o < 4; ++o) sh[k][r][cg + o] = out[k][o];
        }
        __syncthreads();
        float res[3][4];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float in[14];
#pragma unroll
            for (int i = 0; i < 14; ++i) in[i] = sh[k][r0 + i][c];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float s = 0.f;
#pragma unroll
                for (int tp = 0; tp < 11; ++tp) s = fmaf(c_taps[tp], in[o + tp], s);
                res[k][o] = s;
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int y = y0 + r0 + o, x = x0 + c;
            g[o][ch] = 0.f;
            if (y >= d.H || x >= d.W) continue;
            const size_t pix = (size_t)y * d.W + x;
            const float q = rgb[3 * ((size_t)v * npix + pix) + ch], p = timg[(3 * (size_t)v + ch) * npix + pix];
            const float diff = p - q;
            const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            g[o][ch] = res[0][o] + 2.f * q * res[1][o] + p * res[2][o] - l1 * sgn;
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int y = y0 + r0 + o, x = x0 + c;
        if (y >= d.H || x >= d.W) continue;
        const size_t pix = (size_t)v * npix + (size_t)y * d.W + x;
        d_rgb[3 * pix] = g[o][0]; d_rgb[3 * pix + 1] = g[o][1]; d_rgb[3 * pix + 2] = g[o][2];
        const float m = mask[pix];
        d_alpha[pix] = iou_m * m + iou_c * (1.f - m);
    }
    (void)alpha;
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
    loss_reduce_kernel<<<dim3(bx, V), LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, stats);
    const dim3 tiles((W + LT - 1) / LT, (H + LT - 1) / LT, V);
    const double count = 3.0 * (double)(H - 2 * HALO) * (double)(W - 2 * HALO);
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f; // (k * data_range)^2, data_range = 1.0
    int n = 2;
    if (d_rgb) {
        ssim_fwd_kernel<<<tiles, LTHREADS, 0, s>>>(d, rgb, timg, (float)(-(double)ssim_lambda / count), c1, c2, adj, stats);
        loss_bwd_kernel<<<tiles, LTHREADS, 0, s>>>(d, rgb, alpha, timg, mask, adj, stats, img_lambda, d_rgb, d_alpha);
        n += 2;
    } else {
        ssim_fwd_kernel<<<tiles, LTHREADS, 0, s>>>(d, rgb, timg, 0.0f, c1, c2, adj, stats);
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
