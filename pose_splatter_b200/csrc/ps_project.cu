// ps_project.cu -- per-Gaussian stages of the hot path (SURVEY.md 2.2 K1', K2', K7'):
//   project   : fused activation + EWA projection (3D) / activation + extent (2D) -> splat records,
//               tile rectangle, tiles-touched and a per-block sum for the scan      [HBM-bound]
//   scan      : exclusive scan of the per-block sums (one CTA)                        [latency]
//   emit      : in-block scan + (key, value) emission, gid-major / tile row-major     [HBM-bound]
//   project_bwd : chain rule back to the raw rows, atomics into d_params[frame]       [HBM-bound]
// Replaces: adapter activations src/gaussian_renderer.py:183-193 / :314-323 and gsplat's
// fully_fused_projection + isect_tiles (absent from the reference tree).
#include "ps_contract.cuh"
#include "ps_internal.h"

namespace {

__device__ __forceinline__ int block_exclusive_scan_256(int v, int *s_warp, int *block_total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += n;
        }
        s_warp[lane] = wi - w; // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    if (block_total) *block_total = s_warp[32];
    return s_warp[wid] + incl - v;
}

// Stage a block's parameter rows through shared memory with coalesced (128-bit when aligned) loads.
template <int P>
__device__ __forceinline__ void stage_rows(const float *__restrict__ src, int n_rows, float *s_rows)
{
    const int n_float = n_rows * P;
    if ((((uintptr_t)src) & 15u) == 0 && (n_float & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(s_rows);
        for (int i = threadIdx.x; i < (n_float >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = threadIdx.x; i < n_float; i += blockDim.x) s_rows[i] = __ldg(src + i);
    }
}

template <int MODE>
__global__ void __launch_bounds__(PS_PROJ_BLOCK)
project_kernel(PsGeometry g, const float *__restrict__ params, const int32_t *__restrict__ view_frame,
               const float *__restrict__ viewmats, const float *__restrict__ Ks, PsTable t)
{
    constexpr int P = (MODE == PS_MODE_3D) ? 14 : 9;
    __shared__ __align__(16) float s_rows[PS_PROJ_BLOCK * P];
    __shared__ float s_cam[25];
    __shared__ int s_warp[33];
    const int v = blockIdx.y;
    const int g0 = blockIdx.x * PS_PROJ_BLOCK;
    const int n_rows = min(PS_PROJ_BLOCK, g.N - g0);
    const int frame = view_frame[v];
    stage_rows<P>(params + ((size_t)frame * g.N + g0) * P, n_rows, s_rows);
    if (MODE == PS_MODE_3D && threadIdx.x < 25) {
        s_cam[threadIdx.x] = threadIdx.x < 16 ? viewmats[(size_t)v * 16 + threadIdx.x] : Ks[(size_t)v * 9 + threadIdx.x - 16];
    }
    __syncthreads();
    int touched = 0;
    if ((int)threadIdx.x < n_rows) {
        PsRecord rec;
        if (MODE == PS_MODE_3D) {
            PsProj3dAux aux;
            ps_project3d(s_rows + threadIdx.x * P, s_cam, s_cam + 16, g.W, g.H, g.near_plane, g.far_plane, g.radius_clip,
                         g.eps2d, &rec, &aux);
        } else {
            ps_project2d(s_rows + threadIdx.x * P, (uint32_t)(g0 + threadIdx.x), g.W, g.H, &rec);
        }
        const size_t idx = (size_t)v * g.N + g0 + threadIdx.x;
        t.rec0[idx] = make_float4(rec.r0[0], rec.r0[1], rec.r0[2], rec.r0[3]);
        t.rec1[idx] = make_float4(rec.r1[0], rec.r1[1], rec.r1[2], rec.r1[3]);
        t.rec2[idx] = make_float4(rec.r2[0], rec.r2[1], rec.r2[2], rec.r2[3]);
        t.tile_rect[idx] = make_uint2((uint32_t)rec.tile[0] | ((uint32_t)rec.tile[1] << 16),
                                      (uint32_t)rec.tile[2] | ((uint32_t)rec.tile[3] << 16));
        touched = (rec.tile[2] - rec.tile[0]) * (rec.tile[3] - rec.tile[1]);
        t.tiles_touched[idx] = touched;
    }
    int total;
    block_exclusive_scan_256(touched, s_warp, &total);
    if (threadIdx.x == 0) t.block_sums[blockIdx.y * gridDim.x + blockIdx.x] = total;
}

// one CTA: exclusive scan of n block sums in place; sums[n] and *total receive the grand total
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int32_t *sums, int n, int64_t *total)
{
    __shared__ long long s_part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    long long acc = 0;
    for (int i = lo; i < hi; ++i) acc += sums[i];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (int d = 1; d < 1024; d <<= 1) {
        long long add = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    long long run = s_part[threadIdx.x] - acc;
    for (int i = lo; i < hi; ++i) {
        int c = sums[i];
        sums[i] = (int32_t)run;
        run += c;
    }
    if (threadIdx.x == 1023) {
        *total = s_part[1023];
        sums[n] = (int32_t)s_part[1023];
    }
}

template <int MODE>
__global__ void __launch_bounds__(PS_PROJ_BLOCK)
emit_kernel(PsGeometry g, PsTable t, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    __shared__ int s_warp[33];
    const int v = blockIdx.y;
    const int gi = blockIdx.x * PS_PROJ_BLOCK + threadIdx.x;
    const bool live = gi < g.N;
    const size_t idx = (size_t)v * g.N + (live ? gi : 0);
    const int touched = live ? t.tiles_touched[idx] : 0;
    const int excl = block_exclusive_scan_256(touched, s_warp, nullptr);
    if (touched == 0) return;
    size_t out = (size_t)t.block_sums[blockIdx.y * gridDim.x + blockIdx.x] + excl;
    const uint2 tr = t.tile_rect[idx];
    const int tx0 = tr.x & 0xffff, ty0 = tr.x >> 16, tx1 = tr.y & 0xffff, ty1 = tr.y >> 16;
    const uint32_t low = (MODE == PS_MODE_3D) ? __float_as_uint(t.rec2[idx].w) : (uint32_t)gi;
    const uint64_t view_hi = (uint64_t)v << g.tile_bits;
    const uint32_t val = (uint32_t)idx;
    for (int ty = ty0; ty < ty1; ++ty)
        for (int tx = tx0; tx < tx1; ++tx) {
            const uint64_t tile = (uint64_t)(ty * g.tiles_x + tx);
            keys[out] = ((view_hi | tile) << 32) | low;
            vals[out] = val;
            ++out;
        }
}

// ------------------------------------------------------------------------------------------
// Backward of the projection / activations. acc row = 9 sums from the rasterizer backward:
//   3D: v_rgb(3), v_A, v_B, v_C, v_x, v_y, v_opacity        2D: v_rgb(3), sum d_dxr, sum d_dyr, d_theta, d_iax, d_iay, sum G_q
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(PS_PROJ_BLOCK)
project_bwd_kernel(PsGeometry g, const float *__restrict__ params, const int32_t *__restrict__ view_frame,
                   const float *__restrict__ viewmats, const float *__restrict__ Ks, PsTable t,
                   const float *__restrict__ acc, float *__restrict__ d_params)
{
    constexpr int P = (MODE == PS_MODE_3D) ? 14 : 9;
    __shared__ float s_cam[25];
    const int v = blockIdx.y;
    if (MODE == PS_MODE_3D && threadIdx.x < 25)
        s_cam[threadIdx.x] = threadIdx.x < 16 ? viewmats[(size_t)v * 16 + threadIdx.x] : Ks[(size_t)v * 9 + threadIdx.x - 16];
    __syncthreads();
    const int gi = blockIdx.x * PS_PROJ_BLOCK + threadIdx.x;
    if (gi >= g.N) return;
    const size_t idx = (size_t)v * g.N + gi;
    if (t.tiles_touched[idx] == 0) return; // never listed -> no contribution -> zero gradient
    const float4 *a4 = reinterpret_cast<const float4 *>(acc + idx * PS_ACC_STRIDE);
    const float4 q0 = a4[0], q1 = a4[1], q2 = a4[2];
    const float a[9] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x };
    const int frame = view_frame[v];
    const float *row = params + ((size_t)frame * g.N + gi) * P;
    float *d = d_params + ((size_t)frame * g.N + gi) * P;
    float r[P];
#pragma unroll
    for (int k = 0; k < P; ++k) r[k] = __ldg(row + k);
    float out[P];
#pragma unroll
    for (int k = 0; k < P; ++k) out[k] = 0.0f;

    if (MODE == PS_MODE_2D) {
        const float4 r1 = t.rec1[idx];
        const float cs = r1.x, sn = r1.y, iax = r1.z, iay = r1.w;
        const float o = t.rec2[idx].w;
        const float sx = psm_exp(r[2]), sy = psm_exp(r[3]);
        out[0] = -(cs * a[3] - sn * a[4]);
        out[1] = -(sn * a[3] + cs * a[4]);
        out[2] = a[6] * (-(iax * iax) * 4.0f * sx * sx);
        out[3] = a[7] * (-(iay * iay) * 4.0f * sy * sy);
        out[4] = a[5];
#pragma unroll
        for (int k = 0; k < 3; ++k) out[5 + k] = (r[5 + k] >= 0.0f && r[5 + k] <= 1.0f) ? a[k] : 0.0f;
        out[8] = -a[8] * (1.0f - o);
    } else {
        PsRecord rec;
        PsProj3dAux x;
        const int ok = ps_project3d(r, s_cam, s_cam + 16, g.W, g.H, g.near_plane, g.far_plane, g.radius_clip, g.eps2d, &rec, &x);
#pragma unroll
        for (int k = 0; k < 3; ++k) out[10 + k] = (r[10 + k] >= 0.0f && r[10 + k] <= 1.0f) ? a[k] : 0.0f;
        const float o = rec.r1[3];
        out[13] = a[8] * o * (1.0f - o);
        if (ok) {
            const float *V = s_cam, *K = s_cam + 16;
            const float fx = K[0], fy = K[4];
            // conic = inverse(cov2d):  G_S = -X G_X X,  G_X = [[vA, vB/2], [vB/2, vC]]
            const float XA = rec.r1[0], XB = rec.r1[1], XC = rec.r1[2];
            const float gA = a[3], gB = 0.5f * a[4], gC = a[5];
            const float P00 = XA * gA + XB * gB, P01 = XA * gB + XB * gC;
            const float P10 = XB * gA + XC * gB, P11 = XB * gB + XC * gC;
            const float G00 = -(P00 * XA + P01 * XB), G01 = -(P00 * XB + P01 * XC);
            const float G10 = -(P10 * XA + P11 * XB), G11 = -(P10 * XB + P11 * XC);
            const float Gs01 = 0.5f * (G01 + G10);
            const float J[2][3] = { { x.J00, 0.0f, x.J02 }, { 0.0f, x.J11, x.J12 } };
            const float Gs[2][2] = { { G00, Gs01 }, { Gs01, G11 } };
            const float Sc[3][3] = { { x.Sc[0], x.Sc[1], x.Sc[2] }, { x.Sc[1], x.Sc[3], x.Sc[4] }, { x.Sc[2], x.Sc[4], x.Sc[5] } };
            // W = Gs * J (2x3)
            float Wm[2][3];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int c = 0; c < 3; ++c) Wm[p][c] = Gs[p][0] * J[0][c] + Gs[p][1] * J[1][c];
            float GSc[3][3], GJ[2][3];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int c = 0; c < 3; ++c) GSc[rr][c] = J[0][rr] * Wm[0][c] + J[1][rr] * Wm[1][c];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    GJ[p][c] = 2.0f * (Wm[p][0] * Sc[0][c] + Wm[p][1] * Sc[1][c] + Wm[p][2] * Sc[2][c]);
            const float px = x.pc[0], py = x.pc[1], pz = x.pc[2];
            const float rz = 1.0f / pz, rz2 = rz * rz, rz3 = rz2 * rz;
            float vpc0 = fx * rz * a[6];
            float vpc1 = fy * rz * a[7];
            float vpc2 = -(fx * px * a[6] + fy * py * a[7]) * rz2;
            vpc2 += -fx * rz2 * GJ[0][0] - fy * rz2 * GJ[1][1];
            if (!x.clampx) { vpc0 += -fx * rz2 * GJ[0][2]; vpc2 += 2.0f * fx * x.tx * rz3 * GJ[0][2]; }
            else { vpc2 += fx * x.tx * rz3 * GJ[0][2]; }
            if (!x.clampy) { vpc1 += -fy * rz2 * GJ[1][2]; vpc2 += 2.0f * fy * x.ty * rz3 * GJ[1][2]; }
            else { vpc2 += fy * x.ty * rz3 * GJ[1][2]; }
#pragma unroll
            for (int c = 0; c < 3; ++c) out[c] = V[c] * vpc0 + V[4 + c] * vpc1 + V[8 + c] * vpc2;
            // G_Sigma = Rwc^T GSc Rwc
            float Tm[3][3], GS[3][3];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < 3; ++c) Tm[p][c] = GSc[p][0] * V[c] + GSc[p][1] * V[4 + c] + GSc[p][2] * V[8 + c];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int c = 0; c < 3; ++c) GS[rr][c] = V[rr] * Tm[0][c] + V[4 + rr] * Tm[1][c] + V[8 + rr] * Tm[2][c];
            // Sigma = M M^T -> G_M = (G + G^T) M ; M = R diag(s)
            float GM[3][3];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    GM[rr][c] = (GS[rr][0] + GS[0][rr]) * x.M[c] + (GS[rr][1] + GS[1][rr]) * x.M[3 + c] +
                                (GS[rr][2] + GS[2][rr]) * x.M[6 + c];
            float GR[3][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float vs = 0.0f;
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) { GR[rr][c] = GM[rr][c] * x.s[c]; vs += x.R[3 * rr + c] * GM[rr][c]; }
                out[3 + c] = vs * x.s[c];
            }
            const float w = x.qh[0], qx = x.qh[1], qy = x.qh[2], qz = x.qh[3];
            float vq[4];
            vq[0] = 2.0f * (-qz * GR[0][1] + qy * GR[0][2] + qz * GR[1][0] - qx * GR[1][2] - qy * GR[2][0] + qx * GR[2][1]);
            vq[1] = 2.0f * (qy * GR[0][1] + qz * GR[0][2] + qy * GR[1][0] - 2.0f * qx * GR[1][1] - w * GR[1][2] + qz * GR[2][0] + w * GR[2][1] - 2.0f * qx * GR[2][2]);
            vq[2] = 2.0f * (-2.0f * qy * GR[0][0] + qx * GR[0][1] + w * GR[0][2] + qx * GR[1][0] + qz * GR[1][2] - w * GR[2][0] + qz * GR[2][1] - 2.0f * qy * GR[2][2]);
            vq[3] = 2.0f * (-2.0f * qz * GR[0][0] - w * GR[0][1] + qx * GR[0][2] + w * GR[1][0] - 2.0f * qz * GR[1][1] + qy * GR[1][2] + qx * GR[2][0] + qy * GR[2][1]);
            const float dot = vq[0] * w + vq[1] * qx + vq[2] * qy + vq[3] * qz;
            float va[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) va[k] = (vq[k] - dot * x.qh[k]) * x.inv2;
            const float n = x.qn_raw, den = n + 1e-8f;
            const float dq = va[0] * r[6] + va[1] * r[7] + va[2] * r[8] + va[3] * r[9];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g0 = va[k] / den;
                if (n > 0.0f) g0 -= dq / (den * den) * (r[6 + k] / n);
                out[6 + k] = g0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < P; ++k)
        if (out[k] != 0.0f) atomicAdd(d + k, out[k]);
}

__global__ void math_probe_kernel(const float *x, int n, float *y)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sn, cs;
    y[i] = psm_exp(x[i]);
    y[n + i] = psm_log(fabsf(x[i]) + 1e-30f);
    y[2 * n + i] = psm_sigmoid(x[i]);
    psm_sincos(x[i], &sn, &cs);
    y[3 * n + i] = sn;
    y[4 * n + i] = cs;
}

} // namespace

int ps_launch_project(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                      const float *Ks, const PsTable &t, cudaStream_t s)
{
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.V);
    if (g.mode == PS_MODE_3D) project_kernel<PS_MODE_3D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, params, view_frame, viewmats, Ks, t);
    else project_kernel<PS_MODE_2D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, params, view_frame, viewmats, Ks, t);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_scan_block_sums(const PsGeometry &g, const PsTable &t, int64_t *total_out, cudaStream_t s)
{
    const int nb = ((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK) * g.V;
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(t.block_sums, nb, total_out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_emit(const PsGeometry &g, const PsTable &t, uint64_t *keys, uint32_t *vals, cudaStream_t s)
{
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.V);
    if (g.mode == PS_MODE_3D) emit_kernel<PS_MODE_3D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, t, keys, vals);
    else emit_kernel<PS_MODE_2D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, t, keys, vals);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_project_bwd(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                          const float *Ks, const PsTable &t, const float *acc, float *d_params, cudaStream_t s)
{
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.V);
    if (g.mode == PS_MODE_3D)
        project_bwd_kernel<PS_MODE_3D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, params, view_frame, viewmats, Ks, t, acc, d_params);
    else
        project_bwd_kernel<PS_MODE_2D><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, params, view_frame, viewmats, Ks, t, acc, d_params);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_math_probe(const float *x, int n, float *y, cudaStream_t s)
{
    if (n <= 0) return 0;
    math_probe_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, n, y);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
