// ps_project.cu -- per-Gaussian stages of the hot path (SURVEY.md 2.2 K1', K2', K7'):
//   project   : fused activation + EWA projection (3D) / activation + extent (2D) -> splat records,
//               tile rectangle, tiles-touched, and the per-(view,tile) list lengths
//               (CTA histogram in shared memory, one global atomic per (CTA, tile))  [HBM-bound]
//   project_bwd : chain rule back to the raw rows; one thread per (frame, Gaussian) sums the frame's views in
//                 registers and writes the row once -- into d_params, or (K8') with red.global.add into the
//                 owner rank's d_params over NVLink when a frame's cameras are split across GPUs [HBM-bound]
// Replaces: adapter activations src/gaussian_renderer.py:183-193 / :314-323 and gsplat's
// fully_fused_projection + isect_tiles (absent from the reference tree).
#include "ps_contract.cuh"
#include "ps_internal.h"
#include <cstdlib>

namespace {

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
               const float *__restrict__ viewmats, const float *__restrict__ Ks, PsTable t,
               int32_t *__restrict__ tile_counts, int use_smem)
{
    constexpr int P = (MODE == PS_MODE_3D) ? 14 : 9;
    extern __shared__ int s_cnt[]; // [n_tiles] when use_smem
    __shared__ __align__(16) float s_rows[PS_PROJ_BLOCK * P];
    __shared__ float s_cam[25];
    const int v = blockIdx.y;
    const int g0 = blockIdx.x * PS_PROJ_BLOCK;
    const int n_rows = min(PS_PROJ_BLOCK, g.N - g0);
    const int frame = view_frame[v];
    // a frame id outside [0, F) never reads out of bounds: the view renders nothing (the backward's CSR drops it too)
    const bool bad_frame = (unsigned)frame >= (unsigned)g.F;
    if (!bad_frame) stage_rows<P>(params + ((size_t)frame * g.N + g0) * P, n_rows, s_rows);
    if (use_smem)
        for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) s_cnt[i] = 0;
    if (MODE == PS_MODE_3D && threadIdx.x < 25) {
        s_cam[threadIdx.x] = threadIdx.x < 16 ? viewmats[(size_t)v * 16 + threadIdx.x] : Ks[(size_t)v * 9 + threadIdx.x - 16];
    }
    __syncthreads();
    int32_t *cnt_v = tile_counts + (size_t)v * g.n_tiles;
    if ((int)threadIdx.x < n_rows) {
        PsRecord rec;
        if (bad_frame) {
            ps_record_clear(&rec);
        } else if (MODE == PS_MODE_3D) {
            PsProj3dAux aux;
            ps_project3d(s_rows + threadIdx.x * P, s_cam, s_cam + 16, g.W, g.H, g.near_plane, g.far_plane, g.radius_clip,
                         g.eps2d, &rec, &aux, g.activated);
        } else {
            ps_project2d(s_rows + threadIdx.x * P, (uint32_t)(g0 + threadIdx.x), g.W, g.H, &rec);
        }
        const size_t idx = (size_t)v * g.N + g0 + threadIdx.x;
        if (MODE == PS_MODE_3D) {
            float4 *dst = PS_REC(t, idx, 0);
            dst[0] = make_float4(rec.r0[0], rec.r0[1], rec.thr, rec.r1[3]);
            dst[1] = make_float4(psm_mul(0.5f, rec.r1[0]), rec.r1[1], psm_mul(0.5f, rec.r1[2]), 0.0f);
            dst[2] = make_float4(rec.r2[0], rec.r2[1], rec.r2[2], 0.0f);
            dst[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            t.depth[idx] = rec.low;
        } else {
            float4 *dst = PS_REC(t, idx, 0);
            dst[0] = make_float4(rec.r0[0], rec.r0[1], rec.thr, rec.r2[3]);
            dst[1] = make_float4(rec.r1[0], rec.r1[1], rec.r1[2], rec.r1[3]);
            dst[2] = make_float4(rec.r2[0], rec.r2[1], rec.r2[2], 0.0f);
            dst[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        t.tile_rect[idx] = make_uint2((uint32_t)rec.tile[0] | ((uint32_t)rec.tile[1] << 16),
                                      (uint32_t)rec.tile[2] | ((uint32_t)rec.tile[3] << 16));
        t.tiles_touched[idx] = (rec.tile[2] - rec.tile[0]) * (rec.tile[3] - rec.tile[1]);
        for (int ty = rec.tile[1]; ty < rec.tile[3]; ++ty)
            for (int tx = rec.tile[0]; tx < rec.tile[2]; ++tx)
                atomicAdd(use_smem ? &s_cnt[ty * g.tiles_x + tx] : &cnt_v[ty * g.tiles_x + tx], 1);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) {
            const int c = s_cnt[i];
            if (c) atomicAdd(&cnt_v[i], c);
        }
    }
}

// 3D projection with the camera-independent part hoisted: one thread per (frame, Gaussian) applies the adapter and builds
// the world covariance ONCE (ps_gauss3d), then projects into every camera of the frame (ps_view3d; the frame's views
// come from the CSR the forward builds first).  Same records, bit for bit, as one thread per (view, Gaussian).
template <int MINB> // CTAs per SM the register allocation is bounded for
__global__ void __launch_bounds__(PS_PROJ_BLOCK, MINB)
project3d_frames_kernel(PsGeometry g, const float *__restrict__ params, const int32_t *__restrict__ frame_off,
                        const int32_t *__restrict__ frame_views, const float *__restrict__ viewmats,
                        const float *__restrict__ Ks, PsTable t, int32_t *__restrict__ tile_counts, int use_smem)
{
    constexpr int P = 14;
    extern __shared__ int s_cnt[]; // [n_tiles] when use_smem
    __shared__ __align__(16) float s_rows[PS_PROJ_BLOCK * P];
    __shared__ float s_cam[25];
    const int frame = blockIdx.y;
    const int g0 = blockIdx.x * PS_PROJ_BLOCK;
    const int n_rows = min(PS_PROJ_BLOCK, g.N - g0);
    stage_rows<P>(params + ((size_t)frame * g.N + g0) * P, n_rows, s_rows);
    __syncthreads();
    const bool live = (int)threadIdx.x < n_rows;
    const float *row = s_rows + threadIdx.x * P;
    PsRecord rec;
    PsProj3dAux aux;
    float S[6];
    if (live) ps_gauss3d(row, g.activated, &rec, &aux, S);
    const int j0 = frame_off[frame], j1 = frame_off[frame + 1];
    for (int j = j0; j < j1; ++j) {
        const int v = frame_views[j];
        __syncthreads(); // the previous camera's histogram has been flushed, s_cam is free
        if (threadIdx.x < 25)
            s_cam[threadIdx.x] = threadIdx.x < 16 ? viewmats[(size_t)v * 16 + threadIdx.x] : Ks[(size_t)v * 9 + threadIdx.x - 16];
        if (use_smem)
            for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) s_cnt[i] = 0;
        __syncthreads();
        int32_t *cnt_v = tile_counts + (size_t)v * g.n_tiles;
        if (live) {
            ps_view3d(row, S, s_cam, s_cam + 16, g.W, g.H, g.near_plane, g.far_plane, g.radius_clip, g.eps2d, &rec, &aux);
            const size_t idx = (size_t)v * g.N + g0 + threadIdx.x;
            float4 *dst = PS_REC(t, idx, 0);
            dst[0] = make_float4(rec.r0[0], rec.r0[1], rec.thr, rec.r1[3]);
            dst[1] = make_float4(psm_mul(0.5f, rec.r1[0]), rec.r1[1], psm_mul(0.5f, rec.r1[2]), 0.0f);
            dst[2] = make_float4(rec.r2[0], rec.r2[1], rec.r2[2], 0.0f);
            dst[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            t.depth[idx] = rec.low;
            t.tile_rect[idx] = make_uint2((uint32_t)rec.tile[0] | ((uint32_t)rec.tile[1] << 16),
                                          (uint32_t)rec.tile[2] | ((uint32_t)rec.tile[3] << 16));
            t.tiles_touched[idx] = (rec.tile[2] - rec.tile[0]) * (rec.tile[3] - rec.tile[1]);
            for (int ty = rec.tile[1]; ty < rec.tile[3]; ++ty)
                for (int tx = rec.tile[0]; tx < rec.tile[2]; ++tx)
                    atomicAdd(use_smem ? &s_cnt[ty * g.tiles_x + tx] : &cnt_v[ty * g.tiles_x + tx], 1);
        }
        if (use_smem) {
            __syncthreads();
            for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) {
                const int c = s_cnt[i];
                if (c) atomicAdd(&cnt_v[i], c);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward of the projection / activations. acc row = 9 sums from the rasterizer backward:
//   3D: v_rgb(3), v_A, v_B, v_C, v_x, v_y, v_opacity        2D: v_rgb(3), sum d_dxr, sum d_dyr, d_theta, d_iax, d_iay, sum G_q
// ------------------------------------------------------------------------------------------
// Gradient of one (view, Gaussian) pair w.r.t. its raw parameter row, added into out[P].
//   r   the raw row, cam  viewmat[16] | K[9] of the view (3D), a  the nine sums of the rasterizer backward
//   3D: rec / x / S hold the camera-independent part of the projection (ps_gauss3d), computed once per Gaussian by the caller
template <int MODE>
__device__ __forceinline__ void project_bwd_row(const PsGeometry &g, const PsTable &t, size_t idx, const float *r,
                                                const float *cam, const float (&a)[9], PsRecord &rec, PsProj3dAux &x,
                                                const float *S, float (&out)[(MODE == PS_MODE_3D) ? 14 : 9])
{
    constexpr int P = (MODE == PS_MODE_3D) ? 14 : 9;
    float o[P];
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = 0.0f;
    if (MODE == PS_MODE_2D) {
        const float4 r1 = *PS_REC(t, idx, 1);
        const float cs = r1.x, sn = r1.y, iax = r1.z, iay = r1.w;
        const float op = PS_REC(t, idx, 0)->w;
        const float sx = psm_exp(r[2]), sy = psm_exp(r[3]);
        o[0] = -(cs * a[3] - sn * a[4]);
        o[1] = -(sn * a[3] + cs * a[4]);
        o[2] = a[6] * (-(iax * iax) * 4.0f * sx * sx);
        o[3] = a[7] * (-(iay * iay) * 4.0f * sy * sy);
        o[4] = a[5];
#pragma unroll
        for (int k = 0; k < 3; ++k) o[5 + k] = (r[5 + k] >= 0.0f && r[5 + k] <= 1.0f) ? a[k] : 0.0f;
        o[8] = -a[8] * (1.0f - op);
    } else {
        const int ok = ps_view3d(r, S, cam, cam + 16, g.W, g.H, g.near_plane, g.far_plane, g.radius_clip, g.eps2d, &rec, &x);
        // v[14]: gradient w.r.t. the activated values (means | scales | quats | colours | opacity), i.e. what gsplat's
        // backward hands to autograd; the adapter's own vector-Jacobian product maps it to the raw row below
        float v[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) v[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[10 + k] = a[k];
        v[13] = a[8];
        if (ok) {
            const float *V = cam, *K = cam + 16;
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
            for (int c = 0; c < 3; ++c) v[c] = V[c] * vpc0 + V[4 + c] * vpc1 + V[8 + c] * vpc2;
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
                v[3 + c] = vs;
            }
            const float w = x.qh[0], qx = x.qh[1], qy = x.qh[2], qz = x.qh[3];
            float vq[4];
            vq[0] = 2.0f * (-qz * GR[0][1] + qy * GR[0][2] + qz * GR[1][0] - qx * GR[1][2] - qy * GR[2][0] + qx * GR[2][1]);
            vq[1] = 2.0f * (qy * GR[0][1] + qz * GR[0][2] + qy * GR[1][0] - 2.0f * qx * GR[1][1] - w * GR[1][2] + qz * GR[2][0] + w * GR[2][1] - 2.0f * qx * GR[2][2]);
            vq[2] = 2.0f * (-2.0f * qy * GR[0][0] + qx * GR[0][1] + w * GR[0][2] + qx * GR[1][0] + qz * GR[1][2] - w * GR[2][0] + qz * GR[2][1] - 2.0f * qy * GR[2][2]);
            vq[3] = 2.0f * (-2.0f * qz * GR[0][0] - w * GR[0][1] + qx * GR[0][2] + w * GR[1][0] - 2.0f * qz * GR[1][1] + qy * GR[1][2] + qx * GR[2][0] + qy * GR[2][1]);
            const float dot = vq[0] * w + vq[1] * qx + vq[2] * qy + vq[3] * qz;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[6 + k] = (vq[k] - dot * x.qh[k]) * x.inv2;
        }
        if (g.activated) {
#pragma unroll
            for (int k = 0; k < 14; ++k) o[k] = v[k];
        } else {
            ps_adapter3d_vjp(r, x.s, x.qn_raw, rec.r1[3], v, o);
        }
    }
#pragma unroll
    for (int k = 0; k < P; ++k) out[k] += o[k];
}

// One thread per (frame, Gaussian): loops over the views of the frame (CSR frame_off / frame_views built by the
// forward), sums their gradients in registers and writes the row ONCE.
//   peers == nullptr : plain stores into d_params (no atomics, no memset: rows without gradient get zeros)
//   peers != nullptr : K8' fused gradient exchange.  Frame f is owned by rank f % world.  The CTA's 256 finished
//                      rows are staged in shared memory and pushed with coalesced 8-byte stores straight into the
//                      owner's staging buffer, slot [my_rank][f / world] (local memory, or a peer GPU's over
//                      NVLink): the transfer of one block overlaps the arithmetic of the next, no atomics, no
//                      collective launch.  After a barrier the owner adds its `world` slots (ps_peer_sum).
template <int MODE, int MINB> // MINB = CTAs per SM the register allocation is bounded for (measured: see the launcher)
__global__ void __launch_bounds__(PS_PROJ_BLOCK, MINB)
project_bwd_kernel(PsGeometry g, const float *__restrict__ params, const int32_t *__restrict__ frame_off,
                   const int32_t *__restrict__ frame_views, const float *__restrict__ viewmats,
                   const float *__restrict__ Ks, PsTable t, const float *__restrict__ acc, float *__restrict__ d_params,
                   float *const *__restrict__ peers, int my_rank, int world)
{
    constexpr int P = (MODE == PS_MODE_3D) ? 14 : 9;
    __shared__ float s_cam[25];
    __shared__ __align__(16) float s_out[PS_PROJ_BLOCK * P]; // peers mode: the CTA's rows, pushed out coalesced
    const int frame = blockIdx.y;
    const int gi = blockIdx.x * PS_PROJ_BLOCK + threadIdx.x;
    const bool live = gi < g.N;
    float r[P], out[P];
#pragma unroll
    for (int k = 0; k < P; ++k) { r[k] = 0.0f; out[k] = 0.0f; }
    if (live) {
        const float *row = params + ((size_t)frame * g.N + gi) * P;
#pragma unroll
        for (int k = 0; k < P; ++k) r[k] = __ldg(row + k);
    }
    PsRecord rec;
    PsProj3dAux aux;
    float S[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    if (MODE == PS_MODE_3D && live) ps_gauss3d(r, g.activated, &rec, &aux, S); // once per Gaussian, shared by the frame's cameras
    const int j0 = frame_off[frame], j1 = frame_off[frame + 1];
    for (int j = j0; j < j1; ++j) {
        const int v = frame_views[j];
        if (MODE == PS_MODE_3D) {
            __syncthreads();
            if (threadIdx.x < 25)
                s_cam[threadIdx.x] = threadIdx.x < 16 ? viewmats[(size_t)v * 16 + threadIdx.x] : Ks[(size_t)v * 9 + threadIdx.x - 16];
            __syncthreads();
        }
        if (!live) continue;
        const size_t idx = (size_t)v * g.N + gi;
        if (t.tiles_touched[idx] == 0) continue; // never listed -> no contribution -> zero gradient
        float a[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = __ldg(acc + idx * PS_ACC_STRIDE + k);
        bool any = false; // listed but never a contributor (hidden behind saturated pixels): gradient exactly zero
#pragma unroll
        for (int k = 0; k < 9; ++k) any = any || (a[k] != 0.0f);
        if (!any) continue;
        project_bwd_row<MODE>(g, t, idx, r, s_cam, a, rec, aux, S, out);
    }
    if (peers == nullptr) {
        if (!live) return;
        float *d = d_params + ((size_t)frame * g.N + gi) * P;
        if (MODE == PS_MODE_3D) { // 56-byte rows: 8-byte aligned
#pragma unroll
            for (int k = 0; k < P; k += 2) *reinterpret_cast<float2 *>(d + k) = make_float2(out[k], out[k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < P; ++k) d[k] = out[k];
        }
        return;
    }
    // push the block's rows (zeros included: every slot is fully rewritten every step, nothing to clear)
#pragma unroll
    for (int k = 0; k < P; ++k) s_out[threadIdx.x * P + k] = out[k];
    __syncthreads();
    const int g0 = blockIdx.x * PS_PROJ_BLOCK;
    const int n_rows = min(PS_PROJ_BLOCK, g.N - g0);
    const int frames_per_rank = (g.F + world - 1) / world;
    float *dst = peers[frame % world] + (((size_t)my_rank * frames_per_rank + frame / world) * g.N + g0) * P;
    const int n_float = n_rows * P;
    if (((reinterpret_cast<uintptr_t>(dst) & 7u) == 0) && (n_float & 1) == 0) {
        float2 *d2 = reinterpret_cast<float2 *>(dst);
        const float2 *s2 = reinterpret_cast<const float2 *>(s_out);
        for (int i = threadIdx.x; i < (n_float >> 1); i += PS_PROJ_BLOCK) d2[i] = s2[i];
    } else {
        for (int i = threadIdx.x; i < n_float; i += PS_PROJ_BLOCK) dst[i] = s_out[i];
    }
}

// out[i] = sum over the `world` slots of the local staging buffer (the gradients pushed by every rank)
__global__ void __launch_bounds__(256) peer_sum_kernel(const float *__restrict__ stage, int world, size_t n, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.0f;
    for (int r = 0; r < world; ++r) acc += stage[(size_t)r * n + i];
    out[i] = acc;
}

// view -> frame map to CSR (views of every frame): one CTA; order inside a frame is arbitrary
__global__ void __launch_bounds__(1024)
frame_csr_kernel(const int32_t *__restrict__ view_frame, int V, int F, int32_t *__restrict__ frame_off,
                 int32_t *__restrict__ cursor, int32_t *__restrict__ frame_views, const float *__restrict__ background,
                 float *__restrict__ bg_saved)
{
    __shared__ int s_carry;
    __shared__ int s_warp[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 3 && bg_saved) bg_saved[tid] = background[tid]; // the backward composites against the forward's background
    for (int f = tid; f <= F; f += 1024) frame_off[f] = 0;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int v = tid; v < V; v += 1024) {
        const int f = view_frame[v];
        if (f >= 0 && f < F) atomicAdd(&frame_off[f + 1], 1);
    }
    __syncthreads();
    // inclusive scan of frame_off[1..F] in chunks of 1024
    for (int c0 = 1; c0 <= F; c0 += 1024) {
        const int i = c0 + tid;
        const int val = i <= F ? frame_off[i] : 0;
        int incl = val;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int w = s_warp[lane];
            int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += n;
            }
            s_warp[lane] = wi - w;
            if (lane == 31) s_warp[32] = wi;
        }
        __syncthreads();
        if (i <= F) {
            const int res = s_carry + s_warp[wid] + incl;
            frame_off[i] = res;
            cursor[i - 1] = res - val; // start of frame i-1
        }
        __syncthreads();
        if (tid == 0) s_carry += s_warp[32];
        __syncthreads();
    }
    for (int v = tid; v < V; v += 1024) {
        const int f = view_frame[v];
        if (f >= 0 && f < F) frame_views[atomicAdd(&cursor[f], 1)] = v;
    }
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

// adapter stage alone, through the device functions the projection kernels use (parity probe)
__global__ void adapter3d_probe_kernel(const float *__restrict__ rows, int n, const float *__restrict__ v_act,
                                       float *__restrict__ act, float *__restrict__ d_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[14], s[3], qa[4], qn, rgb[3], o;
#pragma unroll
    for (int k = 0; k < 14; ++k) r[k] = rows[(size_t)i * 14 + k];
    ps_adapter3d(r, 0, s, qa, &qn, rgb, &o);
    float *a = act + (size_t)i * 14;
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = r[k]; a[3 + k] = s[k]; a[10 + k] = rgb[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) a[6 + k] = qa[k];
    a[13] = o;
    if (v_act && d_rows) {
        float v[14], out[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) v[k] = v_act[(size_t)i * 14 + k];
        ps_adapter3d_vjp(r, s, qn, o, v, out);
#pragma unroll
        for (int k = 0; k < 14; ++k) d_rows[(size_t)i * 14 + k] = out[k];
    }
}

} // namespace

int ps_launch_adapter3d_probe(const float *rows, int n, const float *v_act, float *act, float *d_rows, cudaStream_t s)
{
    if (n <= 0) return 0;
    adapter3d_probe_kernel<<<(n + 127) / 128, 128, 0, s>>>(rows, n, v_act, act, d_rows);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_project(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                      const float *Ks, const PsTable &t, int32_t *tile_counts, const int32_t *frame_off,
                      const int32_t *frame_views, cudaStream_t s)
{
    if (g.N == 0 || g.V == 0) return 0;
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.V);
    const int use_smem = g.n_tiles <= PS_HIST_SMEM_TILES;
    const size_t dyn = use_smem ? (size_t)g.n_tiles * sizeof(int) : 0;
    static const bool per_view = getenv("PS_PROJECT_PER_VIEW") != nullptr; // A/B switch for measurements
    if (g.mode == PS_MODE_3D && frame_off && !per_view && g.F > 0) {
        // a view whose frame id is out of range is in no frame's list: it must still read as "nothing listed"
        if (cudaMemsetAsync(t.tiles_touched, 0, (size_t)g.V * g.N * sizeof(int32_t), s) != cudaSuccess) return -1;
        dim3 fgrid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.F);
        static const int minb = getenv("PS_PROJ_MINB") ? atoi(getenv("PS_PROJ_MINB")) : 4; // A/B switch for measurements
        if (minb == 3) project3d_frames_kernel<3><<<fgrid, PS_PROJ_BLOCK, dyn, s>>>(g, params, frame_off, frame_views, viewmats, Ks, t, tile_counts, use_smem);
        else if (minb == 5) project3d_frames_kernel<5><<<fgrid, PS_PROJ_BLOCK, dyn, s>>>(g, params, frame_off, frame_views, viewmats, Ks, t, tile_counts, use_smem);
        else project3d_frames_kernel<4><<<fgrid, PS_PROJ_BLOCK, dyn, s>>>(g, params, frame_off, frame_views, viewmats, Ks, t, tile_counts, use_smem);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    if (g.mode == PS_MODE_3D)
        project_kernel<PS_MODE_3D><<<grid, PS_PROJ_BLOCK, dyn, s>>>(g, params, view_frame, viewmats, Ks, t, tile_counts, use_smem);
    else
        project_kernel<PS_MODE_2D><<<grid, PS_PROJ_BLOCK, dyn, s>>>(g, params, view_frame, viewmats, Ks, t, tile_counts, use_smem);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_frame_csr(const PsGeometry &g, const int32_t *view_frame, int32_t *frame_off, int32_t *cursor,
                        int32_t *frame_views, const float *background, float *bg_saved, cudaStream_t s)
{
    frame_csr_kernel<<<1, 1024, 0, s>>>(view_frame, g.V, g.F, frame_off, cursor, frame_views, background, bg_saved);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_project_bwd(const PsGeometry &g, const float *params, const int32_t *frame_off, const int32_t *frame_views,
                          const float *viewmats, const float *Ks, const PsTable &t, const float *acc, float *d_params,
                          float *const *peers, int my_rank, int world, cudaStream_t s)
{
    if (g.N == 0 || g.F == 0) return 0;
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.F);
    // registers bounded for 2 / 3 / 4 CTAs per SM (122 / 80 / 64 registers): measured 1.41 / 1.23 / 1.17 ms at c2
    static const int minb = getenv("PS_PBWD_MINB") ? atoi(getenv("PS_PBWD_MINB")) : 4; // A/B switch for measurements
#define PS_PBWD(MODE, MB) project_bwd_kernel<MODE, MB><<<grid, PS_PROJ_BLOCK, 0, s>>>(g, params, frame_off, frame_views, viewmats, Ks, t, acc, d_params, peers, my_rank, world)
    if (g.mode == PS_MODE_3D) {
        if (minb == 2) PS_PBWD(PS_MODE_3D, 2); else if (minb == 3) PS_PBWD(PS_MODE_3D, 3); else PS_PBWD(PS_MODE_3D, 4);
    } else {
        PS_PBWD(PS_MODE_2D, 3);
    }
#undef PS_PBWD
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_peer_sum(const float *stage, int world, size_t n, float *out, cudaStream_t s)
{
    if (n == 0) return 0;
    peer_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(stage, world, n, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_math_probe(const float *x, int n, float *y, cudaStream_t s)
{
    if (n <= 0) return 0;
    math_probe_kernel<<<(n + 255) / 256, 256, 0, s>>>(x, n, y);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
