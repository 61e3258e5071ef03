// ps_head.cu -- the parameter-head tail that produces the rows render() takes, forward and backward.
// Replaces, in the reference's src/model.py: the activations of get_gaussian_params_from_volume_unified (:207-257:
// colour sigmoid + clip, log-scale offset, opacity logit from the occupancy probability, grid + 2 voxel tanh(delta)) and
// apply_pose_transform_3d (:261-298: yaw + translation of the means, and the quaternion composition through
// quaternion_matrix_torch_batch :368-391 and quaternion_from_matrix_torch_batch :394-421 -- a float64
// torch.linalg.eigh per Gaussian plus ~40 small ATen kernels and their autograd).  One thread per Gaussian; the
// 4x4 symmetric eigenproblem is a register-resident cyclic Jacobi in fp64, and the backward differentiates the
// top eigenvector analytically (sum over the other eigenpairs), so nothing but the rows and their gradients touches HBM.
// The 3x3 block of :380-388 is reproduced entry by entry (it is not the textbook rotation matrix; see oracle/param_head_ref.py).
#include "ps_internal.h"

namespace {

struct HeadArgs {
    int mode, n, pose;
    float voxel2, inv1mpt, pt, clip_lo, clip_hi;
    float c, s, px, py, pz; // yaw cos / sin rounded to fp32 like the reference's rot_mat, translation
    const float *poses;     // optional [F,5] (c, s, px, py, pz) per frame: rows of several frames in one launch
    const int32_t *row_frame; // [n] frame of every row (with poses)
    int n_frames;             // rows of `poses`
};

__device__ __forceinline__ void row_pose(const HeadArgs &a, int i, float &c, float &s, float &px, float &py, float &pz)
{
    c = a.c; s = a.s; px = a.px; py = a.py; pz = a.pz;
    if (a.poses) {
        const unsigned f = (unsigned)a.row_frame[i];
        if (f >= (unsigned)a.n_frames) { // a frame id outside the pose table: never read out of bounds, poison the row
            c = s = px = py = pz = __int_as_float(0x7fc00000);
            return;
        }
        const float *p = a.poses + 5 * (size_t)f;
        c = p[0]; s = p[1]; px = p[2]; py = p[3]; pz = p[4];
    }
}

constexpr int HEAD_THREADS = 128;
constexpr double QEPS = 4.0 * 2.220446049250313e-16;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// one Jacobi rotation annihilating A[P][Q] of a symmetric 4x4 (all indices compile-time: the arrays live in registers)
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&A)[4][4], double (&V)[4][4])
{
    const double apq = A[P][Q];
    if (apq == 0.0) return;
    const double theta = (A[Q][Q] - A[P][P]) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
    for (int k = 0; k < 4; ++k) { // columns P, Q of A
        const double akp = A[k][P], akq = A[k][Q];
        A[k][P] = c * akp - s * akq;
        A[k][Q] = s * akp + c * akq;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { // rows P, Q of A
        const double apk = A[P][k], aqk = A[Q][k];
        A[P][k] = c * apk - s * aqk;
        A[Q][k] = s * apk + c * aqk;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double vkp = V[k][P], vkq = V[k][Q];
        V[k][P] = c * vkp - s * vkq;
        V[k][Q] = s * vkp + c * vkq;
    }
}

// eigen-decomposition of a symmetric 4x4: on return A is diagonal (eigenvalues), the columns of V are the eigenvectors
__device__ __forceinline__ void jacobi4(double (&A)[4][4], double (&V)[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[0][3] * A[0][3] + A[1][2] * A[1][2] + A[1][3] * A[1][3] + A[2][3] * A[2][3];
        const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2] + A[3][3] * A[3][3];
        if (off <= 1e-34 * diag || off == 0.0) break;
        jacobi_rotate<0, 1>(A, V); jacobi_rotate<0, 2>(A, V); jacobi_rotate<0, 3>(A, V);
        jacobi_rotate<1, 2>(A, V); jacobi_rotate<1, 3>(A, V); jacobi_rotate<2, 3>(A, V);
    }
}

// Everything the quaternion path computes between the raw quaternion and the symmetric K / 3.
struct QuatChain {
    double qs[4];   // q * sqrt(2 / n)
    double k, n;    // sqrt(2 / n), q.q
    bool small;     // n < 4 eps: the block is the identity
    double K[4][4];
};

__device__ __forceinline__ void quat_chain(const float (&q)[4], float c, float s, QuatChain &z)
{
    double qd[4] = { (double)q[0], (double)q[1], (double)q[2], (double)q[3] };
    z.n = qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3];
    z.small = z.n < QEPS;
    const double nn = z.small ? 1.0 : z.n;
    z.k = sqrt(2.0 / nn);
#pragma unroll
    for (int i = 0; i < 4; ++i) z.qs[i] = qd[i] * z.k;
#define O(a, b) (z.qs[a] * z.qs[b])
    double r[3][3];
    r[0][0] = 1.0 - O(2, 2) - O(3, 3); r[0][1] = O(1, 2) - O(3, 0);       r[0][2] = O(1, 3) + O(2, 0);
    r[1][0] = O(1, 2) - O(3, 0);       r[1][1] = (1.0 + O(0, 0)) - O(0, 0); r[1][2] = O(2, 3) - O(1, 0);
    r[2][0] = O(1, 3) - O(2, 0);       r[2][1] = O(2, 3) + O(1, 0);       r[2][2] = 1.0 - O(1, 1) - O(2, 2);
#undef O
    float m[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float r0 = z.small ? (j == 0 ? 1.f : 0.f) : (float)r[0][j];
        const float r1 = z.small ? (j == 1 ? 1.f : 0.f) : (float)r[1][j];
        const float r2 = z.small ? (j == 2 ? 1.f : 0.f) : (float)r[2][j];
        m[0][j] = c * r0 - s * r1; // Rz4 @ r in fp32 (the reference's einsum on float32 tensors)
        m[1][j] = s * r0 + c * r1;
        m[2][j] = r2;
    }
    const double m00 = m[0][0], m01 = m[0][1], m02 = m[0][2], m10 = m[1][0], m11 = m[1][1], m12 = m[1][2],
                 m20 = m[2][0], m21 = m[2][1], m22 = m[2][2];
    z.K[0][0] = (m00 - m11 - m22) / 3.0; z.K[1][1] = (m11 - m00 - m22) / 3.0;
    z.K[2][2] = (m22 - m00 - m11) / 3.0; z.K[3][3] = (m00 + m11 + m22) / 3.0;
    z.K[0][1] = z.K[1][0] = (m01 + m10) / 3.0; z.K[0][2] = z.K[2][0] = (m02 + m20) / 3.0;
    z.K[0][3] = z.K[3][0] = (m21 - m12) / 3.0; z.K[1][2] = z.K[2][1] = (m12 + m21) / 3.0;
    z.K[1][3] = z.K[3][1] = (m02 - m20) / 3.0; z.K[2][3] = z.K[3][2] = (m10 - m01) / 3.0;
}

// index of the largest eigenvalue and its eigenvector (select chains: no dynamic register indexing)
__device__ __forceinline__ int top_eigen(const double (&A)[4][4], const double (&V)[4][4], double (&t)[4], double &lam)
{
    int top = 0;
    lam = A[0][0];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = V[k][0];
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        const bool better = A[i][i] > lam;
        lam = better ? A[i][i] : lam;
        top = better ? i : top;
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = better ? V[k][i] : t[k];
    }
    return top;
}

__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_kernel(HeadArgs a, const float *__restrict__ net_out, const float *__restrict__ probs, const float *__restrict__ grid,
                const float *__restrict__ scale0, float *__restrict__ rows)
{
    const int i = blockIdx.x * HEAD_THREADS + threadIdx.x;
    if (i >= a.n) return;
    const float sc = __ldg(scale0);
    const float x = a.inv1mpt * (probs[i] - a.pt);
    const float xc = fminf(fmaxf(x, 1e-6f), 1.0f - 1e-6f);
    const float logit = logf(xc / (1.0f - xc));
    if (a.mode == PS_MODE_2D) {
        const float *in = net_out + 9 * (size_t)i;
        float *o = rows + 9 * (size_t)i;
        o[0] = in[0]; o[1] = in[1]; o[2] = in[2] + sc; o[3] = in[3] + sc; o[4] = in[4];
#pragma unroll
        for (int k = 0; k < 3; ++k) o[5 + k] = fminf(fmaxf(sigmoidf_(in[5 + k]), a.clip_lo), a.clip_hi);
        o[8] = logit;
        return;
    }
    const float *in = net_out + 14 * (size_t)i; // quats 4 | scales 3 | opacity 1 (unused) | colours 3 | delta means 3
    float *o = rows + 14 * (size_t)i;           // means 3 | log scales 3 | quats 4 | colours 3 | logit opacity
    float m[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) m[k] = grid[3 * (size_t)i + k] + a.voxel2 * tanhf(in[11 + k]);
    float q[4] = { in[0], in[1], in[2], in[3] };
    if (a.pose) {
        float pc, ps, px, py, pz;
        row_pose(a, i, pc, ps, px, py, pz);
        const float mx = m[0] * pc - m[1] * ps + px, my = m[0] * ps + m[1] * pc + py;
        m[0] = mx; m[1] = my; m[2] = m[2] + pz;
        QuatChain z;
        quat_chain(q, pc, ps, z);
        double V[4][4], t[4], lam;
        jacobi4(z.K, V);
        top_eigen(z.K, V, t, lam);
        const double sgn = t[3] < 0.0 ? -1.0 : 1.0; // w = component 3 of the (x, y, z, w) eigenvector; flipped to w >= 0
        q[0] = (float)(sgn * t[3]); q[1] = (float)(sgn * t[0]); q[2] = (float)(sgn * t[1]); q[3] = (float)(sgn * t[2]);
    }
    o[0] = m[0]; o[1] = m[1]; o[2] = m[2];
    o[3] = in[4] + sc; o[4] = in[5] + sc; o[5] = in[6] + sc;
    o[6] = q[0]; o[7] = q[1]; o[8] = q[2]; o[9] = q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[10 + k] = fminf(fmaxf(sigmoidf_(in[8 + k]), a.clip_lo), a.clip_hi);
    o[13] = logit;
}

__global__ void __launch_bounds__(HEAD_THREADS)
head_bwd_kernel(HeadArgs a, const float *__restrict__ net_out, const float *__restrict__ probs,
                const float *__restrict__ d_rows, float *__restrict__ d_net, float *__restrict__ d_probs,
                float *__restrict__ d_scale0)
{
    __shared__ float s_part[HEAD_THREADS / 32];
    const int i = blockIdx.x * HEAD_THREADS + threadIdx.x;
    float dsc = 0.0f;
    if (i < a.n) {
        const int P = a.mode == PS_MODE_2D ? 9 : 14;
        const float *g = d_rows + (size_t)P * i;
        const float *in = net_out + (size_t)P * i;
        float *d = d_net + (size_t)P * i;
        const float x = a.inv1mpt * (probs[i] - a.pt);
        const bool pass = x >= 1e-6f && x <= 1.0f - 1e-6f; // clamp passes the gradient on the closed interval
        const float g_logit = g[P - 1];
        d_probs[i] = pass ? g_logit * a.inv1mpt / (x * (1.0f - x)) : 0.0f;
        const int c_in = a.mode == PS_MODE_2D ? 5 : 8, c_out = a.mode == PS_MODE_2D ? 5 : 10;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sg = sigmoidf_(in[c_in + k]);
            d[c_in + k] = (sg >= a.clip_lo && sg <= a.clip_hi) ? g[c_out + k] * sg * (1.0f - sg) : 0.0f;
        }
        if (a.mode == PS_MODE_2D) {
            d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; d[3] = g[3]; d[4] = g[4]; d[8] = 0.0f;
            dsc = g[2] + g[3];
        } else {
            d[4] = g[3]; d[5] = g[4]; d[6] = g[5]; d[7] = 0.0f;
            dsc = g[3] + g[4] + g[5];
            float gm[3] = { g[0], g[1], g[2] };
            float pc = 1.f, ps = 0.f, px, py, pz;
            if (a.pose) row_pose(a, i, pc, ps, px, py, pz);
            if (a.pose) { // means' = Rz m + p
                const float gx = pc * gm[0] + ps * gm[1], gy = -ps * gm[0] + pc * gm[1];
                gm[0] = gx; gm[1] = gy;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float th = tanhf(in[11 + k]);
                d[11 + k] = gm[k] * a.voxel2 * (1.0f - th * th);
            }
            if (!a.pose) {
                d[0] = g[6]; d[1] = g[7]; d[2] = g[8]; d[3] = g[9];
            } else {
                const float q[4] = { in[0], in[1], in[2], in[3] };
                QuatChain z;
                quat_chain(q, pc, ps, z);
                double dq[4] = { 0.0, 0.0, 0.0, 0.0 };
                if (!z.small) {
                    double V[4][4], t[4], lam;
                    jacobi4(z.K, V);
                    const int top = top_eigen(z.K, V, t, lam);
                    const double sgn = t[3] < 0.0 ? -1.0 : 1.0;
                    // cotangent of the eigenvector (x, y, z, w) from the cotangent of the output (w, x, y, z)
                    const double gt[4] = { sgn * g[7], sgn * g[8], sgn * g[9], sgn * g[6] };
                    // G = dL/dK (symmetric) = 1/2 sum_{j != top} c_j (v_j t^T + t v_j^T), c_j = (v_j . gt) / (lam - lam_j)
                    double G[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) G[r][cc] = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double dot = V[0][j] * gt[0] + V[1][j] * gt[1] + V[2][j] * gt[2] + V[3][j] * gt[3];
                        const double gap = lam - z.K[j][j];
                        const double cj = (j == top || gap == 0.0) ? 0.0 : 0.5 * dot / gap;
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) G[r][cc] += cj * (V[r][j] * t[cc] + t[r] * V[cc][j]);
                    }
                    // K / 3 is linear in m = Rz4 @ r
                    double dm[3][3];
                    dm[0][0] = (G[0][0] - G[1][1] - G[2][2] + G[3][3]) / 3.0;
                    dm[1][1] = (-G[0][0] + G[1][1] - G[2][2] + G[3][3]) / 3.0;
                    dm[2][2] = (-G[0][0] - G[1][1] + G[2][2] + G[3][3]) / 3.0;
                    dm[0][1] = 2.0 * (G[0][1] - G[2][3]) / 3.0; dm[1][0] = 2.0 * (G[0][1] + G[2][3]) / 3.0;
                    dm[0][2] = 2.0 * (G[0][2] + G[1][3]) / 3.0; dm[2][0] = 2.0 * (G[0][2] - G[1][3]) / 3.0;
                    dm[1][2] = 2.0 * (G[1][2] - G[0][3]) / 3.0; dm[2][1] = 2.0 * (G[1][2] + G[0][3]) / 3.0;
                    double dr[3][3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        dr[0][j] = (double)pc * dm[0][j] + (double)ps * dm[1][j];
                        dr[1][j] = -(double)ps * dm[0][j] + (double)pc * dm[1][j];
                        dr[2][j] = dm[2][j];
                    }
                    // r from the outer products o_ab = qs_a qs_b (entry (1,1) does not depend on q)
                    double dO[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) dO[r][cc] = 0.0;
                    dO[2][2] -= dr[0][0]; dO[3][3] -= dr[0][0];
                    dO[1][2] += dr[0][1]; dO[3][0] -= dr[0][1];
                    dO[1][3] += dr[0][2]; dO[2][0] += dr[0][2];
                    dO[1][2] += dr[1][0]; dO[3][0] -= dr[1][0];
                    dO[2][3] += dr[1][2]; dO[1][0] -= dr[1][2];
                    dO[1][3] += dr[2][0]; dO[2][0] -= dr[2][0];
                    dO[2][3] += dr[2][1]; dO[1][0] += dr[2][1];
                    dO[1][1] -= dr[2][2]; dO[2][2] -= dr[2][2];
                    double dqs[4], proj = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        dqs[r] = 0.0;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) dqs[r] += (dO[r][cc] + dO[cc][r]) * z.qs[cc];
                    }
                    // qs = q k, k = sqrt(2 / n): dq = k dqs - (dqs . q) (k / n) q
#pragma unroll
                    for (int r = 0; r < 4; ++r) proj += dqs[r] * (double)q[r];
#pragma unroll
                    for (int r = 0; r < 4; ++r) dq[r] = z.k * dqs[r] - proj * (z.k / z.n) * (double)q[r];
                }
                d[0] = (float)dq[0]; d[1] = (float)dq[1]; d[2] = (float)dq[2]; d[3] = (float)dq[3];
            }
        }
    }
#pragma unroll
    for (int dd = 16; dd >= 1; dd >>= 1) dsc += __shfl_xor_sync(0xffffffffu, dsc, dd);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = dsc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.0f;
        for (int w = 0; w < HEAD_THREADS / 32; ++w) tsum += s_part[w];
        if (tsum != 0.0f) atomicAdd(d_scale0, tsum);
    }
}

HeadArgs head_args(int mode, int n, float voxel_size, float pt, float clip_lo, float clip_hi, int pose, double angle, const float *p,
                   const float *poses, const int32_t *row_frame, int n_frames)
{
    HeadArgs a;
    a.n_frames = n_frames;
    a.mode = mode; a.n = n; a.pose = (mode == PS_MODE_3D && pose) ? 1 : 0;
    a.voxel2 = (float)(2.0 * (double)voxel_size);
    a.inv1mpt = (float)(1.0 / (1.0 - (double)pt));
    a.pt = pt; a.clip_lo = clip_lo; a.clip_hi = clip_hi;
    a.c = (float)cos(angle); a.s = (float)sin(angle);
    a.px = p ? p[0] : 0.f; a.py = p ? p[1] : 0.f; a.pz = p ? p[2] : 0.f;
    a.poses = (a.pose && poses && row_frame) ? poses : nullptr;
    a.row_frame = row_frame;
    return a;
}

} // namespace

int ps_launch_head_fwd(int mode, int n, const float *net_out, const float *probs, const float *grid, const float *scale0,
                       float voxel_size, float pt, float clip_lo, float clip_hi, int pose, double angle, const float *p_host,
                       const float *poses, const int32_t *row_frame, int n_frames, float *rows, cudaStream_t s)
{
    if (n <= 0) return 0;
    const HeadArgs a = head_args(mode, n, voxel_size, pt, clip_lo, clip_hi, pose, angle, p_host, poses, row_frame, n_frames);
    head_fwd_kernel<<<(n + HEAD_THREADS - 1) / HEAD_THREADS, HEAD_THREADS, 0, s>>>(a, net_out, probs, grid, scale0, rows);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_head_bwd(int mode, int n, const float *net_out, const float *probs, float voxel_size, float pt, float clip_lo,
                       float clip_hi, int pose, double angle, const float *poses, const int32_t *row_frame, int n_frames,
                       const float *d_rows, float *d_net, float *d_probs, float *d_scale0, cudaStream_t s)
{
    if (cudaMemsetAsync(d_scale0, 0, sizeof(float), s) != cudaSuccess) return -1;
    if (n <= 0) return 0;
    const HeadArgs a = head_args(mode, n, voxel_size, pt, clip_lo, clip_hi, pose, angle, nullptr, poses, row_frame, n_frames);
    head_bwd_kernel<<<(n + HEAD_THREADS - 1) / HEAD_THREADS, HEAD_THREADS, 0, s>>>(a, net_out, probs, d_rows, d_net, d_probs, d_scale0);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
