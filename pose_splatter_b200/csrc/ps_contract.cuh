// ps_contract.cuh -- arithmetic contract "PSM-1" of the renderer hot path.
//
// Everything that feeds a bit-exact output (sort keys, tile ranges, per-pixel contributor
// counts) is written here as an explicit sequence of IEEE fp32 round-to-nearest operations:
// + - * / sqrt and FMA only where psm_fma() is written, plus polynomial exp2 / log / sincos.
// On the device the wrappers map to the _rn intrinsics, which nvcc never contracts or
// reorders; on the host (tests/host_contract.cpp, built with -ffp-contract=off) they map to
// plain C.  The CPU oracle (oracle/ps_oracle.c) restates the same sequences independently.
//
// Reference anchors: adapter activations src/gaussian_renderer.py:183-193 (3D), :314-323
// (2D); 2D pair arithmetic :395-413; 3D core = gsplat 1.5.x fully_fused_projection /
// rasterize_to_pixels semantics (gsplat is absent from the reference tree).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD static inline
#endif

#define PS_TILE 16
#define PS_ALPHA_MIN (1.0f / 255.0f)
#define PS_ALPHA_MAX 0.999f
#define PS_T_STOP_3D 1e-4f
#define PS_TAU_2D 0x1p-28f
#define PS_TAU_INV_2D 0x1p28f
#define PS_T_STOP_2D 0x1p-20f
#define PS_RADIUS_MAX 1.0e9f

#ifdef __CUDA_ARCH__
PS_HD float psm_add(float a, float b) { return __fadd_rn(a, b); }
PS_HD float psm_sub(float a, float b) { return __fsub_rn(a, b); }
PS_HD float psm_mul(float a, float b) { return __fmul_rn(a, b); }
PS_HD float psm_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PS_HD float psm_div(float a, float b) { return __fdiv_rn(a, b); }
PS_HD float psm_sqrt(float a) { return __fsqrt_rn(a); }
PS_HD uint32_t psm_f2u(float f) { return __float_as_uint(f); }
PS_HD float psm_u2f(uint32_t u) { return __uint_as_float(u); }
#else
PS_HD float psm_add(float a, float b) { return a + b; }
PS_HD float psm_sub(float a, float b) { return a - b; }
PS_HD float psm_mul(float a, float b) { return a * b; }
PS_HD float psm_fma(float a, float b, float c) { return fmaf(a, b, c); }
PS_HD float psm_div(float a, float b) { return a / b; }
PS_HD float psm_sqrt(float a) { return sqrtf(a); }
PS_HD uint32_t psm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
PS_HD float psm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

// 2^t for |t| <= 125 (the caller guarantees the range): r = t - rint(t) via the 1.5*2^23 trick, degree-5 Horner.
PS_HD float psm_exp2_inrange(float t)
{
    const float magic = 12582912.0f;
    float z = psm_add(t, magic);
    float n = psm_sub(z, magic);
    float r = psm_sub(t, n);
    float p = 0x1.5f48c8p-10f;
    p = psm_fma(p, r, 0x1.3d107cp-7f);
    p = psm_fma(p, r, 0x1.c6aeccp-5f);
    p = psm_fma(p, r, 0x1.ebf906p-3f);
    p = psm_fma(p, r, 0x1.62e430p-1f);
    p = psm_fma(p, r, 1.0f);
    return psm_u2f(psm_f2u(p) + (psm_f2u(z) << 23));
}
// 2^t, t clamped to [-125, 125]
PS_HD float psm_exp2(float t) { return psm_exp2_inrange(fminf(fmaxf(t, -125.0f), 125.0f)); }
PS_HD float psm_exp(float x) { return psm_exp2(psm_mul(x, 0x1.715476p+0f)); }
PS_HD float psm_sigmoid(float x) { return psm_div(1.0f, psm_add(1.0f, psm_exp(-x))); }

PS_HD float psm_log(float x)
{
    uint32_t u = psm_f2u(x);
    int e = (int)(u >> 23) - 127;
    float m = psm_u2f((u & 0x007fffffu) | 0x3f800000u);
    if (m > 0x1.6a09e6p+0f) { m = psm_mul(m, 0.5f); e += 1; }
    float f = psm_sub(m, 1.0f);
    float s = psm_div(f, psm_add(2.0f, f));
    float s2 = psm_mul(s, s);
    float p = 0x1.c71c72p-4f;
    p = psm_fma(p, s2, 0x1.24924ap-3f);
    p = psm_fma(p, s2, 0x1.99999ap-3f);
    p = psm_fma(p, s2, 0x1.555556p-2f);
    p = psm_fma(p, s2, 1.0f);
    float lm = psm_mul(psm_mul(2.0f, s), p);
    float fe = (float)e;
    return psm_fma(fe, 0x1.62e4p-1f, psm_fma(fe, 0x1.7f7d1cp-20f, lm));
}

PS_HD void psm_sincos(float th, float *sn, float *cs)
{
    float k = rintf(psm_mul(th, 0x1.45f306p-1f));
    float r = psm_fma(-k, 0x1.92p+0f, th);
    r = psm_fma(-k, 0x1.fb4p-12f, r);
    r = psm_fma(-k, 0x1.4442d2p-24f, r);
    float z = psm_mul(r, r);
    float ps = -0x1.9943f2p-13f;
    ps = psm_fma(ps, z, 0x1.11073cp-7f);
    ps = psm_fma(ps, z, -0x1.555546p-3f);
    float sr = psm_fma(psm_mul(ps, z), r, r);
    float pc = 0x1.99eb9cp-16f;
    pc = psm_fma(pc, z, -0x1.6c0c34p-10f);
    pc = psm_fma(pc, z, 0x1.55554ap-5f);
    float cr = psm_fma(psm_mul(pc, z), z, psm_fma(-0.5f, z, 1.0f));
    int q = (int)psm_sub(k, psm_mul(4.0f, floorf(psm_mul(k, 0.25f))));
    float s_, c_;
    if (q == 0) { s_ = sr; c_ = cr; }
    else if (q == 1) { s_ = cr; c_ = -sr; }
    else if (q == 2) { s_ = -sr; c_ = -cr; }
    else { s_ = -cr; c_ = sr; }
    *sn = s_;
    *cs = c_;
}

PS_HD bool psm_finite(float x) { return psm_sub(x, x) == 0.0f; }

PS_HD int ps_tile_bits(int n_tiles)
{
    int b = 0;
    while ((1 << b) <= n_tiles) ++b;
    return b;
}

// ---------------------------------------------------------------------------------------
// Splat record: what the projection stage leaves per (view, Gaussian) for binning and
// rasterization.  In HBM it is three 16-byte words so that a tile rasterizer gathers it with
// three aligned 128-bit copies:
//   3D  rec0 = x, y, thr, opacity   rec1 = A/2, B, C/2, 0   rec2 = r, g, b, 0
//       (A/2, C/2: exact halvings, what the pair arithmetic uses; thr = log(255 * opacity), the
//        largest sigma that can pass alpha >= 1/255; rec0 + rec1 is all the culling needs; the radii
//        only shape the tile rectangle and the depth word goes to its own array)
//   2D  rec0 = u, v, L, opacity   rec1 = cos, sin, iax, iay   rec2 = r, g, b, 0
//       (L = log(opacity / tau): a pixel is in the footprint iff q <= L; the pixel rectangle of the ellipse's
//        bounding box only shapes the tile rectangle)
// tile rect: tx0, ty0, tx1, ty1 (exclusive max); culled <=> empty.
// ---------------------------------------------------------------------------------------
struct PsRecord {
    float r0[4];
    float r1[4];
    float r2[4];
    int tile[4];
    uint32_t low; // low word of the sort key: depth bits (3D) or row index (2D)
    float thr;    // 3D: log(255 * opacity)   2D: log(opacity / tau)
};

struct PsProj3dAux { // intermediates the projection backward re-uses
    float s[3], qa[4], qn_raw, qh[4], inv2, R[9], M[9], Sc[6], pc[3];
    float tx, ty, J00, J02, J11, J12;
    int clampx, clampy;
};

PS_HD void ps_record_clear(PsRecord *rec)
{
    for (int k = 0; k < 4; ++k) { rec->r0[k] = 0.0f; rec->r1[k] = 0.0f; rec->r2[k] = 0.0f; rec->tile[k] = 0; }
    rec->low = 0;
    rec->thr = 0.0f;
}

// The adapter of GaussianRenderer3D.render (src/gaussian_renderer.py:183-193): scales = exp(log_scales),
// quats = q / (|q| + 1e-8), colours = clamp(c, 0, 1), opacities = sigmoid(logit) -- the values the reference hands to
// gsplat.rendering.rasterization (:196-208).  activated != 0: the row already holds them (identity).
// Pinned by tests/golden/adapter3d_reference.npz (the reference's own lines, gsplat stubbed).
PS_HD void ps_adapter3d(const float *row, int activated, float *s, float *qa, float *qn_raw, float *rgb, float *o)
{
    for (int k = 0; k < 3; ++k) s[k] = activated ? row[3 + k] : psm_exp(row[3 + k]);
    float qw = row[6], qx = row[7], qy = row[8], qz = row[9];
    float n2 = psm_fma(qz, qz, psm_fma(qy, qy, psm_fma(qx, qx, psm_mul(qw, qw))));
    float qn = psm_sqrt(n2);
    *qn_raw = qn;
    float den = psm_add(qn, 1e-8f);
    if (activated) { qa[0] = qw; qa[1] = qx; qa[2] = qy; qa[3] = qz; }
    else {
        qa[0] = psm_div(qw, den); qa[1] = psm_div(qx, den);
        qa[2] = psm_div(qy, den); qa[3] = psm_div(qz, den);
    }
    for (int k = 0; k < 3; ++k) rgb[k] = activated ? row[10 + k] : fminf(fmaxf(row[10 + k], 0.0f), 1.0f);
    *o = activated ? row[13] : psm_sigmoid(row[13]);
}

// Vector-Jacobian product of the adapter (what autograd does through :183-193): v_act = gradient w.r.t. the activated
// values (means | scales | quats | colours | opacity, the layout gsplat's backward returns)  ->  out[14] = gradient w.r.t.
// the raw row.  Free-form fp32 (gradients are tolerance-checked).
PS_HD void ps_adapter3d_vjp(const float *row, const float *s, float qn_raw, float o, const float *v_act, float *out)
{
    for (int k = 0; k < 3; ++k) out[k] = v_act[k];
    for (int k = 0; k < 3; ++k) out[3 + k] = v_act[3 + k] * s[k];
    const float n = qn_raw, den = n + 1e-8f;
    const float dq = v_act[6] * row[6] + v_act[7] * row[7] + v_act[8] * row[8] + v_act[9] * row[9];
    for (int k = 0; k < 4; ++k) {
        float g0 = v_act[6 + k] / den;
        if (n > 0.0f) g0 -= dq / (den * den) * (row[6 + k] / n);
        out[6 + k] = g0;
    }
    for (int k = 0; k < 3; ++k) out[10 + k] = (row[10 + k] >= 0.0f && row[10 + k] <= 1.0f) ? v_act[10 + k] : 0.0f;
    out[13] = v_act[13] * o * (1.0f - o);
}

// The projection of one Gaussian for one camera = a camera-independent part (adapter activations, rotation, world
// covariance: ps_gauss3d) followed by a camera-dependent part (ps_view3d).  A frame's six cameras share the first; the
// kernels that loop over a frame's views compute it once per Gaussian.  Same operations in the same order as before the
// split: every bit-exact output is unchanged.
PS_HD void ps_gauss3d(const float *row, int activated, PsRecord *rec, PsProj3dAux *t, float *S /* [6] world covariance */)
{
    ps_record_clear(rec);
    float o;
    ps_adapter3d(row, activated, t->s, t->qa, &t->qn_raw, rec->r2, &o);
    rec->r1[3] = o;

    float a0 = t->qa[0], a1 = t->qa[1], a2 = t->qa[2], a3 = t->qa[3];
    float m2 = psm_fma(a3, a3, psm_fma(a2, a2, psm_fma(a1, a1, psm_mul(a0, a0))));
    float inv = psm_div(1.0f, psm_sqrt(m2));
    t->inv2 = inv;
    float w = psm_mul(a0, inv), x = psm_mul(a1, inv), y = psm_mul(a2, inv), z = psm_mul(a3, inv);
    t->qh[0] = w; t->qh[1] = x; t->qh[2] = y; t->qh[3] = z;
    float x2 = psm_mul(x, x), y2 = psm_mul(y, y), z2 = psm_mul(z, z);
    float xy = psm_mul(x, y), xz = psm_mul(x, z), yz = psm_mul(y, z);
    float wx = psm_mul(w, x), wy = psm_mul(w, y), wz = psm_mul(w, z);
    float *R = t->R;
    R[0] = psm_sub(1.0f, psm_mul(2.0f, psm_add(y2, z2)));
    R[1] = psm_mul(2.0f, psm_sub(xy, wz));
    R[2] = psm_mul(2.0f, psm_add(xz, wy));
    R[3] = psm_mul(2.0f, psm_add(xy, wz));
    R[4] = psm_sub(1.0f, psm_mul(2.0f, psm_add(x2, z2)));
    R[5] = psm_mul(2.0f, psm_sub(yz, wx));
    R[6] = psm_mul(2.0f, psm_sub(xz, wy));
    R[7] = psm_mul(2.0f, psm_add(yz, wx));
    R[8] = psm_sub(1.0f, psm_mul(2.0f, psm_add(x2, y2)));
    float *M = t->M;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = psm_mul(R[3 * i + j], t->s[j]);
    {
        int idx = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                S[idx++] = psm_fma(M[3 * i + 2], M[3 * j + 2],
                                   psm_fma(M[3 * i + 1], M[3 * j + 1], psm_mul(M[3 * i], M[3 * j])));
    }
}

// camera-dependent part: rec / t hold what ps_gauss3d left (colours, opacity, rotation ...).  Returns 1 if visible.
PS_HD int ps_view3d(const float *row, const float *S, const float *V, const float *K, int W, int H, float near_plane,
                    float far_plane, float radius_clip, float eps2d, PsRecord *rec, PsProj3dAux *t)
{
    const float o = rec->r1[3];
    // the camera-dependent fields start from zero for every camera (a Gaussian culled by this camera keeps only its
    // colours and opacity, whatever an earlier camera of the same frame left here)
    for (int k = 0; k < 4; ++k) { rec->r0[k] = 0.0f; rec->tile[k] = 0; }
    rec->r1[0] = rec->r1[1] = rec->r1[2] = 0.0f;
    rec->r2[3] = 0.0f;
    rec->low = 0;
    rec->thr = 0.0f;
    for (int i = 0; i < 3; ++i)
        t->pc[i] = psm_fma(V[4 * i + 2], row[2], psm_fma(V[4 * i + 1], row[1], psm_fma(V[4 * i], row[0], V[4 * i + 3])));
    float zc = t->pc[2];
    if (!(zc >= near_plane) || !(zc <= far_plane)) return 0;

    float Sf[9] = { S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5] };
    float Tm[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Tm[3 * i + j] = psm_fma(V[4 * i + 2], Sf[6 + j], psm_fma(V[4 * i + 1], Sf[3 + j], psm_mul(V[4 * i], Sf[j])));
    {
        int idx = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                t->Sc[idx++] = psm_fma(Tm[3 * i + 2], V[4 * j + 2],
                                       psm_fma(Tm[3 * i + 1], V[4 * j + 1], psm_mul(Tm[3 * i], V[4 * j])));
    }
    const float *Sc = t->Sc;

    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    float Wf = (float)W, Hf = (float)H;
    float tanx = psm_div(psm_mul(0.5f, Wf), fx), tany = psm_div(psm_mul(0.5f, Hf), fy);
    float limxp = psm_add(psm_div(psm_sub(Wf, cx), fx), psm_mul(0.3f, tanx));
    float limxn = psm_add(psm_div(cx, fx), psm_mul(0.3f, tanx));
    float limyp = psm_add(psm_div(psm_sub(Hf, cy), fy), psm_mul(0.3f, tany));
    float limyn = psm_add(psm_div(cy, fy), psm_mul(0.3f, tany));
    float rz = psm_div(1.0f, zc), rz2 = psm_mul(rz, rz);
    float xr = psm_mul(t->pc[0], rz), yr = psm_mul(t->pc[1], rz);
    float xcl = fminf(limxp, fmaxf(-limxn, xr)), ycl = fminf(limyp, fmaxf(-limyn, yr));
    t->clampx = (xcl != xr); t->clampy = (ycl != yr);
    float tx = psm_mul(zc, xcl), ty = psm_mul(zc, ycl);
    t->tx = tx; t->ty = ty;
    float J00 = psm_mul(fx, rz), J02 = psm_mul(-psm_mul(fx, tx), rz2);
    float J11 = psm_mul(fy, rz), J12 = psm_mul(-psm_mul(fy, ty), rz2);
    t->J00 = J00; t->J02 = J02; t->J11 = J11; t->J12 = J12;
    float a_0 = psm_fma(J02, Sc[2], psm_mul(J00, Sc[0]));
    float a_1 = psm_fma(J02, Sc[4], psm_mul(J00, Sc[1]));
    float a_2 = psm_fma(J02, Sc[5], psm_mul(J00, Sc[2]));
    float b_1 = psm_fma(J12, Sc[4], psm_mul(J11, Sc[3]));
    float b_2 = psm_fma(J12, Sc[5], psm_mul(J11, Sc[4]));
    float c00 = psm_fma(a_2, J02, psm_mul(a_0, J00));
    float c01 = psm_fma(a_2, J12, psm_mul(a_1, J11));
    float c11 = psm_fma(b_2, J12, psm_mul(b_1, J11));
    float mx = psm_fma(psm_mul(fx, t->pc[0]), rz, cx), my = psm_fma(psm_mul(fy, t->pc[1]), rz, cy);
    c00 = psm_add(c00, eps2d); c11 = psm_add(c11, eps2d);
    float det = psm_fma(c00, c11, -psm_mul(c01, c01));
    if (!(det > 0.0f)) return 0;
    float cA = psm_div(c11, det), cB = psm_div(-c01, det), cC = psm_div(c00, det);

    if (!(o >= PS_ALPHA_MIN)) return 0;
    float thr = psm_log(psm_mul(o, 255.0f));
    float ext = fminf(3.33f, psm_sqrt(psm_mul(2.0f, thr)));
    float bh = psm_mul(0.5f, psm_add(c00, c11));
    float v1 = psm_add(bh, psm_sqrt(fmaxf(0.01f, psm_fma(bh, bh, -det))));
    float r1 = psm_mul(ext, psm_sqrt(v1));
    float rx = ceilf(fminf(psm_mul(ext, psm_sqrt(c00)), r1));
    float ry = ceilf(fminf(psm_mul(ext, psm_sqrt(c11)), r1));
    if (!psm_finite(mx) || !psm_finite(my) || !psm_finite(cA) || !psm_finite(cB) || !psm_finite(cC) ||
        !(rx == rx) || !(ry == ry))
        return 0;
    rx = fminf(rx, PS_RADIUS_MAX); ry = fminf(ry, PS_RADIUS_MAX);
    if (rx <= radius_clip && ry <= radius_clip) return 0;
    if (psm_add(mx, rx) <= 0.0f || psm_sub(mx, rx) >= Wf || psm_add(my, ry) <= 0.0f || psm_sub(my, ry) >= Hf) return 0;

    rec->r0[0] = mx; rec->r0[1] = my; rec->r0[2] = rx; rec->r0[3] = ry;
    rec->r1[0] = cA; rec->r1[1] = cB; rec->r1[2] = cC; rec->r1[3] = o;
    rec->r2[3] = zc;
    rec->low = psm_f2u(zc);
    rec->thr = thr;

    int tw = (W + PS_TILE - 1) / PS_TILE, th = (H + PS_TILE - 1) / PS_TILE;
    float txc = psm_mul(mx, 0.0625f), tyc = psm_mul(my, 0.0625f);
    float trx = psm_mul(rx, 0.0625f), try_ = psm_mul(ry, 0.0625f);
    float fx0 = fminf(fmaxf(floorf(psm_sub(txc, trx)), 0.0f), (float)tw);
    float fx1 = fminf(fmaxf(ceilf(psm_add(txc, trx)), 0.0f), (float)tw);
    float fy0 = fminf(fmaxf(floorf(psm_sub(tyc, try_)), 0.0f), (float)th);
    float fy1 = fminf(fmaxf(ceilf(psm_add(tyc, try_)), 0.0f), (float)th);
    rec->tile[0] = (int)fx0; rec->tile[1] = (int)fy0; rec->tile[2] = (int)fx1; rec->tile[3] = (int)fy1;
    if (rec->tile[2] <= rec->tile[0] || rec->tile[3] <= rec->tile[1]) {
        rec->tile[2] = rec->tile[0]; rec->tile[3] = rec->tile[1];
    }
    return 1;
}

// Adapter activations + EWA projection of one Gaussian for one camera. Returns 1 if visible.
// activated != 0: the row already holds scales / quaternion / colours / opacity as gsplat's rasterization() takes
// them (the legacy PoseSplatter.splat call, src/model.py:342-361): no exp, no q/(|q|+1e-8), no clamp, no sigmoid.
PS_HD int ps_project3d(const float *row, const float *V, const float *K, int W, int H, float near_plane,
                       float far_plane, float radius_clip, float eps2d, PsRecord *rec, PsProj3dAux *t, int activated = 0)
{
    float S[6];
    ps_gauss3d(row, activated, rec, t, S);
    return ps_view3d(row, S, V, K, W, H, near_plane, far_plane, radius_clip, eps2d, rec, t);
}

// 2D activations + binning extent. Returns 1 if listed anywhere.
// PsRecord (contract level): r0 = u, v, bits(x0|y0<<16), bits(x1|y1<<16); r1 = cos, sin, iax, iay; r2 = r, g, b, o.
PS_HD int ps_project2d(const float *row, uint32_t row_index, int W, int H, PsRecord *rec)
{
    ps_record_clear(rec);
    rec->low = row_index;
    float u = row[0], v = row[1];
    float sx = psm_exp(row[2]), sy = psm_exp(row[3]);
    float sn, cs;
    psm_sincos(row[4], &sn, &cs);
    for (int k = 0; k < 3; ++k) rec->r2[k] = fminf(fmaxf(row[5 + k], 0.0f), 1.0f);
    float o = psm_sigmoid(row[8]);
    float ax = psm_add(psm_mul(2.0f, psm_mul(sx, sx)), 1e-8f);
    float ay = psm_add(psm_mul(2.0f, psm_mul(sy, sy)), 1e-8f);
    float iax = psm_div(1.0f, ax), iay = psm_div(1.0f, ay);
    rec->r2[3] = o;
    if (!(o > PS_TAU_2D)) return 0;
    float chk = psm_add(psm_add(psm_add(psm_add(psm_add(u, v), iax), iay), sn), cs);
    if (!psm_finite(chk)) return 0;
    // g >= tau <=> q <= L = ln(o / tau): an ellipse, listed on the tiles met by its bounding box (+1 px of slack)
    float L = psm_log(psm_mul(o, PS_TAU_INV_2D));
    float cc = psm_mul(cs, cs), ss = psm_mul(sn, sn);
    float mxx = psm_fma(ay, ss, psm_mul(ax, cc)), myy = psm_fma(ay, cc, psm_mul(ax, ss));
    float hx = psm_add(ceilf(psm_sqrt(psm_mul(L, mxx))), 1.0f), hy = psm_add(ceilf(psm_sqrt(psm_mul(L, myy))), 1.0f);
    if (!(hx == hx) || !(hy == hy)) return 0;
    hx = fminf(hx, PS_RADIUS_MAX); hy = fminf(hy, PS_RADIUS_MAX);
    float x0 = fmaxf(ceilf(psm_sub(u, hx)), 0.0f), x1 = fminf(floorf(psm_add(u, hx)), (float)(W - 1));
    float y0 = fmaxf(ceilf(psm_sub(v, hy)), 0.0f), y1 = fminf(floorf(psm_add(v, hy)), (float)(H - 1));
    if (!(x0 <= x1) || !(y0 <= y1)) return 0;
    rec->thr = L;
    int ix0 = (int)x0, iy0 = (int)y0, ix1 = (int)x1, iy1 = (int)y1;
    rec->r0[0] = u; rec->r0[1] = v;
    rec->r0[2] = psm_u2f((uint32_t)ix0 | ((uint32_t)iy0 << 16));
    rec->r0[3] = psm_u2f((uint32_t)ix1 | ((uint32_t)iy1 << 16));
    rec->r1[0] = cs; rec->r1[1] = sn; rec->r1[2] = iax; rec->r1[3] = iay;
    rec->tile[0] = ix0 / PS_TILE; rec->tile[1] = iy0 / PS_TILE;
    rec->tile[2] = ix1 / PS_TILE + 1; rec->tile[3] = iy1 / PS_TILE + 1;
    return 1;
}

// ---------------------------------------------------------------------------------------
// Per (pixel, Gaussian) pair arithmetic.
// ---------------------------------------------------------------------------------------
// 3D: sigma = 1/2 (A dx^2 + C dy^2) + B dx dy with d = mean2d - pixel centre (px+0.5, py+0.5);
// hA = A/2 and hC = C/2 (exact) come from the record
PS_HD float ps_sigma3d(float gx, float gy, float hA, float B, float hC, float px, float py, float *dx_, float *dy_)
{
    float dx = psm_sub(gx, px), dy = psm_sub(gy, py);
    float uu = psm_fma(B, dy, psm_mul(hA, dx));
    float s = psm_mul(uu, dx);
    float wv = psm_mul(hC, dy);
    s = psm_fma(wv, dy, s);
    *dx_ = dx; *dy_ = dy;
    return s;
}

// 2D: q = dxr^2 * iax + dyr^2 * iay with the rotation convention of src/gaussian_renderer.py:401-402
PS_HD float ps_q2d(float u, float v, float cs, float sn, float iax, float iay, float x, float y, float *dxr_, float *dyr_)
{
    float dx = psm_sub(x, u), dy = psm_sub(y, v);
    float dxr = psm_fma(sn, dy, psm_mul(cs, dx));
    float dyr = psm_fma(cs, dy, psm_mul(-sn, dx));
    float q = psm_fma(psm_mul(dyr, dyr), iay, psm_mul(psm_mul(dxr, dxr), iax));
    *dxr_ = dxr; *dyr_ = dyr;
    return q;
}
