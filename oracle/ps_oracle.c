/*
 * ps_oracle.c -- CPU ORACLE for the pose-splatter renderer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (pose_splatter_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates (citations are path:line under the reference tree):
 *   2D mode  src/gaussian_renderer.py:291-334 (wrapper: slicing, exp / clamp / sigmoid,
 *            background composite) and :351-427 (dense arithmetic: integer pixel centres,
 *            rotation convention :401-402, 2*s^2+1e-8 denominators :408-410, additive alpha
 *            :416-425).  The reference is dense O(N*H*W); the oracle restates the SAME sum
 *            restricted by the binning definition of DESIGN.md section 5, whose
 *            error against the dense sum is bounded by the tau budget.  PARITY PINNED: the
 *            output is checked against fixtures produced by running the reference class
 *            itself (tests/golden/make_golden.py).
 *   3D mode  adapter src/gaussian_renderer.py:175-211 literally (exp, q/(|q|+1e-8), clamp,
 *            sigmoid), then gsplat.rendering.rasterization(packed=False, classic, RGB) as
 *            called at :196-208.  gsplat (requirements.txt:10 "gsplat>=0.1.0", un-vendored,
 *            not installable offline) is ABSENT from the reference tree; its 1.5.x algorithm
 *            is restated from its published structure (fully_fused_projection, isect_tiles,
 *            isect_offset_encode, rasterize_to_pixels fwd/bwd; rules in SURVEY 8c-c5).
 *            PARITY UNPINNED for 3D: the reference's own tests pin only shapes
 *            (tests/test_gaussian_renderer.py:207-229).  The oracle is cross-checked against
 *            an independent fp64 torch restatement with autograd (oracle/ref3d_torch.py).
 *
 * Arithmetic contract ("PSM-1"): fp32, round-to-nearest, IEEE + - * / sqrt,
 * FMA only where fmaf() is written, and deterministic exp2 / log / sincos polynomials so
 * that integer outputs (sort keys, tile ranges, contributor counts) are bit-reproducible
 * on CPU and GPU.  Build with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_TILE 16
#define ORA_ALPHA_MIN (1.0f / 255.0f)
#define ORA_ALPHA_MAX 0.999f
#define ORA_T_STOP_3D 1e-4f
#define ORA_TAU_2D 0x1p-28f   /* 2D binning error budget per Gaussian */
#define ORA_TAU_INV_2D 0x1p28f
#define ORA_T_STOP_2D 0x1p-20f /* 2D early stop: remaining light <= 9.6e-7 */
#define ORA_RADIUS_MAX 1.0e9f

/* ------------------------------------------------------------------------------------ */
/* PSM-1 deterministic math                                                             */
/* ------------------------------------------------------------------------------------ */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* 2^t for t clamped to [-125, 125]; degree-5 polynomial on r = t - rint(t). */
static inline float d_exp2(float t)
{
    t = fminf(fmaxf(t, -125.0f), 125.0f);
    const float magic = 12582912.0f; /* 1.5 * 2^23 */
    float z = t + magic;
    float n = z - magic;
    float r = t - n;
    float p = 0x1.5f48c8p-10f;
    p = fmaf(p, r, 0x1.3d107cp-7f);
    p = fmaf(p, r, 0x1.c6aeccp-5f);
    p = fmaf(p, r, 0x1.ebf906p-3f);
    p = fmaf(p, r, 0x1.62e430p-1f);
    p = fmaf(p, r, 1.0f);
    return u2f(f2u(p) + (f2u(z) << 23));
}
static inline float d_exp(float x) { return d_exp2(x * 0x1.715476p+0f); }
static inline float d_sigmoid(float x) { return 1.0f / (1.0f + d_exp(-x)); }

/* natural log for finite x > 0 (normal range); 2*atanh series on m in [sqrt(.5), sqrt(2)). */
static inline float d_log(float x)
{
    uint32_t u = f2u(x);
    int e = (int)(u >> 23) - 127;
    uint32_t mb = (u & 0x007fffffu) | 0x3f800000u;
    float m = u2f(mb);
    if (m > 0x1.6a09e6p+0f) { m = m * 0.5f; e += 1; }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float s2 = s * s;
    float p = 0x1.c71c72p-4f;            /* 1/9 */
    p = fmaf(p, s2, 0x1.24924ap-3f);     /* 1/7 */
    p = fmaf(p, s2, 0x1.99999ap-3f);     /* 1/5 */
    p = fmaf(p, s2, 0x1.555556p-2f);     /* 1/3 */
    p = fmaf(p, s2, 1.0f);
    float lm = (2.0f * s) * p;
    float fe = (float)e;
    return fmaf(fe, 0x1.62e4p-1f, fmaf(fe, 0x1.7f7d1cp-20f, lm));
}

/* sin, cos with 3-term Cody-Waite reduction by pi/2 and cephes-style polynomials. */
static inline void d_sincos(float th, float *sn, float *cs)
{
    float k = rintf(th * 0x1.45f306p-1f);       /* 2/pi */
    float r = fmaf(-k, 0x1.92p+0f, th);          /* 1.5703125 */
    r = fmaf(-k, 0x1.fb4p-12f, r);               /* 4.837512969970703125e-4 */
    r = fmaf(-k, 0x1.4442d2p-24f, r);            /* 7.54978995489188216e-8 */
    float z = r * r;
    float ps = -0x1.9943f2p-13f;                  /* -1.9515295891e-4 */
    ps = fmaf(ps, z, 0x1.11073cp-7f);            /*  8.3321608736e-3 */
    ps = fmaf(ps, z, -0x1.555546p-3f);           /* -1.6666654611e-1 */
    float sr = fmaf(ps * z, r, r);
    float pc = 0x1.99eb9cp-16f;                  /*  2.443315711809948e-5 */
    pc = fmaf(pc, z, -0x1.6c0c34p-10f);          /* -1.388731625493765e-3 */
    pc = fmaf(pc, z, 0x1.55554ap-5f);            /*  4.166664568298827e-2 */
    float cr = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    int q = (int)(k - 4.0f * floorf(k * 0.25f)); /* k mod 4, exact for integer-valued floats */
    float s_, c_;
    switch (q) {
        case 0: s_ = sr; c_ = cr; break;
        case 1: s_ = cr; c_ = -sr; break;
        case 2: s_ = -sr; c_ = -cr; break;
        default: s_ = -cr; c_ = sr; break;
    }
    *sn = s_;
    *cs = c_;
}

/* exported so the tests can probe the contract functions directly */
void ora_math_probe(const float *x, int n, float *o_exp, float *o_log, float *o_sig, float *o_sin, float *o_cos)
{
    for (int i = 0; i < n; ++i) {
        o_exp[i] = d_exp(x[i]);
        o_log[i] = d_log(fabsf(x[i]) + 1e-30f);
        o_sig[i] = d_sigmoid(x[i]);
        d_sincos(x[i], &o_sin[i], &o_cos[i]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Per-view splat table shared by both modes                                            */
/* ------------------------------------------------------------------------------------ */
/*
 * SoA per (view, Gaussian).  geom is 8 floats:
 *   3D: x, y (mean2d px), A, B, C (conic), opacity, depth, 0
 *   2D: u, v, cos, sin, iax, iay, opacity, 0
 * rect is 4 ints: 3D = radius_x, radius_y, 0, 0 ; 2D = x0, y0, x1, y1 (inclusive pixel rect)
 * tile_rect is tx0, ty0, tx1, ty1 (exclusive max); a culled Gaussian has tx1 == tx0.
 * low is the low word of the sort key (3D: depth bits, 2D: row index).
 */

static inline int tile_bits_for(int n_tiles)
{
    int b = 0;
    while ((1 << b) <= n_tiles) ++b; /* floor(log2(n_tiles)) + 1 */
    return b;
}
int ora_tile_bits(int W, int H)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    return tile_bits_for(tw * th);
}

/* ------------------------------------------------------------------------------------ */
/* 3D: adapter activations + gsplat fully_fused_projection (EWA)                        */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    float s[3];      /* activated scales */
    float qa[4];     /* quaternion after the adapter's q/(|q|+1e-8) */
    float qn_raw;    /* |q| of the raw row */
    float qh[4];     /* unit quaternion used for R (gsplat normalises again) */
    float inv2;      /* 1/|qa| */
    float R[9], M[9], S[6];          /* world covariance, upper triangle 00 01 02 11 12 22 */
    float pc[3];                     /* camera-space mean */
    float Sc[6];                     /* camera-space covariance */
    float rz, tx, ty; int clampx, clampy;
    float J00, J02, J11, J12;
    float c00, c01, c11, det;        /* blurred 2D covariance */
} ora_proj3d_tmp;

/* The adapter of GaussianRenderer3D.render, src/gaussian_renderer.py:183-193: scales = exp(log_scales),
 * quats = q / (|q| + 1e-8), colours = clamp(c, 0, 1), opacities = sigmoid(logit).  activated != 0: identity. */
static void adapter3d_one(const float *row, int activated, float *s, float *qa, float *qn_raw, float *rgb, float *o)
{
    for (int k = 0; k < 3; ++k) s[k] = activated ? row[3 + k] : d_exp(row[3 + k]);
    float qw = row[6], qx = row[7], qy = row[8], qz = row[9];
    float n2 = fmaf(qz, qz, fmaf(qy, qy, fmaf(qx, qx, qw * qw)));
    float qn = sqrtf(n2);
    *qn_raw = qn;
    float den = qn + 1e-8f;
    if (activated) { qa[0] = qw; qa[1] = qx; qa[2] = qy; qa[3] = qz; }
    else { qa[0] = qw / den; qa[1] = qx / den; qa[2] = qy / den; qa[3] = qz / den; }
    for (int k = 0; k < 3; ++k) rgb[k] = activated ? row[10 + k] : fminf(fmaxf(row[10 + k], 0.0f), 1.0f);
    *o = activated ? row[13] : d_sigmoid(row[13]);
}

/* vector-Jacobian product of the adapter: v_act = dL/d(means | scales | quats | colours | opacity) as
 * gsplat.rendering.rasterization receives them  ->  d += dL/d(raw row) (what autograd does through :183-193) */
static void adapter3d_vjp_one(const float *row, const float *s, float qn_raw, float o, const double *v_act, double *d)
{
    for (int k = 0; k < 3; ++k) d[k] += v_act[k];
    for (int k = 0; k < 3; ++k) d[3 + k] += v_act[3 + k] * s[k]; /* scale = exp(log_scale) */
    /* qa = q / (|q| + 1e-8) */
    double n = qn_raw, den = n + 1e-8;
    double dq = v_act[6] * row[6] + v_act[7] * row[7] + v_act[8] * row[8] + v_act[9] * row[9];
    for (int k = 0; k < 4; ++k) {
        double g0 = v_act[6 + k] / den;
        if (n > 0.0) g0 -= dq / (den * den) * (row[6 + k] / n);
        d[6 + k] += g0;
    }
    for (int k = 0; k < 3; ++k)
        if (row[10 + k] >= 0.0f && row[10 + k] <= 1.0f) d[10 + k] += v_act[10 + k];
    d[13] += v_act[13] * (double)o * (1.0 - (double)o);
}

/* Adapter stage alone (parity fixture tests/golden/adapter3d_reference.npz): act [N,14] and, if v_act is given,
 * d_params [N,14] += J^T v_act */
void ora3d_adapter(const float *params, int N, float *act, const double *v_act, double *d_params)
{
    for (int i = 0; i < N; ++i) {
        const float *row = params + 14 * (size_t)i;
        float s[3], qa[4], qn, rgb[3], o;
        adapter3d_one(row, 0, s, qa, &qn, rgb, &o);
        float *a = act + 14 * (size_t)i;
        for (int k = 0; k < 3; ++k) { a[k] = row[k]; a[3 + k] = s[k]; a[10 + k] = rgb[k]; }
        for (int k = 0; k < 4; ++k) a[6 + k] = qa[k];
        a[13] = o;
        if (v_act) adapter3d_vjp_one(row, s, qn, o, v_act + 14 * (size_t)i, d_params + 14 * (size_t)i);
    }
}

static int project3d_one(const float *row, const float *V, const float *K, int W, int H,
                         float near_plane, float far_plane, float radius_clip, float eps2d,
                         float *geom, float *rgb, int *rect, int *tile_rect, uint32_t *low,
                         ora_proj3d_tmp *t, int activated)
{
    /* adapter: src/gaussian_renderer.py:183-193.  activated != 0: the row already holds scales, quaternion,
     * colours and opacity as gsplat.rendering.rasterization receives them (src/model.py:342-361): no exp, no
     * q/(|q|+1e-8), no clamp, no sigmoid (gsplat still normalises the quaternion itself) */
    float o;
    adapter3d_one(row, activated, t->s, t->qa, &t->qn_raw, rgb, &o);

    for (int k = 0; k < 8; ++k) geom[k] = 0.0f;
    rect[0] = rect[1] = rect[2] = rect[3] = 0;
    tile_rect[0] = tile_rect[1] = tile_rect[2] = tile_rect[3] = 0;
    *low = 0;
    geom[5] = o;

    /* gsplat quat_to_rotmat: normalise again (rsqrt upstream; IEEE 1/sqrt here) */
    float a0 = t->qa[0], a1 = t->qa[1], a2 = t->qa[2], a3 = t->qa[3];
    float m2 = fmaf(a3, a3, fmaf(a2, a2, fmaf(a1, a1, a0 * a0)));
    float inv = 1.0f / sqrtf(m2);
    t->inv2 = inv;
    float w = a0 * inv, x = a1 * inv, y = a2 * inv, z = a3 * inv;
    t->qh[0] = w; t->qh[1] = x; t->qh[2] = y; t->qh[3] = z;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    float *R = t->R;
    R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
    R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
    R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
    float *M = t->M;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * t->s[j];
    /* Sigma = M M^T */
    float *S = t->S;
    {
        int idx = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                S[idx++] = fmaf(M[3 * i + 2], M[3 * j + 2], fmaf(M[3 * i + 1], M[3 * j + 1], M[3 * i] * M[3 * j]));
    }
    /* world -> camera (V row-major 4x4, OpenCV convention) */
    const float *p = row;
    for (int i = 0; i < 3; ++i)
        t->pc[i] = fmaf(V[4 * i + 2], p[2], fmaf(V[4 * i + 1], p[1], fmaf(V[4 * i], p[0], V[4 * i + 3])));
    float zc = t->pc[2];
    if (!(zc >= near_plane) || !(zc <= far_plane)) return 0;

    /* Sigma_c = Rwc Sigma Rwc^T */
    float Sf[9] = { S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5] };
    float Tm[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Tm[3 * i + j] = fmaf(V[4 * i + 2], Sf[6 + j], fmaf(V[4 * i + 1], Sf[3 + j], V[4 * i] * Sf[j]));
    {
        int idx = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                t->Sc[idx++] = fmaf(Tm[3 * i + 2], V[4 * j + 2], fmaf(Tm[3 * i + 1], V[4 * j + 1], Tm[3 * i] * V[4 * j]));
    }
    const float *Sc = t->Sc; /* 00 01 02 11 12 22 */

    /* gsplat persp_proj */
    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    float Wf = (float)W, Hf = (float)H;
    float tanx = (0.5f * Wf) / fx, tany = (0.5f * Hf) / fy;
    float limxp = (Wf - cx) / fx + 0.3f * tanx, limxn = cx / fx + 0.3f * tanx;
    float limyp = (Hf - cy) / fy + 0.3f * tany, limyn = cy / fy + 0.3f * tany;
    float rz = 1.0f / zc, rz2 = rz * rz;
    float xr = t->pc[0] * rz, yr = t->pc[1] * rz;
    float xcl = fminf(limxp, fmaxf(-limxn, xr)), ycl = fminf(limyp, fmaxf(-limyn, yr));
    t->clampx = (xcl != xr); t->clampy = (ycl != yr);
    float tx = zc * xcl, ty = zc * ycl;
    t->rz = rz; t->tx = tx; t->ty = ty;
    float J00 = fx * rz, J02 = -(fx * tx) * rz2, J11 = fy * rz, J12 = -(fy * ty) * rz2;
    t->J00 = J00; t->J02 = J02; t->J11 = J11; t->J12 = J12;
    /* rows of J*Sc */
    float a_0 = fmaf(J02, Sc[2], J00 * Sc[0]);
    float a_1 = fmaf(J02, Sc[4], J00 * Sc[1]);
    float a_2 = fmaf(J02, Sc[5], J00 * Sc[2]);
    float b_1 = fmaf(J12, Sc[4], J11 * Sc[3]);
    float b_2 = fmaf(J12, Sc[5], J11 * Sc[4]);
    float c00 = fmaf(a_2, J02, a_0 * J00);
    float c01 = fmaf(a_2, J12, a_1 * J11);
    float c11 = fmaf(b_2, J12, b_1 * J11);
    float mx = fmaf(fx * t->pc[0], rz, cx), my = fmaf(fy * t->pc[1], rz, cy);
    c00 = c00 + eps2d; c11 = c11 + eps2d;
    float det = fmaf(c00, c11, -(c01 * c01));
    t->c00 = c00; t->c01 = c01; t->c11 = c11; t->det = det;
    if (!(det > 0.0f)) return 0;
    float cA = c11 / det, cB = -c01 / det, cC = c00 / det;

    /* gsplat 1.5 opacity-aware rectangular radius */
    if (!(o >= ORA_ALPHA_MIN)) return 0;
    float ext = fminf(3.33f, sqrtf(2.0f * d_log(o * 255.0f)));
    float bh = 0.5f * (c00 + c11);
    float v1 = bh + sqrtf(fmaxf(0.01f, fmaf(bh, bh, -det)));
    float r1 = ext * sqrtf(v1);
    float rx = ceilf(fminf(ext * sqrtf(c00), r1));
    float ry = ceilf(fminf(ext * sqrtf(c11), r1));
    if (!(mx - mx == 0.0f) || !(my - my == 0.0f) || !(cA - cA == 0.0f) || !(cB - cB == 0.0f) ||
        !(cC - cC == 0.0f) || !(rx == rx) || !(ry == ry))
        return 0; /* non-finite -> culled (contract; upstream would propagate NaN) */
    rx = fminf(rx, ORA_RADIUS_MAX); ry = fminf(ry, ORA_RADIUS_MAX);
    if (rx <= radius_clip && ry <= radius_clip) return 0;
    if (mx + rx <= 0.0f || mx - rx >= Wf || my + ry <= 0.0f || my - ry >= Hf) return 0;

    geom[0] = mx; geom[1] = my; geom[2] = cA; geom[3] = cB; geom[4] = cC; geom[5] = o; geom[6] = zc;
    rect[0] = (int)rx; rect[1] = (int)ry;
    *low = f2u(zc);

    /* gsplat isect_tiles: tile rectangle, exclusive max, clamped to the grid */
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    float txc = mx * 0.0625f, tyc = my * 0.0625f, trx = rx * 0.0625f, try_ = ry * 0.0625f;
    float x0 = fminf(fmaxf(floorf(txc - trx), 0.0f), (float)tw);
    float x1 = fminf(fmaxf(ceilf(txc + trx), 0.0f), (float)tw);
    float y0 = fminf(fmaxf(floorf(tyc - try_), 0.0f), (float)th);
    float y1 = fminf(fmaxf(ceilf(tyc + try_), 0.0f), (float)th);
    tile_rect[0] = (int)x0; tile_rect[1] = (int)y0; tile_rect[2] = (int)x1; tile_rect[3] = (int)y1;
    if (tile_rect[2] <= tile_rect[0] || tile_rect[3] <= tile_rect[1]) {
        tile_rect[2] = tile_rect[0]; tile_rect[3] = tile_rect[1];
    }
    return 1;
}

/* params [N,14] row-major; V [16]; K [9].  Outputs per Gaussian (see table comment). */
void ora3d_project(const float *params, int N, const float *V, const float *K, int W, int H,
                   float near_plane, float far_plane, float radius_clip, float eps2d,
                   float *geom, float *rgb, int *rect, int *tile_rect, uint32_t *low, int *tiles_touched, int activated)
{
    ora_proj3d_tmp t;
    for (int i = 0; i < N; ++i) {
        project3d_one(params + 14 * (size_t)i, V, K, W, H, near_plane, far_plane, radius_clip, eps2d,
                      geom + 8 * (size_t)i, rgb + 3 * (size_t)i, rect + 4 * (size_t)i,
                      tile_rect + 4 * (size_t)i, low + i, &t, activated);
        const int *tr = tile_rect + 4 * (size_t)i;
        tiles_touched[i] = (tr[2] - tr[0]) * (tr[3] - tr[1]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* 2D: activations (src/gaussian_renderer.py:314-323) + binning extent    */
/* ------------------------------------------------------------------------------------ */
void ora2d_project(const float *params, int N, int W, int H,
                   float *geom, float *rgb, int *rect, int *tile_rect, uint32_t *low, int *tiles_touched)
{
    for (int i = 0; i < N; ++i) {
        const float *row = params + 9 * (size_t)i;
        float *g = geom + 8 * (size_t)i;
        int *rc = rect + 4 * (size_t)i, *tr = tile_rect + 4 * (size_t)i;
        for (int k = 0; k < 8; ++k) g[k] = 0.0f;
        rc[0] = rc[1] = rc[2] = rc[3] = 0;
        tr[0] = tr[1] = tr[2] = tr[3] = 0;
        low[i] = (uint32_t)i;
        tiles_touched[i] = 0;
        float u = row[0], v = row[1];
        float sx = d_exp(row[2]), sy = d_exp(row[3]);
        float sn, cs;
        d_sincos(row[4], &sn, &cs);
        for (int k = 0; k < 3; ++k) rgb[3 * (size_t)i + k] = fminf(fmaxf(row[5 + k], 0.0f), 1.0f);
        float o = d_sigmoid(row[8]);
        float ax = (2.0f * (sx * sx)) + 1e-8f, ay = (2.0f * (sy * sy)) + 1e-8f; /* :409 */
        float iax = 1.0f / ax, iay = 1.0f / ay;
        g[6] = o;
        if (!(o > ORA_TAU_2D)) continue;
        float chk = ((((u + v) + iax) + iay) + sn) + cs;
        if (!(chk - chk == 0.0f)) continue; /* non-finite -> culled */
        /* g >= tau  <=>  q <= L = ln(o/tau): an ellipse.  Listed on the tiles met by its bounding box,
         * half-widths sqrt(L * (M^-1)_xx), sqrt(L * (M^-1)_yy) with M^-1 = R^T diag(ax, ay) R, plus one pixel
         * of slack so that no pixel passing the fp32 test q <= L can fall outside the listed tiles */
        float L = d_log(o * ORA_TAU_INV_2D);
        float cc = cs * cs, ss = sn * sn;
        float mxx = fmaf(ay, ss, ax * cc), myy = fmaf(ay, cc, ax * ss);
        float hx = ceilf(sqrtf(L * mxx)) + 1.0f, hy = ceilf(sqrtf(L * myy)) + 1.0f;
        if (!(hx == hx) || !(hy == hy)) continue;
        hx = fminf(hx, ORA_RADIUS_MAX); hy = fminf(hy, ORA_RADIUS_MAX);
        float x0 = fmaxf(ceilf(u - hx), 0.0f), x1 = fminf(floorf(u + hx), (float)(W - 1));
        float y0 = fmaxf(ceilf(v - hy), 0.0f), y1 = fminf(floorf(v + hy), (float)(H - 1));
        if (!(x0 <= x1) || !(y0 <= y1)) continue;
        g[7] = L;
        g[0] = u; g[1] = v; g[2] = cs; g[3] = sn; g[4] = iax; g[5] = iay;
        rc[0] = (int)x0; rc[1] = (int)y0; rc[2] = (int)x1; rc[3] = (int)y1;
        tr[0] = rc[0] / ORA_TILE; tr[1] = rc[1] / ORA_TILE;
        tr[2] = rc[2] / ORA_TILE + 1; tr[3] = rc[3] / ORA_TILE + 1;
        tiles_touched[i] = (tr[2] - tr[0]) * (tr[3] - tr[1]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Tile binning: emit (key, value), stable sort, tile ranges (gsplat isect_tiles /       */
/* isect_offset_encode).  key = view<<(32+tile_bits) | tile<<32 | low.     */
/* ------------------------------------------------------------------------------------ */
void ora_emit(int N, int W, int view, int view_stride_n, const int *tile_rect, const uint32_t *low,
              int tile_bits, int64_t *keys, int32_t *vals)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE;
    size_t m = 0;
    for (int i = 0; i < N; ++i) {
        const int *tr = tile_rect + 4 * (size_t)i;
        for (int ty = tr[1]; ty < tr[3]; ++ty)
            for (int tx = tr[0]; tx < tr[2]; ++tx) {
                int64_t tile = (int64_t)ty * tw + tx;
                keys[m] = ((int64_t)view << (32 + tile_bits)) | (tile << 32) | (int64_t)low[i];
                vals[m] = view * view_stride_n + i;
                ++m;
            }
    }
}

/* stable LSD radix sort on the unsigned 64-bit key (reference semantics of a stable sort) */
void ora_sort_pairs(int64_t *keys, int32_t *vals, size_t M)
{
    if (M < 2) return;
    uint64_t *k0 = (uint64_t *)keys, *k1 = (uint64_t *)malloc(M * 8);
    int32_t *v0 = vals, *v1 = (int32_t *)malloc(M * 4);
    for (int pass = 0; pass < 8; ++pass) {
        size_t cnt[257];
        memset(cnt, 0, sizeof cnt);
        int sh = pass * 8;
        for (size_t i = 0; i < M; ++i) cnt[((k0[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < M; ++i) {
            size_t d = (k0[i] >> sh) & 255;
            k1[cnt[d]] = k0[i]; v1[cnt[d]] = v0[i]; cnt[d]++;
        }
        uint64_t *tk = k0; k0 = k1; k1 = tk;
        int32_t *tv = v0; v0 = v1; v1 = tv;
    }
    /* 8 passes: data is back in the caller's buffers */
    free(k1); free(v1);
}

/* offsets[t] = first sorted index whose (view,tile) >= t; offsets has n_views*n_tiles + 1 entries */
void ora_tile_ranges(const int64_t *keys, size_t M, int n_views, int n_tiles, int tile_bits, int32_t *offsets)
{
    size_t total = (size_t)n_views * n_tiles;
    size_t j = 0;
    for (size_t t = 0; t < total; ++t) {
        int64_t view = (int64_t)(t / n_tiles), tile = (int64_t)(t % n_tiles);
        int64_t want = (view << tile_bits) | tile;
        while (j < M && (keys[j] >> 32) < want) ++j;
        offsets[t] = (int32_t)j;
    }
    offsets[total] = (int32_t)M;
}

/* ------------------------------------------------------------------------------------ */
/* Rasterizers (per view).  vals index the splat table directly (view-local ids).        */
/* ------------------------------------------------------------------------------------ */
static inline float sigma3d(const float *g, float px, float py, float *dx_, float *dy_)
{
    float dx = g[0] - px, dy = g[1] - py;
    float hA = 0.5f * g[2], hC = 0.5f * g[4];
    float uu = fmaf(g[3], dy, hA * dx);
    float s = uu * dx;
    float wv = hC * dy;
    s = fmaf(wv, dy, s);
    *dx_ = dx; *dy_ = dy;
    return s;
}

/*
 * gsplat rasterize_to_pixels_3dgs_fwd restated: pixel centre (+0.5),
 * alpha = min(0.999, o e^-sigma), skip sigma<0 or alpha<1/255, stop BEFORE adding when
 * T(1-alpha) <= 1e-4.  last[p] = 1 + sorted index of the last contributing entry, or the
 * tile's range start when nothing contributed.
 */
void ora3d_raster_fwd(int W, int H, const float *geom, const float *rgbtab, const int32_t *vals, int id_base,
                      const int32_t *offsets /* n_tiles+1 for this view */, const float *bg,
                      float *out_rgb, float *out_alpha, int32_t *n_contrib, int32_t *last)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    for (int ty = 0; ty < th; ++ty)
        for (int tx = 0; tx < tw; ++tx) {
            int s = offsets[ty * tw + tx], e = offsets[ty * tw + tx + 1];
            for (int i = ty * ORA_TILE; i < (ty + 1) * ORA_TILE && i < H; ++i)
                for (int j = tx * ORA_TILE; j < (tx + 1) * ORA_TILE && j < W; ++j) {
                    float px = (float)j + 0.5f, py = (float)i + 0.5f;
                    float T = 1.0f, r = 0.0f, gcol = 0.0f, b = 0.0f;
                    int cnt = 0, lst = s;
                    for (int k = s; k < e; ++k) {
                        const float *g = geom + 8 * (size_t)(vals[k] - id_base);
                        float dx, dy;
                        float sg = sigma3d(g, px, py, &dx, &dy);
                        float alpha = fminf(ORA_ALPHA_MAX, g[5] * d_exp(-sg));
                        if (sg < 0.0f || alpha < ORA_ALPHA_MIN) continue;
                        float nT = T * (1.0f - alpha);
                        if (nT <= ORA_T_STOP_3D) break;
                        float vis = alpha * T;
                        const float *c = rgbtab + 3 * (size_t)(vals[k] - id_base);
                        r = fmaf(vis, c[0], r); gcol = fmaf(vis, c[1], gcol); b = fmaf(vis, c[2], b);
                        T = nT; ++cnt; lst = k + 1;
                    }
                    size_t p = (size_t)i * W + j;
                    out_rgb[3 * p + 0] = fmaf(T, bg[0], r);
                    out_rgb[3 * p + 1] = fmaf(T, bg[1], gcol);
                    out_rgb[3 * p + 2] = fmaf(T, bg[2], b);
                    out_alpha[p] = 1.0f - T;
                    if (n_contrib) n_contrib[p] = cnt;
                    if (last) last[p] = lst;
                }
        }
}

/*
 * Backward of the above for L = sum w_rgb*rgb + sum w_a*alpha.  Division-free form
 * (SURVEY 3.4 / 8c-c5 formulas re-derived): with S_i = B_i.w_rgb - R_i w_a,
 *   dL/dalpha_i = T_{i-1} (c_i.w_rgb - S_i),  S_{i-1} = S_i + alpha_i (c_i.w_rgb - S_i),
 *   S_end = bg.w_rgb - w_a.
 * Per-(view,Gaussian) accumulators acc[9] = v_rgb(3), v_A, v_B, v_C, v_x, v_y, v_opacity.
 * Accumulates in double so the oracle is a tight reference for the fp32 atomics.
 */
void ora3d_raster_bwd(int W, int H, const float *geom, const float *rgbtab, const int32_t *vals, int id_base,
                      const int32_t *offsets, const float *bg, const int32_t *last,
                      const float *w_rgb, const float *w_a, double *acc /* [N,9] */)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    int cap = 1024;
    float *sT = (float *)malloc(cap * 4), *sA = (float *)malloc(cap * 4), *sE = (float *)malloc(cap * 4);
    int *sK = (int *)malloc(cap * 4);
    for (int ty = 0; ty < th; ++ty)
        for (int tx = 0; tx < tw; ++tx) {
            int s = offsets[ty * tw + tx];
            for (int i = ty * ORA_TILE; i < (ty + 1) * ORA_TILE && i < H; ++i)
                for (int j = tx * ORA_TILE; j < (tx + 1) * ORA_TILE && j < W; ++j) {
                    size_t p = (size_t)i * W + j;
                    int e = last[p];
                    float px = (float)j + 0.5f, py = (float)i + 0.5f;
                    float T = 1.0f;
                    int n = 0;
                    for (int k = s; k < e; ++k) {
                        const float *g = geom + 8 * (size_t)(vals[k] - id_base);
                        float dx, dy;
                        float sg = sigma3d(g, px, py, &dx, &dy);
                        float ex = d_exp(-sg);
                        float alpha = fminf(ORA_ALPHA_MAX, g[5] * ex);
                        if (sg < 0.0f || alpha < ORA_ALPHA_MIN) continue;
                        if (n == cap) {
                            cap *= 2;
                            sT = (float *)realloc(sT, cap * 4); sA = (float *)realloc(sA, cap * 4);
                            sE = (float *)realloc(sE, cap * 4); sK = (int *)realloc(sK, cap * 4);
                        }
                        sT[n] = T; sA[n] = alpha; sE[n] = ex; sK[n] = k; ++n;
                        T = T * (1.0f - alpha);
                    }
                    const float *wr = w_rgb + 3 * p;
                    double S = (double)bg[0] * wr[0] + (double)bg[1] * wr[1] + (double)bg[2] * wr[2] - (double)w_a[p];
                    for (int q = n - 1; q >= 0; --q) {
                        int gid = vals[sK[q]] - id_base;
                        const float *g = geom + 8 * (size_t)gid;
                        const float *c = rgbtab + 3 * (size_t)gid;
                        double *a = acc + 9 * (size_t)gid;
                        double Tm = sT[q], al = sA[q];
                        double cw = (double)c[0] * wr[0] + (double)c[1] * wr[1] + (double)c[2] * wr[2];
                        double v_alpha = Tm * (cw - S);
                        double vis = al * Tm;
                        a[0] += vis * wr[0]; a[1] += vis * wr[1]; a[2] += vis * wr[2];
                        if (g[5] * sE[q] <= ORA_ALPHA_MAX) {
                            double dx = (double)g[0] - px, dy = (double)g[1] - py;
                            double v_sigma = -(double)g[5] * sE[q] * v_alpha;
                            a[3] += 0.5 * v_sigma * dx * dx;
                            a[4] += v_sigma * dx * dy;
                            a[5] += 0.5 * v_sigma * dy * dy;
                            a[6] += v_sigma * ((double)g[2] * dx + (double)g[3] * dy);
                            a[7] += v_sigma * ((double)g[3] * dx + (double)g[4] * dy);
                            a[8] += (double)sE[q] * v_alpha;
                        }
                        S = S + al * (cw - S);
                    }
                }
        }
    free(sT); free(sA); free(sE); free(sK);
}

/*
 * 2D: src/gaussian_renderer.py:395-425 restricted to listed Gaussians whose pixel rect
 * contains the pixel, in row order.  The reference accumulates A += g(1-A) (:417-425); the
 * contract carries T = 1-A multiplicatively, T <- T(1-g) (identical in exact arithmetic,
 * <= 3e-6 apart in fp32 over 16000 layers) so that the backward can replay T by
 * division.  Stop once T <= 2^-20 (what is left can add at most 9.6e-7).
 */
/* returns 0 if the pixel is outside the Gaussian's footprint q <= L, else 1 and g = o * exp(-q) */
static inline int g2d(const float *g, float x, float y, float *dxr_, float *dyr_, float *gv)
{
    float dx = x - g[0], dy = y - g[1];
    float dxr = fmaf(g[3], dy, g[2] * dx);
    float dyr = fmaf(g[2], dy, (-g[3]) * dx);
    float q = fmaf(dyr * dyr, g[5], (dxr * dxr) * g[4]);
    *dxr_ = dxr; *dyr_ = dyr;
    if (!(q <= g[7])) return 0;
    *gv = g[6] * d_exp(-q);
    return 1;
}

void ora2d_raster_fwd(int W, int H, const float *geom, const float *rgbtab, const int *rect,
                      const int32_t *vals, int id_base, const int32_t *offsets, const float *bg,
                      float *out_rgb, float *out_alpha, int32_t *n_contrib, int32_t *last)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    for (int ty = 0; ty < th; ++ty)
        for (int tx = 0; tx < tw; ++tx) {
            int s = offsets[ty * tw + tx], e = offsets[ty * tw + tx + 1];
            for (int i = ty * ORA_TILE; i < (ty + 1) * ORA_TILE && i < H; ++i)
                for (int j = tx * ORA_TILE; j < (tx + 1) * ORA_TILE && j < W; ++j) {
                    float T = 1.0f, r = 0.0f, gc = 0.0f, b = 0.0f;
                    int cnt = 0, lst = s;
                    for (int k = s; k < e; ++k) {
                        int gid = vals[k] - id_base;
                        float dxr, dyr, gv;
                        if (!g2d(geom + 8 * (size_t)gid, (float)j, (float)i, &dxr, &dyr, &gv)) continue;
                        float contrib = gv * T;
                        const float *c = rgbtab + 3 * (size_t)gid;
                        r = fmaf(contrib, c[0], r); gc = fmaf(contrib, c[1], gc); b = fmaf(contrib, c[2], b);
                        T = T * (1.0f - gv);
                        ++cnt; lst = k + 1;
                        if (T <= ORA_T_STOP_2D) break;
                    }
                    size_t p = (size_t)i * W + j;
                    out_rgb[3 * p + 0] = fmaf(T, bg[0], r);
                    out_rgb[3 * p + 1] = fmaf(T, bg[1], gc);
                    out_rgb[3 * p + 2] = fmaf(T, bg[2], b);
                    out_alpha[p] = 1.0f - T;
                    if (n_contrib) n_contrib[p] = cnt;
                    if (last) last[p] = lst;
                }
        }
}

/* acc[9] per Gaussian: v_rgb(3), sum d_dxr, sum d_dyr, d_theta, d_iax, d_iay, sum G_q */
void ora2d_raster_bwd(int W, int H, const float *geom, const float *rgbtab, const int *rect,
                      const int32_t *vals, int id_base, const int32_t *offsets, const float *bg,
                      const int32_t *last, const float *w_rgb, const float *w_a, double *acc)
{
    int tw = (W + ORA_TILE - 1) / ORA_TILE, th = (H + ORA_TILE - 1) / ORA_TILE;
    int cap = 1024;
    float *sT = (float *)malloc(cap * 4), *sG = (float *)malloc(cap * 4);
    float *sX = (float *)malloc(cap * 4), *sY = (float *)malloc(cap * 4);
    int *sK = (int *)malloc(cap * 4);
    for (int ty = 0; ty < th; ++ty)
        for (int tx = 0; tx < tw; ++tx) {
            int s = offsets[ty * tw + tx];
            for (int i = ty * ORA_TILE; i < (ty + 1) * ORA_TILE && i < H; ++i)
                for (int j = tx * ORA_TILE; j < (tx + 1) * ORA_TILE && j < W; ++j) {
                    size_t p = (size_t)i * W + j;
                    int e = last[p];
                    float T = 1.0f;
                    int n = 0;
                    for (int k = s; k < e; ++k) {
                        int gid = vals[k] - id_base;
                        float dxr, dyr, gv;
                        if (!g2d(geom + 8 * (size_t)gid, (float)j, (float)i, &dxr, &dyr, &gv)) continue;
                        if (n == cap) {
                            cap *= 2;
                            sT = (float *)realloc(sT, cap * 4); sG = (float *)realloc(sG, cap * 4);
                            sX = (float *)realloc(sX, cap * 4); sY = (float *)realloc(sY, cap * 4);
                            sK = (int *)realloc(sK, cap * 4);
                        }
                        sT[n] = T; sG[n] = gv; sX[n] = dxr; sY[n] = dyr; sK[n] = k; ++n;
                        T = T * (1.0f - gv);
                    }
                    const float *wr = w_rgb + 3 * p;
                    double S = (double)bg[0] * wr[0] + (double)bg[1] * wr[1] + (double)bg[2] * wr[2] - (double)w_a[p];
                    for (int q = n - 1; q >= 0; --q) {
                        int gid = vals[sK[q]] - id_base;
                        const float *g = geom + 8 * (size_t)gid;
                        const float *c = rgbtab + 3 * (size_t)gid;
                        double *a = acc + 9 * (size_t)gid;
                        double Tm = sT[q], gv = sG[q];
                        double cw = (double)c[0] * wr[0] + (double)c[1] * wr[1] + (double)c[2] * wr[2];
                        double dLdg = Tm * (cw - S);
                        double contrib = gv * Tm;
                        a[0] += contrib * wr[0]; a[1] += contrib * wr[1]; a[2] += contrib * wr[2];
                        double Gq = -gv * dLdg;
                        double ddxr = 2.0 * sX[q] * (double)g[4] * Gq, ddyr = 2.0 * sY[q] * (double)g[5] * Gq;
                        a[3] += ddxr; a[4] += ddyr;
                        a[5] += ddxr * sY[q] - ddyr * sX[q];
                        a[6] += (double)sX[q] * sX[q] * Gq;
                        a[7] += (double)sY[q] * sY[q] * Gq;
                        a[8] += Gq;
                        S = S + gv * (cw - S);
                    }
                }
        }
    free(sT); free(sG); free(sX); free(sY); free(sK);
}

/* chain rule back to the raw [N,9] rows (activations of :321-323; clamp passes 0<=c<=1) */
void ora2d_project_bwd(const float *params, int N, const float *geom, const double *acc, double *d_params)
{
    for (int i = 0; i < N; ++i) {
        const float *row = params + 9 * (size_t)i;
        const float *g = geom + 8 * (size_t)i;
        const double *a = acc + 9 * (size_t)i;
        double *d = d_params + 9 * (size_t)i;
        double cs = g[2], sn = g[3], iax = g[4], iay = g[5], o = g[6];
        double sx = d_exp(row[2]), sy = d_exp(row[3]);
        d[0] += -(cs * a[3] - sn * a[4]);
        d[1] += -(sn * a[3] + cs * a[4]);
        d[2] += a[6] * (-(iax * iax) * 4.0 * sx * sx);
        d[3] += a[7] * (-(iay * iay) * 4.0 * sy * sy);
        d[4] += a[5];
        for (int k = 0; k < 3; ++k)
            if (row[5 + k] >= 0.0f && row[5 + k] <= 1.0f) d[5 + k] += a[k];
        d[8] += -a[8] * (1.0 - o);
    }
}

/* gsplat fully_fused_projection backward + adapter activations, in double from fp32 saves */
void ora3d_project_bwd(const float *params, int N, const float *V, const float *K, int W, int H,
                       float near_plane, float far_plane, float radius_clip, float eps2d,
                       const double *acc, double *d_params, int activated)
{
    float geom[8], rgb[3]; int rect[4], trect[4]; uint32_t low;
    ora_proj3d_tmp t;
    for (int i = 0; i < N; ++i) {
        const float *row = params + 14 * (size_t)i;
        const double *a = acc + 9 * (size_t)i;
        double *d = d_params + 14 * (size_t)i;
        /* colour and opacity do not depend on visibility of the projection */
        int ok = project3d_one(row, V, K, W, H, near_plane, far_plane, radius_clip, eps2d,
                               geom, rgb, rect, trect, &low, &t, activated);
        double v_act[14] = { 0 }; /* gradient w.r.t. the activated values (what gsplat's backward returns) */
        for (int k = 0; k < 3; ++k) v_act[10 + k] = a[k];
        v_act[13] = a[8];
        if (!ok) {
            if (activated) { for (int k = 0; k < 14; ++k) d[k] += v_act[k]; }
            else adapter3d_vjp_one(row, t.s, t.qn_raw, geom[5], v_act, d);
            continue;
        }
        double fx = K[0], fy = K[4];
        /* conic = inverse(cov2d): G_S = -X G_X X with G_X = [[vA, vB/2],[vB/2, vC]] */
        double XA = geom[2], XB = geom[3], XC = geom[4];
        double gA = a[3], gB = 0.5 * a[4], gC = a[5];
        /* P = X * G_X */
        double P00 = XA * gA + XB * gB, P01 = XA * gB + XB * gC;
        double P10 = XB * gA + XC * gB, P11 = XB * gB + XC * gC;
        double G00 = -(P00 * XA + P01 * XB), G01 = -(P00 * XB + P01 * XC);
        double G10 = -(P10 * XA + P11 * XB), G11 = -(P10 * XB + P11 * XC);
        double Gs01 = 0.5 * (G01 + G10);
        /* cov2d = J Sc J^T ; J = [[J00,0,J02],[0,J11,J12]] */
        double J[2][3] = { { t.J00, 0.0, t.J02 }, { 0.0, t.J11, t.J12 } };
        double Gs[2][2] = { { G00, Gs01 }, { Gs01, G11 } };
        double Scf[3][3] = { { t.Sc[0], t.Sc[1], t.Sc[2] }, { t.Sc[1], t.Sc[3], t.Sc[4] }, { t.Sc[2], t.Sc[4], t.Sc[5] } };
        double GSc[3][3], GJ[2][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int p = 0; p < 2; ++p)
                    for (int q = 0; q < 2; ++q) s += J[p][r] * Gs[p][q] * J[q][c];
                GSc[r][c] = s;
            }
        for (int p = 0; p < 2; ++p)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int q = 0; q < 2; ++q)
                    for (int r = 0; r < 3; ++r) s += Gs[p][q] * J[q][r] * Scf[r][c];
                GJ[p][c] = 2.0 * s;
            }
        double x = t.pc[0], y = t.pc[1], z = t.pc[2];
        double rz = 1.0 / z, rz2 = rz * rz, rz3 = rz2 * rz;
        double tx = t.tx, ty = t.ty;
        double vpc[3] = { 0, 0, 0 };
        /* mean2d */
        vpc[0] += fx * rz * a[6];
        vpc[1] += fy * rz * a[7];
        vpc[2] += -(fx * x * a[6] + fy * y * a[7]) * rz2;
        /* J entries */
        vpc[2] += -fx * rz2 * GJ[0][0] - fy * rz2 * GJ[1][1];
        if (!t.clampx) { vpc[0] += -fx * rz2 * GJ[0][2]; vpc[2] += 2.0 * fx * tx * rz3 * GJ[0][2]; }
        else           { vpc[2] += fx * tx * rz3 * GJ[0][2]; }
        if (!t.clampy) { vpc[1] += -fy * rz2 * GJ[1][2]; vpc[2] += 2.0 * fy * ty * rz3 * GJ[1][2]; }
        else           { vpc[2] += fy * ty * rz3 * GJ[1][2]; }
        /* world <- camera */
        double Rw[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rw[r][c] = V[4 * r + c];
        for (int c = 0; c < 3; ++c) v_act[c] = Rw[0][c] * vpc[0] + Rw[1][c] * vpc[1] + Rw[2][c] * vpc[2];
        double GS[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int p = 0; p < 3; ++p)
                    for (int q = 0; q < 3; ++q) s += Rw[p][r] * GSc[p][q] * Rw[q][c];
                GS[r][c] = s;
            }
        /* Sigma = M M^T -> G_M = (G + G^T) M */
        double Mm[3][3], Rq[3][3], GM[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) { Mm[r][c] = t.M[3 * r + c]; Rq[r][c] = t.R[3 * r + c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int p = 0; p < 3; ++p) s += (GS[r][p] + GS[p][r]) * Mm[p][c];
                GM[r][c] = s;
            }
        double GR[3][3], vs[3] = { 0, 0, 0 };
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) { GR[r][c] = GM[r][c] * t.s[c]; vs[c] += Rq[r][c] * GM[r][c]; }
        for (int k = 0; k < 3; ++k) v_act[3 + k] = vs[k];
        /* R(qh) */
        double w = t.qh[0], qx = t.qh[1], qy = t.qh[2], qz = t.qh[3];
        double vq[4];
        vq[0] = 2.0 * (-qz * GR[0][1] + qy * GR[0][2] + qz * GR[1][0] - qx * GR[1][2] - qy * GR[2][0] + qx * GR[2][1]);
        vq[1] = 2.0 * (qy * GR[0][1] + qz * GR[0][2] + qy * GR[1][0] - 2.0 * qx * GR[1][1] - w * GR[1][2] + qz * GR[2][0] + w * GR[2][1] - 2.0 * qx * GR[2][2]);
        vq[2] = 2.0 * (-2.0 * qy * GR[0][0] + qx * GR[0][1] + w * GR[0][2] + qx * GR[1][0] + qz * GR[1][2] - w * GR[2][0] + qz * GR[2][1] - 2.0 * qy * GR[2][2]);
        vq[3] = 2.0 * (-2.0 * qz * GR[0][0] - w * GR[0][1] + qx * GR[0][2] + w * GR[1][0] - 2.0 * qz * GR[1][1] + qy * GR[1][2] + qx * GR[2][0] + qy * GR[2][1]);
        /* qh = qa / |qa| */
        double dot = vq[0] * w + vq[1] * qx + vq[2] * qy + vq[3] * qz;
        double va[4];
        double qhv[4] = { w, qx, qy, qz };
        for (int k = 0; k < 4; ++k) va[k] = (vq[k] - dot * qhv[k]) * t.inv2;
        for (int k = 0; k < 4; ++k) v_act[6 + k] = va[k];
        if (activated) { for (int k = 0; k < 14; ++k) d[k] += v_act[k]; }
        else adapter3d_vjp_one(row, t.s, t.qn_raw, geom[5], v_act, d);
    }
}
