// Host build of the product's arithmetic contract (pose_splatter_b200/csrc/ps_contract.cuh) so that the
// CPU test-suite can compare it bit-for-bit with the independent oracle restatement without a GPU.
// Build: g++ -O2 -shared -fPIC -ffp-contract=off -o tests/_host_contract.so tests/host_contract.cpp
#include "../pose_splatter_b200/csrc/ps_contract.cuh"

extern "C" {
void hc_math_probe(const float *x, int n, float *o_exp, float *o_log, float *o_sig, float *o_sin, float *o_cos)
{
    for (int i = 0; i < n; ++i) {
        o_exp[i] = psm_exp(x[i]);
        o_log[i] = psm_log(fabsf(x[i]) + 1e-30f);
        o_sig[i] = psm_sigmoid(x[i]);
        psm_sincos(x[i], &o_sin[i], &o_cos[i]);
    }
}
// records out: r [N,12] floats, tile [N,4] ints, low [N]
void hc_project(int mode, const float *params, int N, const float *V, const float *K, int W, int H, float near_plane,
                float far_plane, float radius_clip, float eps2d, float *r, int *tile, uint32_t *low, int activated)
{
    PsRecord rec;
    PsProj3dAux aux;
    for (int i = 0; i < N; ++i) {
        if (mode == 3) ps_project3d(params + 14 * (size_t)i, V, K, W, H, near_plane, far_plane, radius_clip, eps2d, &rec, &aux, activated);
        else ps_project2d(params + 9 * (size_t)i, (uint32_t)i, W, H, &rec);
        for (int k = 0; k < 4; ++k) {
            r[12 * (size_t)i + k] = rec.r0[k]; r[12 * (size_t)i + 4 + k] = rec.r1[k]; r[12 * (size_t)i + 8 + k] = rec.r2[k];
            tile[4 * (size_t)i + k] = rec.tile[k];
        }
        low[i] = rec.low;
    }
}
void hc_pairs3d(const float *g /*x y A B C*/, const float *px, const float *py, int n, float *sigma)
{
    float dx, dy;
    for (int i = 0; i < n; ++i) sigma[i] = ps_sigma3d(g[0], g[1], psm_mul(0.5f, g[2]), g[3], psm_mul(0.5f, g[4]), px[i], py[i], &dx, &dy);
}
void hc_pairs2d(const float *g /*u v cs sn iax iay*/, const float *x, const float *y, int n, float *q)
{
    float a, b;
    for (int i = 0; i < n; ++i) q[i] = ps_q2d(g[0], g[1], g[2], g[3], g[4], g[5], x[i], y[i], &a, &b);
}
}
