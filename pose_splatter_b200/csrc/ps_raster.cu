// ps_raster.cu -- per-tile rasterizers (SURVEY.md 2.2 K5', K6'), both gaussian modes.  (v4)
//
// block_lists: every non-empty (view, 16x16 tile) list is split, in list order, into the lists of its eight
//   8x4 pixel blocks: an entry goes to a block only if the splat can contribute there (3D: exact
//   ellipse-vs-box test, min of sigma over the block's pixel centres against log(255*opacity); 2D: the
//   splat's pixel rectangle meets the block).  One CTA per tile, one list entry per thread, ordered
//   multi-split with ballots.  The block lists hold view * N + Gaussian (4 B each).  [fallback for N too large for the
//   bitmap split of ps_bin.cu, which builds the same lists without reading a record]
// forward / backward: the unit of work is one WARP = one pixel block; a CTA is two such warps that share
//   nothing (no __syncthreads), so a block that finishes early frees its slot: this workload is dominated by
//   a few very long lists in which, at any depth, only some blocks still have live pixels.
//   Tasks are launched in the size order of the work list.  A warp streams its block list 32 entries at a
//   time (one per lane): Gaussian id -> cp.async of the 48-byte splat record into a warp-private
//   3-stage ring in shared memory, two chunks ahead of use; the lane re-culls its entry against the bounding
//   box of the pixels that are still live (forward: not terminated; backward: still have contributors at or
//   before this chunk); survivors (ballot) are walked in list order with broadcast reads of the ring.
//   Culling never changes a result: it only removes pairs whose alpha test (3D) / rectangle test (2D) is
//   guaranteed to fail for every live pixel.
//     forward : front-to-back compositing, two survivors in flight, per-warp early exit   [FP32 pipe + smem]
//     backward: reverse replay from the block's last contributor; transmittance recovered by division from
//               the saved "T before the last contributor"; the nine per-splat gradient sums are reduced over
//               the warp (multi-value butterfly) and added to the (view, Gaussian) accumulator row with
//               red.global.add                                                          [FP32 pipe + shfl]
// Replaces gsplat rasterize_to_pixels_3dgs_fwd/bwd (absent from the reference tree) and the
// torch element-wise loop src/gaussian_renderer.py:379-425 plus its autograd.
#include "ps_contract.cuh"
#include "ps_cull.cuh"
#include "ps_internal.h"
#include <cstdlib>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr float THR_SLACK = PS_THR_SLACK;
constexpr int CH = 32;              // list entries per chunk (one per lane)
constexpr int NS = 3;               // ring stages per warp
constexpr int WPC = 2;              // warps (pixel blocks) per CTA
constexpr int RT_THREADS = WPC * 32;
constexpr int TASKS_PER_TILE = 8 / WPC;

struct BlockCtx {
    int view, start, len;   // tile list [start, start + len)
    int nb;                 // entries of this block's list
    const uint32_t *bl;     // this block's list: view * N + Gaussian of its entries, in tile-list order
    const uint32_t *bp;     // tile-list positions (relative to `start`) of the entries, or NULL (last-id tap only)
    int px, py;             // this lane's pixel
    int bx, by;             // block origin
    bool inside;
};

// task = 8 * work-list item + block (0..7: x half = blk & 1, y quarter = blk >> 1)
__device__ __forceinline__ BlockCtx block_ctx_task(const PsGeometry &g, const int32_t *offsets, const int32_t *worklist,
                                                   const uint32_t *blist, const uint32_t *bpos, const int32_t *bcount, unsigned task)
{
    BlockCtx c;
    const int item = (int)(task >> 3);
    const int blk = (int)(task & 7u);
    const int lin = worklist[item]; // non-empty (view, tile) lists, longest size class first
    c.view = lin / g.n_tiles;
    const int tile = lin - c.view * g.n_tiles;
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    c.start = offsets[lin];
    c.len = offsets[lin + 1] - c.start;
    c.nb = bcount[item * 8 + blk];
    c.bl = blist + 8 * (size_t)c.start + (size_t)blk * c.len;
    c.bp = bpos ? bpos + 8 * (size_t)c.start + (size_t)blk * c.len : nullptr;
    const int lane = threadIdx.x & 31;
    c.bx = tx * PS_TILE + (blk & 1) * 8;
    c.by = ty * PS_TILE + (blk >> 1) * 4;
    c.px = c.bx + (lane & 7);
    c.py = c.by + (lane >> 3);
    c.inside = c.px < g.W && c.py < g.H;
    return c;
}
template <int W> // W = warps (pixel blocks) per CTA; CTA b handles blocks (b % (8 / W)) * W ... of item b / (8 / W)
__device__ __forceinline__ BlockCtx block_ctx(const PsGeometry &g, const int32_t *offsets, const int32_t *worklist,
                                              const uint32_t *blist, const uint32_t *bpos, const int32_t *bcount)
{
    return block_ctx_task(g, offsets, worklist, blist, bpos, bcount, blockIdx.x * W + (threadIdx.x >> 5));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1-D bulk copies (the TMA engine's non-tensor form, cp.async.bulk -> UBLKCP) completing on an mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem, const void *gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// bounding box (in block-local pixel coordinates) of the lanes set in `active` (lane = y * 8 + x); active != 0
__device__ __forceinline__ void active_box(uint32_t active, int &x0, int &x1, int &y0, int &y1)
{
    const uint32_t cols = (active | (active >> 8) | (active >> 16) | (active >> 24)) & 0xffu;
    x0 = __ffs(cols) - 1;
    x1 = 31 - __clz(cols);
    y0 = (__ffs(active) - 1) >> 3;
    y1 = (31 - __clz(active)) >> 3;
}

// uint8 RGBA exactly as the reference's evaluation writer quantises it (scripts/utils/evaluate_model.py:110-113):
// (255 * clip(x, 0, 1)).astype(uint8), i.e. fp32 multiply, truncation
__device__ __forceinline__ uint32_t quantise_rgba8(float r, float g, float b, float a)
{
    const uint32_t qr = (uint32_t)psm_mul(255.0f, fminf(fmaxf(r, 0.0f), 1.0f));
    const uint32_t qg = (uint32_t)psm_mul(255.0f, fminf(fmaxf(g, 0.0f), 1.0f));
    const uint32_t qb = (uint32_t)psm_mul(255.0f, fminf(fmaxf(b, 0.0f), 1.0f));
    const uint32_t qa = (uint32_t)psm_mul(255.0f, fminf(fmaxf(a, 0.0f), 1.0f));
    return qr | (qg << 8) | (qb << 16) | (qa << 24);
}

// The warp-private ring: stage st holds chunk data for 32 entries.
struct Ring {
    float4 (*a)[CH], (*b)[CH], (*c)[CH];
};

// lane-parallel cull of the staged chunk against the live-pixel box -> ballot of surviving entries
template <int MODE>
__device__ __forceinline__ uint32_t cull_chunk(const Ring &q, int st, int lane, bool valid, const BlockCtx &c, uint32_t live)
{
    int ax0, ax1, ay0, ay1;
    active_box(live, ax0, ax1, ay0, ay1);
    bool hit = false;
    if (valid) {
        const float4 a0 = q.a[st][lane], a1 = q.b[st][lane];
        const float half = (MODE == PS_MODE_3D) ? 0.5f : 0.0f; // pixel centres: +0.5 in 3D, integers in 2D
        float hA = a1.x, B = a1.y, hC = a1.z;
        if (MODE == PS_MODE_2D) ps_conic2d(a1, hA, B, hC);
        hit = ps_ellipse_hits_box(a0.x, a0.y, hA, B, hC, a0.z, (float)(c.bx + ax0) + half, (float)(c.bx + ax1) + half,
                               (float)(c.by + ay0) + half, (float)(c.by + ay1) + half);
    }
    return __ballot_sync(FULL, hit);
}

template <int MODE, bool STATS, int W>
__global__ void __launch_bounds__(W * 32)
raster_fwd_kernel(PsGeometry g, PsTable t, const uint32_t *__restrict__ vals, const int32_t *__restrict__ offsets,
                  const int32_t *__restrict__ worklist, const float *__restrict__ background,
                  float *__restrict__ rgb, float *__restrict__ alpha,
                  int32_t *__restrict__ n_contrib, int32_t *__restrict__ last, int32_t *__restrict__ blast,
                  float *__restrict__ t_pen, const uint32_t *__restrict__ blist, const uint32_t *__restrict__ bpos,
                  const int32_t *__restrict__ bcount, uint32_t *__restrict__ rgba8, unsigned long long *__restrict__ stats,
                  const int32_t *__restrict__ n_lists, uint32_t *__restrict__ cmask, uint32_t *__restrict__ cids,
                  int32_t *__restrict__ ccount)
{
    __shared__ float4 s_a[W][NS][CH], s_b[W][NS][CH], s_c[W][NS][CH];
    // the grid may be sized from an upper bound of the number of non-empty lists (sync-free small calls)
    if ((int)((blockIdx.x * W + (threadIdx.x >> 5)) >> 3) >= __ldg(n_lists)) return;
    const BlockCtx c = block_ctx<W>(g, offsets, worklist, blist, bpos, bcount);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool done = !c.inside;
    if (__all_sync(FULL, done)) { // block entirely outside the image
        if (ccount && lane == 0) ccount[blockIdx.x * W + wid] = 0;
        return;
    }
    const Ring q = { s_a[wid], s_b[wid], s_c[wid] };
    const int len = c.nb;
    const int nchunks = (len + CH - 1) / CH;
    // chunk cj: Gaussian id (ld) -> record copies (cp.async); software-pipelined: ids run 4 chunks ahead, records 2
    auto issue = [&](int cj, uint32_t id) { // one commit group per chunk, also when it does not exist
        if (cj * CH + lane < len) {
            const int st = cj % NS;
            const float4 *src = PS_REC(t, id, 0);
            cp_async16(&q.a[st][lane], src);
            cp_async16(&q.b[st][lane], src + 1);
            cp_async16(&q.c[st][lane], src + 2);
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int cj) -> uint32_t { return (cj * CH + lane < len) ? __ldg(c.bl + cj * CH + lane) : 0u; };
    uint32_t idn, idnn, id_c, id_c1; // ids of chunks ci + 2, ci + 3 (to be staged) and ci, ci + 1 (for the contributor list)
    {
        const uint32_t i0 = fetch_id(0), i1 = fetch_id(1);
        idn = fetch_id(2);
        idnn = fetch_id(3);
        issue(0, i0);
        issue(1, i1);
        id_c = i0; id_c1 = i1;
    }
    int n_cl = 0; // entries of this block's contributor list so far
    const size_t cl_off = (size_t)(c.bl - blist);

    unsigned long long st_eval = 0, st_walk = 0, st_staged = 0;
    const float pxf = (MODE == PS_MODE_3D) ? (float)c.px + 0.5f : (float)c.px;
    const float pyf = (MODE == PS_MODE_3D) ? (float)c.py + 0.5f : (float)c.py;
    float T = 1.0f, Tpen = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
    int cnt = 0;
    int blastpos = 0; // 1 + index (in the block list) of the last contributor
    for (int ci = 0; ci < nchunks; ++ci) {
        issue(ci + 2, idn);
        const uint32_t id_cur = id_c; // this lane's entry of chunk ci
        id_c = id_c1; id_c1 = idn;
        idn = idnn;
        idnn = fetch_id(ci + 4);
        cp_async_wait_group<2>(); // chunk ci has landed (this lane's copies); the barrier makes all lanes' visible
        __syncwarp();
        const int st = ci % NS;
        const int first = ci * CH;
        uint32_t mask = cull_chunk<MODE>(q, st, lane, ci * CH + lane < len, c, __ballot_sync(FULL, !done));
        if (STATS) { st_staged += 1; if (lane == 0) st_walk += __popc(mask); st_eval += done ? 0 : __popc(mask); }
        const float4 *q0 = q.a[st], *q1 = q.b[st], *q2 = q.c[st];
        uint32_t cm_mine = 0; // lane e: the pixels entry e of this chunk contributes to (saved for the backward)
        // survivors two at a time: the two alpha evaluations are independent chains, the compositing is ordered
        while (mask) {
            const int ea = __ffs(mask) - 1;
            mask &= mask - 1;
            const bool two = mask != 0;
            const int eb = two ? __ffs(mask) - 1 : ea;
            mask &= mask - 1;
            const float4 r0a = q0[ea], r1a = q1[ea], r0b = q0[eb], r1b = q1[eb];
            bool conta = false, contb = false;
            if (MODE == PS_MODE_3D) {
                float dx, dy;
                const float thra = r0a.z, thrb = r0b.z;
                const float sga = ps_sigma3d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, pxf, pyf, &dx, &dy);
                const float sgb = ps_sigma3d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, pxf, pyf, &dx, &dy);
                const bool canda = !done && sga >= 0.0f && sga <= thra + THR_SLACK;
                const bool candb = two && !done && sgb >= 0.0f && sgb <= thrb + THR_SLACK;
                if (!__any_sync(FULL, canda || candb)) continue;
                // candidates have 0 <= sigma <= ~5.6: the clamp inside psm_exp2 is the identity for them
                const float aa = fminf(PS_ALPHA_MAX, psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-sga, 0x1.715476p+0f))));
                const float ab = fminf(PS_ALPHA_MAX, psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-sgb, 0x1.715476p+0f))));
                if (canda && aa >= PS_ALPHA_MIN) {
                    const float nT = psm_mul(T, psm_sub(1.0f, aa));
                    if (nT <= PS_T_STOP_3D) {
                        done = true;
                    } else {
                        const float4 r2 = q2[ea];
                        const float vis = psm_mul(aa, T);
                        cr = psm_fma(vis, r2.x, cr); cg = psm_fma(vis, r2.y, cg); cb = psm_fma(vis, r2.z, cb);
                        Tpen = T; T = nT; ++cnt; blastpos = first + ea + 1; conta = true;
                    }
                }
                if (candb && !done && ab >= PS_ALPHA_MIN) {
                    const float nT = psm_mul(T, psm_sub(1.0f, ab));
                    if (nT <= PS_T_STOP_3D) {
                        done = true;
                    } else {
                        const float4 r2 = q2[eb];
                        const float vis = psm_mul(ab, T);
                        cr = psm_fma(vis, r2.x, cr); cg = psm_fma(vis, r2.y, cg); cb = psm_fma(vis, r2.z, cb);
                        Tpen = T; T = nT; ++cnt; blastpos = first + eb + 1; contb = true;
                    }
                }
            } else {
                float dxr, dyr;
                const float qa = ps_q2d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, r1a.w, pxf, pyf, &dxr, &dyr);
                const float qb = ps_q2d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, r1b.w, pxf, pyf, &dxr, &dyr);
                const bool ina = !done && qa <= r0a.z;        // inside the footprint q <= L
                const bool inb = two && !done && qb <= r0b.z;
                if (!__any_sync(FULL, ina || inb)) continue;
                const float4 r2a = q2[ea], r2b = q2[eb];
                // 0 <= q <= L <= ~20: the clamp inside psm_exp is the identity
                const float gva = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-qa, 0x1.715476p+0f)));
                const float gvb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-qb, 0x1.715476p+0f)));
                if (ina) {
                    const float contrib = psm_mul(gva, T);
                    cr = psm_fma(contrib, r2a.x, cr); cg = psm_fma(contrib, r2a.y, cg); cb = psm_fma(contrib, r2a.z, cb);
                    Tpen = T; T = psm_mul(T, psm_sub(1.0f, gva)); ++cnt; blastpos = first + ea + 1; conta = true;
                    if (T <= PS_T_STOP_2D) done = true;
                }
                if (inb && !done) {
                    const float contrib = psm_mul(gvb, T);
                    cr = psm_fma(contrib, r2b.x, cr); cg = psm_fma(contrib, r2b.y, cg); cb = psm_fma(contrib, r2b.z, cb);
                    Tpen = T; T = psm_mul(T, psm_sub(1.0f, gvb)); ++cnt; blastpos = first + eb + 1; contb = true;
                    if (T <= PS_T_STOP_2D) done = true;
                }
            }
            if (cmask) { // the walk is warp-uniform here: both ballots are taken by all lanes
                const uint32_t cma = __ballot_sync(FULL, conta), cmb = __ballot_sync(FULL, contb);
                cm_mine = (lane == ea) ? cma : cm_mine;
                cm_mine = (two && lane == eb) ? cmb : cm_mine;
            }
        }
        if (cmask) { // append the entries that contributed to some pixel, in list order: (id, pixel mask)
            const uint32_t nzb = __ballot_sync(FULL, cm_mine != 0u);
            if (cm_mine != 0u) {
                const size_t o = cl_off + n_cl + __popc(nzb & ((1u << lane) - 1u));
                cids[o] = id_cur;
                cmask[o] = cm_mine;
            }
            n_cl += __popc(nzb);
        }
        if (__all_sync(FULL, done)) break;
        __syncwarp(); // every lane is finished with stage st before chunk ci + 3 is copied into it
    }
    cp_async_wait_group<0>();
    if (ccount && lane == 0) ccount[blockIdx.x * W + wid] = n_cl;
    if (c.inside) {
        const size_t p = ((size_t)c.view * g.H + c.py) * g.W + c.px;
        const float b0 = __ldg(background), b1 = __ldg(background + 1), b2 = __ldg(background + 2);
        const float o0 = psm_fma(T, b0, cr), o1 = psm_fma(T, b1, cg), o2 = psm_fma(T, b2, cb), oa = psm_sub(1.0f, T);
        if (rgb) { rgb[3 * p + 0] = o0; rgb[3 * p + 1] = o1; rgb[3 * p + 2] = o2; }
        if (alpha) alpha[p] = oa;
        if (rgba8) rgba8[p] = quantise_rgba8(o0, o1, o2, oa);
        if (n_contrib) n_contrib[p] = cnt;
        if (last) last[p] = c.start + ((blastpos && c.bp) ? (int)__ldg(c.bp + blastpos - 1) + 1 : 0); // tile-list position
        if (blast) blast[p] = blastpos;
        if (t_pen) t_pen[p] = Tpen;
    }
    if (STATS) {
        // lanes that finish inside a chunk are still counted for the whole chunk: an upper bound
        // within 32 pairs per (lane, termination), negligible against the totals
        unsigned long long contributing = (unsigned long long)cnt;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            st_eval += __shfl_xor_sync(FULL, st_eval, d);
            contributing += __shfl_xor_sync(FULL, contributing, d);
        }
        if (lane == 0) {
            atomicAdd(stats + 0, st_eval);
            atomicAdd(stats + 1, contributing);
            atomicAdd(stats + 2, st_walk);
            atomicAdd(stats + 3, st_staged * CH);
        }
    }
}

// ---- forward v6: candidate masks + per-pixel walk --------------------------------------------------------------------
// v4 walks every surviving entry with all 32 lanes although ~5 of them are candidates.  Here lane e first computes, from
// the splat's conic, the exact set of block pixels its ellipse sigma <= thr can reach (per pixel row the solution of a
// quadratic, widened by 1e-3 px and 0.02 in sigma: a superset of the pixels that pass the fp32 test of v4, so no
// result changes); the 32 x 32 bit matrix (entry, pixel) is transposed across the warp with five shuffles, and every
// pixel lane then walks only ITS OWN candidate entries, in list order, with the same arithmetic as v4.

// pixels (bit = y * 8 + x of the 8x4 block whose first pixel centre is (bxf, byf)) where the quadratic form
// hA ux^2 + B ux uy + hC uy^2 <= lim can hold, u = pixel centre - (gx, gy)
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t span_mask(float gx, float gy, float hA, float B, float hC, float lim, float bxf, float byf)
{
    if (!(hA > 0.0f)) return 0xffffffffu; // not an ellipse in x: let the exact test decide everywhere
    // per pixel row uy: hA ux^2 + (B uy) ux + (hC uy^2 - lim) <= 0  <=>  |ux - c| <= h with
    // c = -B uy / (2 hA), h = sqrt(disc) / (2 hA), disc = (B^2 - 4 hA hC) uy^2 + 4 hA lim.  The root comes from
    // sqrt.approx (2 ulp): the interval is widened by 2e-3 px + 1e-6 relative, a superset of the exact test's pixels
    const float inv2a = __fdividef(0.5f, hA);
    const float gxr = gx - bxf;
    const float k2 = B * B - 4.0f * hA * hC, k0 = 4.0f * hA * lim;
    const float cs = -B * inv2a;
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float uy = (byf + (float)j) - gy;
        const float disc = fmaf(k2, uy * uy, k0);
        if (disc >= 0.0f) { // NaN (degenerate records) fails: such entries never pass the exact test either
            const float h = sqrt_approx(disc) * inv2a * 1.000001f + 2e-3f;
            const float ctr = fmaf(cs, uy, gxr);
            const float flo = fmaxf(ctr - h, -1.0f), fhi = fminf(ctr + h, 8.0f);
            const int ilo = max(0, (int)ceilf(flo)), ihi = min(7, (int)floorf(fhi));
            if (ilo <= ihi) m |= (((2u << ihi) - 1u) & ~((1u << ilo) - 1u)) << (8 * j);
        }
    }
    return m;
}

// 32 x 32 bit-matrix transpose across the warp: lane l gives row l, gets column l.  Five butterfly stages; the two
// byte-granular ones are one PRMT each (selector by the lane's bit), the three sub-byte ones a rotate and a bit-select:
//   lower lane of a pair: x = (x & M) | ((y << J) & ~M),  upper lane: x = ((y >> J) & M) | (x & ~M)
// (a rotation by J resp. 32 - J puts the wanted bits in place; the bits that wrap around fall under the discarded half
// of the mask).  Emulated against the definition for all 1024 single-bit matrices and random ones.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane)
{
    uint32_t y = __shfl_xor_sync(FULL, x, 16);
    x = __byte_perm(x, y, (lane & 16) ? 0x3276u : 0x5410u);
    y = __shfl_xor_sync(FULL, x, 8);
    x = __byte_perm(x, y, (lane & 8) ? 0x3715u : 0x6240u);
#define PS_TSTAGE(J, M)                                                                          \
    {                                                                                            \
        y = __shfl_xor_sync(FULL, x, J);                                                         \
        const bool up = (lane & J) != 0;                                                         \
        const uint32_t t = __funnelshift_l(y, y, up ? 32 - J : J);                               \
        const uint32_t m = up ? ~M : M;                                                          \
        x = (x & m) | (t & ~m);                                                                  \
    }
    PS_TSTAGE(4, 0x0f0f0f0fu) PS_TSTAGE(2, 0x33333333u) PS_TSTAGE(1, 0x55555555u)
#undef PS_TSTAGE
    return x;
}

template <int MODE, bool STATS, int W>
__global__ void __launch_bounds__(W * 32, 32 / W)
raster_fwd6_kernel(PsGeometry g, PsTable t, const uint32_t *__restrict__ vals, const int32_t *__restrict__ offsets,
                   const int32_t *__restrict__ worklist, const float *__restrict__ background,
                   float *__restrict__ rgb, float *__restrict__ alpha,
                   int32_t *__restrict__ n_contrib, int32_t *__restrict__ last, int32_t *__restrict__ blast,
                   float *__restrict__ t_pen, const uint32_t *__restrict__ blist, const uint32_t *__restrict__ bpos,
                   const int32_t *__restrict__ bcount, uint32_t *__restrict__ rgba8, unsigned long long *__restrict__ stats,
                  const int32_t *__restrict__ n_lists, uint32_t *__restrict__ cmask, uint32_t *__restrict__ cids,
                  int32_t *__restrict__ ccount)
{
    __shared__ float4 s_a[W][NS][CH], s_b[W][NS][CH], s_c[W][NS][CH];
    // the grid may be sized from an upper bound of the number of non-empty lists (sync-free small calls)
    if ((int)((blockIdx.x * W + (threadIdx.x >> 5)) >> 3) >= __ldg(n_lists)) return;
    const BlockCtx c = block_ctx<W>(g, offsets, worklist, blist, bpos, bcount);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool done = !c.inside;
    if (__all_sync(FULL, done)) {
        if (ccount && lane == 0) ccount[blockIdx.x * W + wid] = 0;
        return;
    }
    const Ring q = { s_a[wid], s_b[wid], s_c[wid] };
    const int len = c.nb;
    const int nchunks = (len + CH - 1) / CH;
    // chunk cj: Gaussian id (ld) -> record copies (cp.async); software-pipelined: ids run 4 chunks ahead, records 2
    auto issue = [&](int cj, uint32_t id) { // one commit group per chunk, also when it does not exist
        if (cj * CH + lane < len) {
            const int st = cj % NS;
            const float4 *src = PS_REC(t, id, 0);
            cp_async16(&q.a[st][lane], src);
            cp_async16(&q.b[st][lane], src + 1);
            cp_async16(&q.c[st][lane], src + 2);
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int cj) -> uint32_t { return (cj * CH + lane < len) ? __ldg(c.bl + cj * CH + lane) : 0u; };
    uint32_t idn, idnn, id_c, id_c1; // ids of chunks ci + 2, ci + 3 (to be staged) and ci, ci + 1 (for the contributor list)
    {
        const uint32_t i0 = fetch_id(0), i1 = fetch_id(1);
        idn = fetch_id(2);
        idnn = fetch_id(3);
        issue(0, i0);
        issue(1, i1);
        id_c = i0; id_c1 = i1;
    }
    int n_cl = 0; // entries of this block's contributor list so far
    const size_t cl_off = (size_t)(c.bl - blist);
    unsigned long long st_eval = 0, st_walk = 0, st_staged = 0;
    const float half = (MODE == PS_MODE_3D) ? 0.5f : 0.0f;
    const float pxf = (float)c.px + half, pyf = (float)c.py + half;
    const float bxf = (float)c.bx + half, byf = (float)c.by + half;
    float T = 1.0f, Tpen = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
    int cnt = 0;
    int blastpos = 0;
    for (int ci = 0; ci < nchunks; ++ci) {
        issue(ci + 2, idn);
        const uint32_t id_cur = id_c; // this lane's entry of chunk ci
        id_c = id_c1; id_c1 = idn;
        idn = idnn;
        idnn = fetch_id(ci + 4);
        cp_async_wait_group<2>();
        __syncwarp();
        const int st = ci % NS;
        const int first = ci * CH;
        const float4 *q0 = q.a[st], *q1 = q.b[st], *q2 = q.c[st];
        // lane = entry: which live pixels of the block can it reach?
        const uint32_t live = __ballot_sync(FULL, !done);
        uint32_t em = 0;
        if (ci * CH + lane < len) {
            const float4 a0 = q0[lane], a1 = q1[lane];
            float hA = a1.x, B = a1.y, hC = a1.z;
            if (MODE == PS_MODE_2D) ps_conic2d(a1, hA, B, hC);
            em = span_mask(a0.x, a0.y, hA, B, hC, a0.z * 1.0001f + 2.0f * THR_SLACK, bxf, byf) & live;
        }
        // lane = pixel: my candidate entries of this chunk, ascending = list order
        uint32_t pm = transpose32(em, lane);
        if (STATS) {
            const uint32_t reach = __ballot_sync(FULL, em != 0); // entries that reach a live pixel
            st_staged += 1;
            if (lane == 0) st_walk += __popc(reach);
            st_eval += __popc(pm);
        }
        uint32_t cmp = 0; // this pixel's contributing entries of the chunk (saved, transposed, for the backward)
        while (pm && !done) {
            // two candidates in flight: independent alpha chains, compositing in order
            const int ea = __ffs(pm) - 1;
            pm &= pm - 1;
            const bool two = pm != 0;
            const int eb = two ? __ffs(pm) - 1 : ea;
            pm &= pm - 1;
            const float4 r0a = q0[ea], r1a = q1[ea], r0b = q0[eb], r1b = q1[eb];
            // the colours are fetched with the rest (95 % of the candidates contribute): their latency hides behind the
            // alpha evaluation instead of stalling the compositing
            const float4 r2a = q2[ea], r2b = q2[eb];
            if (MODE == PS_MODE_3D) {
                float dx, dy;
                const float sga = ps_sigma3d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, pxf, pyf, &dx, &dy);
                const float sgb = ps_sigma3d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, pxf, pyf, &dx, &dy);
                const bool canda = sga >= 0.0f && sga <= r0a.z + THR_SLACK;
                const bool candb = two && sgb >= 0.0f && sgb <= r0b.z + THR_SLACK;
                if (!(canda || candb)) continue;
                const float aa = fminf(PS_ALPHA_MAX, psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-sga, 0x1.715476p+0f))));
                const float ab = fminf(PS_ALPHA_MAX, psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-sgb, 0x1.715476p+0f))));
                if (canda && aa >= PS_ALPHA_MIN) {
                    const float nT = psm_mul(T, psm_sub(1.0f, aa));
                    if (nT <= PS_T_STOP_3D) {
                        done = true;
                    } else {
                        const float vis = psm_mul(aa, T);
                        cr = psm_fma(vis, r2a.x, cr); cg = psm_fma(vis, r2a.y, cg); cb = psm_fma(vis, r2a.z, cb);
                        Tpen = T; T = nT; cmp |= 1u << ea;
                    }
                }
                if (candb && !done && ab >= PS_ALPHA_MIN) {
                    const float nT = psm_mul(T, psm_sub(1.0f, ab));
                    if (nT <= PS_T_STOP_3D) {
                        done = true;
                    } else {
                        const float vis = psm_mul(ab, T);
                        cr = psm_fma(vis, r2b.x, cr); cg = psm_fma(vis, r2b.y, cg); cb = psm_fma(vis, r2b.z, cb);
                        Tpen = T; T = nT; cmp |= 1u << eb;
                    }
                }
            } else {
                float dxr, dyr;
                const float qa = ps_q2d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, r1a.w, pxf, pyf, &dxr, &dyr);
                const float qb = ps_q2d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, r1b.w, pxf, pyf, &dxr, &dyr);
                const bool ina = qa <= r0a.z;
                const bool inb = two && qb <= r0b.z;
                if (!(ina || inb)) continue;
                const float gva = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-qa, 0x1.715476p+0f)));
                const float gvb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-qb, 0x1.715476p+0f)));
                if (ina) {
                    const float contrib = psm_mul(gva, T);
                    cr = psm_fma(contrib, r2a.x, cr); cg = psm_fma(contrib, r2a.y, cg); cb = psm_fma(contrib, r2a.z, cb);
                    Tpen = T; T = psm_mul(T, psm_sub(1.0f, gva)); cmp |= 1u << ea;
                    if (T <= PS_T_STOP_2D) done = true;
                }
                if (inb && !done) {
                    const float contrib = psm_mul(gvb, T);
                    cr = psm_fma(contrib, r2b.x, cr); cg = psm_fma(contrib, r2b.y, cg); cb = psm_fma(contrib, r2b.z, cb);
                    Tpen = T; T = psm_mul(T, psm_sub(1.0f, gvb)); cmp |= 1u << eb;
                    if (T <= PS_T_STOP_2D) done = true;
                }
            }
        }
        // this pixel's contributors of the chunk are the set bits of cmp: count and last position once per chunk
        // instead of once per pair
        if (cmp) { cnt += __popc(cmp); blastpos = first + 32 - __clz(cmp); }
        __syncwarp(); // lanes reconverge; every lane is finished with stage st before chunk ci + 3 is copied into it
        if (cmask) { // (pixel, entry) -> (entry, pixel); the contributing entries are appended, in list order: (id, pixel mask)
            const uint32_t cme = transpose32(cmp, lane);
            const uint32_t nzb = __ballot_sync(FULL, cme != 0u);
            if (cme != 0u) {
                const size_t o = cl_off + n_cl + __popc(nzb & ((1u << lane) - 1u));
                cids[o] = id_cur;
                cmask[o] = cme;
            }
            n_cl += __popc(nzb);
        }
        if (__all_sync(FULL, done)) break;
    }
    cp_async_wait_group<0>();
    if (ccount && lane == 0) ccount[blockIdx.x * W + wid] = n_cl;
    if (c.inside) {
        const size_t p = ((size_t)c.view * g.H + c.py) * g.W + c.px;
        const float b0 = __ldg(background), b1 = __ldg(background + 1), b2 = __ldg(background + 2);
        const float o0 = psm_fma(T, b0, cr), o1 = psm_fma(T, b1, cg), o2 = psm_fma(T, b2, cb), oa = psm_sub(1.0f, T);
        if (rgb) { rgb[3 * p + 0] = o0; rgb[3 * p + 1] = o1; rgb[3 * p + 2] = o2; }
        if (alpha) alpha[p] = oa;
        if (rgba8) rgba8[p] = quantise_rgba8(o0, o1, o2, oa);
        if (n_contrib) n_contrib[p] = cnt;
        if (last) last[p] = c.start + ((blastpos && c.bp) ? (int)__ldg(c.bp + blastpos - 1) + 1 : 0);
        if (blast) blast[p] = blastpos;
        if (t_pen) t_pen[p] = Tpen;
    }
    if (STATS) {
        unsigned long long contributing = (unsigned long long)cnt;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            st_eval += __shfl_xor_sync(FULL, st_eval, d);
            contributing += __shfl_xor_sync(FULL, contributing, d);
        }
        if (lane == 0) {
            atomicAdd(stats + 0, st_eval);
            atomicAdd(stats + 1, contributing);
            atomicAdd(stats + 2, st_walk);
            atomicAdd(stats + 3, st_staged * CH);
        }
    }
}

// ---- backward v5: two phases per group of SL contributing entries --------------------------------------------------
// Phase A (lane = pixel) replays the block list in reverse exactly like v4 but keeps only the two sequential
// per-pixel quantities: for every contributing (pixel, entry) pair it stores (alpha*T, dL/dsigma) [3D] /
// (g*T, dL/dq) [2D] into a warp-private pair table and hands the entry a slot (its record and accumulator row are
// copied to the slot by one lane; the slot's two owner lanes -- lane & 15 == slot, one per 16-pixel half of the
// block -- keep the ballot of its contributing pixels).  Two survivors are in flight per iteration (independent
// sigma / exp chains), their sequential updates are applied in list order.
// Phase B (lane = entry half) walks the set bits of its own mask: 3 colour FMAs + six moments
// (sum s, s dx, s dy, s dx^2, s dx dy, s dy^2) per contributing pair, accumulated in registers -- no warp reduction,
// and no gradient arithmetic at all for the (pixel, entry) pairs that do not contribute (84 % of them at c2).  The
// nine gradient sums of v4 are linear in these moments and are formed once per entry; rows leave through a
// shared-memory transpose so that the red.global.add of one entry's nine floats stay adjacent.
constexpr int SL = 16;              // slots (entries) per phase-B group
constexpr int PAIR_STRIDE = 33;     // float2 per slot row (32 pixels + 1 pad)

// one region, two uses that never overlap in time (warp barriers in between)
union PairTable {
    float2 pair[SL * PAIR_STRIDE];
    float out[32 * 9];
};

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE, int BW, bool STATS> // BW = warps (independent pixel blocks) per CTA
__global__ void __launch_bounds__(BW * 32, 20 / BW)
raster_bwd2_kernel(PsGeometry g, PsTable t, const uint32_t *__restrict__ vals, const int32_t *__restrict__ offsets,
                   const int32_t *__restrict__ worklist, const float *__restrict__ background, const int32_t *__restrict__ last,
                   const float *__restrict__ t_pen, const float *__restrict__ d_rgb, const float *__restrict__ d_alpha,
                   const uint32_t *__restrict__ blist, const int32_t *__restrict__ bcount, float *__restrict__ acc,
                   const int32_t *__restrict__ n_lists, unsigned *__restrict__ next_task, unsigned long long *__restrict__ stats)
{
    __shared__ float4 s_a[BW][NS][CH], s_b[BW][NS][CH], s_c[BW][NS][CH];
    __shared__ uint32_t s_id[BW][NS][CH];
    __shared__ PairTable s_pair[BW];                 // phase A -> B; reused as the 32 x 9 output transpose
    __shared__ float4 s_w[BW][32];                  // d_rgb of the block's pixels
    __shared__ float4 s_sr0[BW][SL], s_sr1[BW][SL]; // rec0 / rec1 of every slot's entry
    __shared__ uint32_t s_sid[BW][SL];              // accumulator row (view * N + Gaussian) of every slot
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long st_eval = 0, st_contrib = 0, st_walk = 0, st_staged = 0; // STATS: this kernel's own pair counters
    const unsigned n_tasks = 8u * (unsigned)__ldg(n_lists);
    // persistent warps: every warp pulls (tile, block) tasks off one counter, in work-list order (longest size class
    // first), so a warp slot is never idle while work remains (no CTA pairing, no launch gaps)
    for (;;) {
    unsigned task = 0;
    if (lane == 0) task = atomicAdd(next_task, 1u);
    task = __shfl_sync(FULL, task, 0);
    if (task >= n_tasks) break;
    const BlockCtx c = block_ctx_task(g, offsets, worklist, blist, nullptr, bcount, task);
    int my_last = 0;
    float Tcur = 1.0f, w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, S = 0.0f;
    if (c.inside) {
        const size_t p = ((size_t)c.view * g.H + c.py) * g.W + c.px;
        my_last = last[p];
        if (my_last > 0) {
            Tcur = t_pen[p];
            w0 = d_rgb[3 * p]; w1 = d_rgb[3 * p + 1]; w2 = d_rgb[3 * p + 2];
            S = __ldg(background) * w0 + __ldg(background + 1) * w1 + __ldg(background + 2) * w2 - d_alpha[p];
        }
    }
    int wmax = my_last;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) wmax = max(wmax, __shfl_xor_sync(FULL, wmax, d));
    if (wmax <= 0) continue;
    const Ring q = { s_a[wid], s_b[wid], s_c[wid] };
    uint32_t (*qid)[CH] = s_id[wid];
    float2 *pairs = s_pair[wid].pair;
    float *outt = s_pair[wid].out;
    float4 *sr0 = s_sr0[wid], *sr1 = s_sr1[wid];
    uint32_t *sid = s_sid[wid];
    s_w[wid][lane] = make_float4(w0, w1, w2, 0.0f);
    const int len = wmax;
    const int nchunks = (len + CH - 1) / CH;
    auto issue = [&](int r, uint32_t id) {
        const int cj = nchunks - 1 - r;
        if (cj >= 0 && cj * CH + lane < len) {
            const int st = r % NS;
            qid[st][lane] = id;
            const float4 *src = PS_REC(t, id, 0);
            cp_async16(&q.a[st][lane], src);
            cp_async16(&q.b[st][lane], src + 1);
            cp_async16(&q.c[st][lane], src + 2);
        }
        cp_async_commit();
    };
    auto fetch_id = [&](int r) -> uint32_t {
        const int cj = nchunks - 1 - r;
        return (cj >= 0 && cj * CH + lane < len) ? __ldg(c.bl + cj * CH + lane) : 0u;
    };
    uint32_t idn, idnn;
    {
        const uint32_t i0 = fetch_id(0), i1 = fetch_id(1);
        idn = fetch_id(2);
        idnn = fetch_id(3);
        issue(0, i0);
        issue(1, i1);
    }
    const float half = (MODE == PS_MODE_3D) ? 0.5f : 0.0f;
    const float pxf = (float)c.px + half, pyf = (float)c.py + half;
    const float bxf = (float)c.bx + half, byf = (float)c.by + half;
    const int slot_of_lane = lane & (SL - 1);
    const int pbase = lane & 16; // first pixel of this lane's half of the block
    uint32_t smask = 0;          // contributing pixels (of this lane's half) of the entry in slot `slot_of_lane`
    int nslots = 0;

    auto flush = [&]() {
        __syncwarp(); // phase A's pair / slot stores are visible
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, ms = 0.0f, mx = 0.0f, my = 0.0f, mxx = 0.0f, mxy = 0.0f, myy = 0.0f;
        uint32_t m = smask;
        const float4 e0 = sr0[slot_of_lane], e1 = sr1[slot_of_lane];
        const float sgx = e0.x - bxf, sgy = e0.y - byf; // mean relative to the block's first pixel centre
        const float2 *prow = pairs + slot_of_lane * PAIR_STRIDE + pbase;
        const float4 *wrow = s_w[wid] + pbase;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float2 pr = prow[b];
            const float4 w = wrow[b];
            const int p = pbase + b;
            // small integers -> float through the 2^23 mantissa trick (FP32 pipe instead of the conversion unit)
            const float fx = __uint_as_float(0x4b000000u | (uint32_t)(p & 7)) - 8388608.0f;
            const float fy = __uint_as_float(0x4b000000u | (uint32_t)(p >> 3)) - 8388608.0f;
            a0 = fmaf(pr.x, w.x, a0); a1 = fmaf(pr.x, w.y, a1); a2 = fmaf(pr.x, w.z, a2);
            float ex, ey;
            if (MODE == PS_MODE_3D) {
                ex = sgx - fx; ey = sgy - fy;                 // mean - pixel centre
            } else {
                const float dx = -(sgx - fx), dy = -(sgy - fy); // pixel - mean, rotated into the splat's axes
                ex = fmaf(e1.y, dy, e1.x * dx);
                ey = fmaf(e1.x, dy, -e1.y * dx);
            }
            const float tx = pr.y * ex, ty = pr.y * ey;
            ms += pr.y; mx += tx; my += ty;
            mxx = fmaf(tx, ex, mxx); mxy = fmaf(tx, ey, mxy); myy = fmaf(ty, ey, myy);
        }
        float v3, v4, v5, v6, v7, v8;
        if (MODE == PS_MODE_3D) {
            v3 = 0.5f * mxx; v4 = mxy; v5 = 0.5f * myy;
            v6 = 2.0f * e1.x * mx + e1.y * my;
            v7 = e1.y * mx + 2.0f * e1.z * my;
            v8 = -ms * rcp_approx(e0.w); // sum of exp(-sigma) * v_alpha over the unclamped pairs = -(sum v_sigma) / o
        } else {
            v3 = 2.0f * e1.z * mx; v4 = 2.0f * e1.w * my;
            v5 = 2.0f * (e1.z - e1.w) * mxy;
            v6 = mxx; v7 = myy; v8 = ms;
        }
        __syncwarp(); // every lane has read its pairs: the table becomes the output transpose
        float *o = outt + lane * 9;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = v3; o[4] = v4; o[5] = v5; o[6] = v6; o[7] = v7; o[8] = v8;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < (SL * 9 + 31) / 32; ++i) {
            const int f = i * 32 + lane;
            if (f < nslots * 9) {
                const float val = outt[f] + outt[f + SL * 9]; // the entry's two halves
                const int sl = f / 9;
                if (val != 0.0f) atomicAdd(acc + (size_t)sid[sl] * PS_ACC_STRIDE + (f - sl * 9), val);
            }
        }
        __syncwarp(); // before phase A writes pairs / slots again
        smask = 0;
        nslots = 0;
    };
    // hand the entry (ring stage st, index e) with contributor ballot cm a slot; pa / pb = this lane's pair
    auto take_slot = [&](int st, int e, const float4 &r0, const float4 &r1, uint32_t cm, bool contrib, float pa, float pb) {
        if (contrib) pairs[nslots * PAIR_STRIDE + lane] = make_float2(pa, pb);
        if (lane == 0) { sr0[nslots] = r0; sr1[nslots] = r1; sid[nslots] = qid[st][e]; }
        smask = (slot_of_lane == nslots) ? ((cm >> pbase) & 0xffffu) : smask;
        ++nslots;
    };

    for (int r = 0; r < nchunks; ++r) {
        issue(r + 2, idn);
        idn = idnn;
        idnn = fetch_id(r + 4);
        cp_async_wait_group<2>();
        __syncwarp();
        const int st = r % NS;
        const int ci = nchunks - 1 - r;
        const int first = ci * CH;
        const uint32_t live = __ballot_sync(FULL, my_last > first);
        uint32_t mask = cull_chunk<MODE>(q, st, lane, ci * CH + lane < len, c, live);
        if (STATS) { st_staged += CH; if (lane == 0) st_walk += __popc(mask); }
        const float4 *q0 = q.a[st], *q1 = q.b[st], *q2 = q.c[st];
        while (mask) {
            if (nslots > SL - 2) flush(); // room for both survivors of this iteration
            // two survivors in flight, ea (later in the list) is applied first
            const int ea = 31 - __clz(mask);
            mask &= ~(1u << ea);
            const bool two = mask != 0;
            const int eb = two ? 31 - __clz(mask) : ea;
            mask &= ~(1u << eb);
            const int posa = first + ea, posb = first + eb;
            if (STATS) st_eval += (posa < my_last) + (two && posb < my_last); // pairs whose sigma / q is evaluated
            const float4 r0a = q0[ea], r1a = q1[ea], r0b = q0[eb], r1b = q1[eb];
            float ga, gb; // alpha (3D) / g (2D) of the two pairs
            bool ca, cb;  // contributes
            float oea = 0.0f, oeb = 0.0f;
            if (MODE == PS_MODE_3D) {
                float dx, dy;
                const float sga = ps_sigma3d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, pxf, pyf, &dx, &dy);
                const float sgb = ps_sigma3d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, pxf, pyf, &dx, &dy);
                const bool canda = posa < my_last && sga >= 0.0f && sga <= r0a.z + THR_SLACK;
                const bool candb = two && posb < my_last && sgb >= 0.0f && sgb <= r0b.z + THR_SLACK;
                if (!__any_sync(FULL, canda || candb)) continue;
                oea = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-sga, 0x1.715476p+0f)));
                oeb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-sgb, 0x1.715476p+0f)));
                ga = fminf(PS_ALPHA_MAX, oea);
                gb = fminf(PS_ALPHA_MAX, oeb);
                ca = canda && ga >= PS_ALPHA_MIN;
                cb = candb && gb >= PS_ALPHA_MIN;
            } else {
                float dxr, dyr;
                const float qa = ps_q2d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, r1a.w, pxf, pyf, &dxr, &dyr);
                const float qb = ps_q2d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, r1b.w, pxf, pyf, &dxr, &dyr);
                ca = posa < my_last && qa <= r0a.z;
                cb = two && posb < my_last && qb <= r0b.z;
                if (!__any_sync(FULL, ca || cb)) continue;
                ga = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-qa, 0x1.715476p+0f)));
                gb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-qb, 0x1.715476p+0f)));
            }
            const uint32_t cma = __ballot_sync(FULL, ca), cmb = __ballot_sync(FULL, cb);
            if (STATS) st_contrib += ca + cb;
            if (!(cma | cmb)) continue;
            const float4 r2a = q2[ea], r2b = q2[eb];
            const float cwa = r2a.x * w0 + r2a.y * w1 + r2a.z * w2;
            const float cwb = r2b.x * w0 + r2b.y * w1 + r2b.z * w2;
            const float ia = rcp_approx(1.0f - ga), ib = rcp_approx(1.0f - gb); // 1 - g >= 1e-3 (3D) / > 0 for contributors
            // entry a
            const float Tba = (posa == my_last - 1) ? Tcur : Tcur * ia;
            const float va = Tba * (cwa - S);          // dL/dalpha (3D) / dL/dg (2D)
            const float paa = ga * Tba;
            const float pba = (MODE == PS_MODE_3D) ? ((oea <= PS_ALPHA_MAX) ? -oea * va : 0.0f) : -ga * va;
            Tcur = ca ? Tba : Tcur;
            S = ca ? S + ga * (cwa - S) : S;
            // entry b
            const float Tbb = (posb == my_last - 1) ? Tcur : Tcur * ib;
            const float vb = Tbb * (cwb - S);
            const float pab = gb * Tbb;
            const float pbb = (MODE == PS_MODE_3D) ? ((oeb <= PS_ALPHA_MAX) ? -oeb * vb : 0.0f) : -gb * vb;
            Tcur = cb ? Tbb : Tcur;
            S = cb ? S + gb * (cwb - S) : S;
            if (cma) take_slot(st, ea, r0a, r1a, cma, ca, paa, pba);
            if (cmb) take_slot(st, eb, r0b, r1b, cmb, cb, pab, pbb);
        }
        __syncwarp(); // every lane is finished with stage st before step r + 3 is copied into it
    }
    cp_async_wait_group<0>();
    if (nslots) flush();
    __syncwarp();
    } // task loop
    if (STATS) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            st_eval += __shfl_xor_sync(FULL, st_eval, d);
            st_contrib += __shfl_xor_sync(FULL, st_contrib, d);
        }
        if (lane == 0) {
            atomicAdd(stats + 4, st_eval);
            atomicAdd(stats + 5, st_contrib);
            atomicAdd(stats + 6, st_walk);
            atomicAdd(stats + 7, st_staged);
        }
    }
}

// ---- backward v6: contributor masks from the forward, per-pixel chain, per-entry moments ------------------------------
// The forward leaves, per pixel block, the list of the entries it composited into at least one of the block's pixels, each
// with the 32-bit mask of those pixels (cids / cmask / ccount: the contributor list, about half of the block list at c2).
// The replay therefore touches contributing (pixel, entry) pairs only -- no culling, no candidate tests, no ballots:
//   A (lane = pixel)  the chunk's 32 entry masks are transposed across the warp (five shuffles); every pixel lane walks
//                     ITS OWN contributors in reverse list order, recomputes alpha with the arithmetic of the contract and
//                     advances the two sequential per-pixel quantities (T by rcp.approx from the saved "T before the last
//                     contributor", S = sum behind); it leaves (alpha T, dL/dsigma) [3D] / (g T, dL/dq) [2D] per pair in a
//                     32 x 32 table in shared memory
//   B (lane = entry)  every entry lane walks the set bits of its own mask: 3 colour FMAs + six moments per pair in
//                     registers; the nine gradient sums are linear in them, formed once per entry, transposed through
//                     shared memory so that one entry's nine floats leave with adjacent red.global.add
// Persistent warps pulling (tile, block) tasks off one counter, records streamed through the cp.async ring as before.
constexpr int TAB_STRIDE = 33; // float2 per table row (32 pixels + 1 pad)

// BULK = true: the records are staged with one 48-byte cp.async.bulk per entry (TMA engine, 1-D form) completing on a
// per-stage mbarrier instead of three 16-byte cp.async per lane (A/B experiment).
template <int MODE, int BW, bool STATS, bool BULK>
__global__ void __launch_bounds__(BW * 32, 16 / BW)
raster_bwd3_kernel(PsGeometry g, PsTable t, const int32_t *__restrict__ offsets, const int32_t *__restrict__ worklist,
                   const float *__restrict__ background, const int32_t *__restrict__ last, const float *__restrict__ t_pen,
                   const float *__restrict__ d_rgb, const float *__restrict__ d_alpha, const uint32_t *__restrict__ blist,
                   const uint32_t *__restrict__ cids, const uint32_t *__restrict__ cmask, const int32_t *__restrict__ ccount,
                   const int32_t *__restrict__ bcount, float *__restrict__ acc,
                   const int32_t *__restrict__ n_lists, unsigned *__restrict__ next_task, unsigned long long *__restrict__ stats)
{
    // cp.async: three planes of float4 (rec0 | rec1 | rec2); bulk: 48 contiguous bytes per entry (s_a[w][st][3 e + k])
    __shared__ __align__(16) float4 s_a[BW][NS][BULK ? 3 * CH : CH], s_b[BW][NS][BULK ? 1 : CH], s_c[BW][NS][BULK ? 1 : CH];
    __shared__ __align__(8) uint64_t s_bar[BW][NS];
    __shared__ float2 s_tab[BW][32 * TAB_STRIDE];   // phase A -> B pair table [entry][pixel]; reused as the 32 x 9 output transpose
    __shared__ float4 s_w[BW][32];                  // d_rgb of the block's pixels
    __shared__ uint32_t s_id[BW][32];               // accumulator rows (view * N + Gaussian) of the chunk's entries
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long st_pairs = 0, st_walk = 0, st_staged = 0;
    const unsigned n_tasks = 8u * (unsigned)__ldg(n_lists);
    if (BULK) {
        if (lane < NS) mbar_init(&s_bar[wid][lane], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    unsigned bar_phase = 0; // bit st: parity the next wait on stage st must see
    float2 *tab = s_tab[wid];
    float *outt = reinterpret_cast<float *>(s_tab[wid]);
    for (;;) {
    unsigned task = 0;
    if (lane == 0) task = atomicAdd(next_task, 1u);
    task = __shfl_sync(FULL, task, 0);
    if (task >= n_tasks) break;
    const BlockCtx c = block_ctx_task(g, offsets, worklist, blist, nullptr, bcount, task);
    int my_last = 0;
    float Tcur = 1.0f, w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, S = 0.0f;
    if (c.inside) {
        const size_t p = ((size_t)c.view * g.H + c.py) * g.W + c.px;
        my_last = last[p];
        if (my_last > 0) {
            Tcur = t_pen[p];
            w0 = d_rgb[3 * p]; w1 = d_rgb[3 * p + 1]; w2 = d_rgb[3 * p + 2];
            S = __ldg(background) * w0 + __ldg(background + 1) * w1 + __ldg(background + 2) * w2 - d_alpha[p];
        }
    }
    // the forward's contributor list of this block: the entries composited into at least one of its pixels
    const int len = __ldg(ccount + task);
    if (len <= 0) continue;
    __syncwarp();
    s_w[wid][lane] = make_float4(w0, w1, w2, 0.0f);
    const int nchunks = (len + CH - 1) / CH;
    const uint32_t *cil = cids + (c.bl - blist);
    const uint32_t *cml = cmask + (c.bl - blist);
    // reverse step r handles chunk nchunks - 1 - r; its ring stage is r % NS; ids and masks run ahead in registers
    auto issue = [&](int r, uint32_t id) {
        const int cj = nchunks - 1 - r;
        const int st = r % NS;
        if (BULK) {
            if (cj < 0) return;
            const int n_valid = min(CH, len - cj * CH);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage's last generic-proxy reads precede the async writes
            if (lane == 0) mbar_expect_tx(&s_bar[wid][st], 48u * (unsigned)n_valid);
            __syncwarp(); // the transaction count is armed before any copy can complete on the barrier
            if (lane < n_valid) bulk_copy_g2s(&s_a[wid][st][3 * lane], PS_REC(t, id, 0), 48u, &s_bar[wid][st]);
            return;
        }
        if (cj >= 0 && cj * CH + lane < len) {
            const float4 *src = PS_REC(t, id, 0);
            cp_async16(&s_a[wid][st][lane], src);
            cp_async16(&s_b[wid][st][lane], src + 1);
            cp_async16(&s_c[wid][st][lane], src + 2);
        }
        cp_async_commit();
    };
    auto fetch = [&](int r, uint32_t &id, uint32_t &cm) {
        const int cj = nchunks - 1 - r;
        const bool ok = cj >= 0 && cj * CH + lane < len;
        id = ok ? __ldg(cil + cj * CH + lane) : 0u;
        cm = ok ? __ldg(cml + cj * CH + lane) : 0u;
    };
    uint32_t id0, cm0, id1, cm1, id2, cm2, id3, cm3; // steps r, r + 1, r + 2, r + 3
    fetch(0, id0, cm0); fetch(1, id1, cm1); fetch(2, id2, cm2); fetch(3, id3, cm3);
    issue(0, id0);
    issue(1, id1);
    const float half = (MODE == PS_MODE_3D) ? 0.5f : 0.0f;
    const float pxf = (float)c.px + half, pyf = (float)c.py + half;
    const float bxf = (float)c.bx + half, byf = (float)c.by + half;
    bool first_c = true; // the next contributor met is this pixel's last one: its T is the saved t_pen itself

    for (int r = 0; r < nchunks; ++r) {
        issue(r + 2, id2);
        const uint32_t id_c = id0, cme = cm0;
        id0 = id1; cm0 = cm1; id1 = id2; cm1 = cm2; id2 = id3; cm2 = cm3;
        fetch(r + 4, id3, cm3);
        const int st = r % NS;
        if (BULK) {
            mbar_wait(&s_bar[wid][st], (bar_phase >> st) & 1u);
            bar_phase ^= 1u << st;
        } else {
            cp_async_wait_group<2>();
        }
        __syncwarp();
        if (STATS) st_staged += CH;
        const uint32_t nz = __ballot_sync(FULL, cme != 0u);
        if (nz == 0u) continue; // no pixel of the block composited any entry of this chunk
        if (STATS && lane == 0) st_walk += __popc(nz);
        // record words of entry e: cp.async planes q0[e], q1[e], q2[e]; bulk rows q0[3 e], q0[3 e + 1], q0[3 e + 2]
        const float4 *q0 = s_a[wid][st], *q1 = BULK ? s_a[wid][st] + 1 : s_b[wid][st], *q2 = BULK ? s_a[wid][st] + 2 : s_c[wid][st];
        constexpr int RS = BULK ? 3 : 1; // float4 stride between consecutive entries
        s_id[wid][lane] = id_c;
        // ---- phase A: lane = pixel
        uint32_t pm = transpose32(cme, lane);
        if (STATS) st_pairs += __popc(pm);
        while (pm) {
            // two contributors in flight (independent sigma / exp chains), applied in reverse list order: a, then b
            const int ea = 31 - __clz(pm);
            pm &= ~(1u << ea);
            const bool two = pm != 0u;
            const int eb = two ? 31 - __clz(pm) : ea;
            pm &= ~(1u << eb);
            const float4 r0a = q0[RS * ea], r1a = q1[RS * ea], r2a = q2[RS * ea], r0b = q0[RS * eb], r1b = q1[RS * eb], r2b = q2[RS * eb];
            float ga, gb, sa, sb; // alpha (3D) / g (2D); factor of -v in dL/dsigma resp. dL/dq
            if (MODE == PS_MODE_3D) {
                float dx, dy;
                const float sga = ps_sigma3d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, pxf, pyf, &dx, &dy);
                const float sgb = ps_sigma3d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, pxf, pyf, &dx, &dy);
                const float oea = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-sga, 0x1.715476p+0f)));
                const float oeb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-sgb, 0x1.715476p+0f)));
                ga = fminf(PS_ALPHA_MAX, oea); gb = fminf(PS_ALPHA_MAX, oeb);
                sa = (oea <= PS_ALPHA_MAX) ? oea : 0.0f; sb = (oeb <= PS_ALPHA_MAX) ? oeb : 0.0f;
            } else {
                float dxr, dyr;
                const float qa = ps_q2d(r0a.x, r0a.y, r1a.x, r1a.y, r1a.z, r1a.w, pxf, pyf, &dxr, &dyr);
                const float qb = ps_q2d(r0b.x, r0b.y, r1b.x, r1b.y, r1b.z, r1b.w, pxf, pyf, &dxr, &dyr);
                ga = psm_mul(r0a.w, psm_exp2_inrange(psm_mul(-qa, 0x1.715476p+0f)));
                gb = psm_mul(r0b.w, psm_exp2_inrange(psm_mul(-qb, 0x1.715476p+0f)));
                sa = ga; sb = gb;
            }
            const float cwa = r2a.x * w0 + r2a.y * w1 + r2a.z * w2;
            const float cwb = r2b.x * w0 + r2b.y * w1 + r2b.z * w2;
            const float ia = rcp_approx(1.0f - ga), ib = rcp_approx(1.0f - gb); // 1 - g >= 1e-3 (3D) / > 0 for contributors
            const float Tba = first_c ? Tcur : Tcur * ia;
            first_c = false;
            const float va = Tba * (cwa - S); // dL/dalpha (3D) / dL/dg (2D)
            tab[ea * TAB_STRIDE + lane] = make_float2(ga * Tba, -sa * va);
            Tcur = Tba;
            S = S + ga * (cwa - S);
            if (two) {
                const float Tbb = Tcur * ib;
                const float vb = Tbb * (cwb - S);
                tab[eb * TAB_STRIDE + lane] = make_float2(gb * Tbb, -sb * vb);
                Tcur = Tbb;
                S = S + gb * (cwb - S);
            }
        }
        __syncwarp(); // the pair table is complete
        // ---- phase B: lane = entry
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, ms = 0.0f, mx = 0.0f, my = 0.0f, mxx = 0.0f, mxy = 0.0f, myy = 0.0f;
        const float4 e0 = q0[RS * lane], e1 = q1[RS * lane];
        {
            uint32_t m = cme;
            const float sgx = e0.x - bxf, sgy = e0.y - byf; // mean relative to the block's first pixel centre
            const float2 *prow = tab + lane * TAB_STRIDE;
            const float4 *wrow = s_w[wid];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const float2 pr = prow[b];
                const float4 w = wrow[b];
                // small integers -> float through the 2^23 mantissa trick (FP32 pipe instead of the conversion unit)
                const float fx = __uint_as_float(0x4b000000u | (uint32_t)(b & 7)) - 8388608.0f;
                const float fy = __uint_as_float(0x4b000000u | (uint32_t)(b >> 3)) - 8388608.0f;
                a0 = fmaf(pr.x, w.x, a0); a1 = fmaf(pr.x, w.y, a1); a2 = fmaf(pr.x, w.z, a2);
                float ex, ey;
                if (MODE == PS_MODE_3D) {
                    ex = sgx - fx; ey = sgy - fy;                 // mean - pixel centre
                } else {
                    const float dx = -(sgx - fx), dy = -(sgy - fy); // pixel - mean, rotated into the splat's axes
                    ex = fmaf(e1.y, dy, e1.x * dx);
                    ey = fmaf(e1.x, dy, -e1.y * dx);
                }
                const float tx = pr.y * ex, ty = pr.y * ey;
                ms += pr.y; mx += tx; my += ty;
                mxx = fmaf(tx, ex, mxx); mxy = fmaf(tx, ey, mxy); myy = fmaf(ty, ey, myy);
            }
        }
        float v3, v4, v5, v6, v7, v8;
        if (MODE == PS_MODE_3D) {
            v3 = 0.5f * mxx; v4 = mxy; v5 = 0.5f * myy;
            v6 = 2.0f * e1.x * mx + e1.y * my;
            v7 = e1.y * mx + 2.0f * e1.z * my;
            v8 = -ms * rcp_approx(e0.w); // sum of exp(-sigma) * v_alpha over the unclamped pairs = -(sum v_sigma) / o
        } else {
            v3 = 2.0f * e1.z * mx; v4 = 2.0f * e1.w * my;
            v5 = 2.0f * (e1.z - e1.w) * mxy;
            v6 = mxx; v7 = myy; v8 = ms;
        }
        __syncwarp(); // every lane has read its pairs: the table becomes the output transpose
        {
            float *o = outt + lane * 9;
            o[0] = a0; o[1] = a1; o[2] = a2; o[3] = v3; o[4] = v4; o[5] = v5; o[6] = v6; o[7] = v7; o[8] = v8;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int f = i * 32 + lane;
            const int en = f / 9;
            if ((nz >> en) & 1u) {
                const float val = outt[f];
                if (val != 0.0f) atomicAdd(acc + (size_t)s_id[wid][en] * PS_ACC_STRIDE + (f - en * 9), val);
            }
        }
        __syncwarp(); // before the next chunk's phase A writes the table / s_id again; ring stage st is free as well
    }
    if (!BULK) cp_async_wait_group<0>();
    __syncwarp();
    } // task loop
    if (STATS) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) st_pairs += __shfl_xor_sync(FULL, st_pairs, d);
        if (lane == 0) {
            atomicAdd(stats + 4, st_pairs); // every pair replayed is a contributing pair
            atomicAdd(stats + 5, st_pairs);
            atomicAdd(stats + 6, st_walk);
            atomicAdd(stats + 7, st_staged);
        }
    }
}

// Split every non-empty tile list, in order, into the lists of its eight 8x4 pixel blocks.
// m8s != NULL: the block masks were computed by the partition kernel and sorted along (one byte per list entry):
// a pure streaming split of (id, mask) pairs, no record is touched.
template <int MODE, bool BPOS> // BPOS: also write the tile-list positions (last-id parity tap only)
__global__ void __launch_bounds__(256)
block_lists_kernel(PsGeometry g, PsTable t, const uint32_t *__restrict__ vals, const int32_t *__restrict__ offsets,
                   const int32_t *__restrict__ worklist, uint32_t *__restrict__ blist, uint32_t *__restrict__ bpos,
                   int32_t *__restrict__ bcount, const uint8_t *__restrict__ m8s, const int32_t *__restrict__ n_lists)
{
    if ((int)blockIdx.x >= __ldg(n_lists)) return; // the grid may be an upper bound (sync-free small calls)
    __shared__ int s_cnt[8][8]; // [warp][block]
    __shared__ __align__(16) int s_pre[8][8]; // [warp][block] output cursor of the round
    __shared__ int s_run[8], s_tot[8];
    const int item = blockIdx.x;
    const int lin = worklist[item];
    const int view = lin / g.n_tiles, tile = lin - view * g.n_tiles;
    const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
    const int start = offsets[lin], len = offsets[lin + 1] - start;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *out = blist + 8 * (size_t)start;
    uint32_t *outp = BPOS ? bpos + 8 * (size_t)start : nullptr;
    const uint32_t *list = vals + start;
    uint32_t inside8 = 0; // blocks that have at least one pixel inside the image
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (tx * PS_TILE + (k & 1) * 8 < g.W && ty * PS_TILE + (k >> 1) * 4 < g.H) inside8 |= 1u << k;
    if (threadIdx.x < 8) s_run[threadIdx.x] = 0;
    // software pipeline: ids two rounds ahead, records one round ahead
    const int tid = threadIdx.x;
    const uint8_t *mlist = m8s ? m8s + start : nullptr;
    uint32_t id_n = (tid < len) ? __ldg(list + tid) : 0u;
    uint32_t id_nn = (256 + tid < len) ? __ldg(list + 256 + tid) : 0u;
    uint32_t mk_n = (mlist && tid < len) ? mlist[tid] : 0u;
    float4 r0_n = make_float4(0.f, 0.f, 0.f, 0.f), r1_n = r0_n;
    if (!mlist) { r0_n = __ldg(PS_REC(t, id_n, 0)); r1_n = __ldg(PS_REC(t, id_n, 1)); }
    __syncthreads();
    for (int first = 0; first < len; first += 256) {
        const int j = first + tid;
        const float4 r0 = r0_n, r1 = r1_n;
        const uint32_t id_cur = id_n;
        uint32_t m8 = mk_n;
        {
            const uint32_t id_next = id_nn;
            id_n = id_next;
            id_nn = (first + 512 + tid < len) ? __ldg(list + first + 512 + tid) : 0u;
            if (mlist) mk_n = (first + 256 + tid < len) ? mlist[first + 256 + tid] : 0u;
            else { r0_n = __ldg(PS_REC(t, id_next, 0)); r1_n = __ldg(PS_REC(t, id_next, 1)); }
        }
        if (!mlist) {
            m8 = 0;
            if (j < len) {
                float hA = r1.x, B = r1.y, hC = r1.z;
                if (MODE == PS_MODE_2D) ps_conic2d(r1, hA, B, hC);
                m8 = ps_block_mask8(r0.x, r0.y, hA, B, hC, r0.z, (MODE == PS_MODE_3D) ? 0.5f : 0.0f, tx, ty);
                m8 &= inside8;
            }
        }
        uint32_t bal[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bal[k] = __ballot_sync(FULL, (m8 >> k) & 1u);
            if (lane == 0) s_cnt[wid][k] = __popc(bal[k]);
        }
        __syncthreads();
        if (threadIdx.x < 64) { // output cursor of every (warp, block) for this round: 64 threads, 8 adds each
            const int w = threadIdx.x >> 3, k = threadIdx.x & 7;
            int pre = s_run[k];
#pragma unroll
            for (int w2 = 0; w2 < 8; ++w2) pre += (w2 < w) ? s_cnt[w2][k] : 0;
            s_pre[w][k] = pre;
            if (w == 7) s_tot[k] = pre + s_cnt[7][k]; // running total after this round
        }
        __syncthreads();
        {
            // the warp's eight cursors in two 16-byte loads; offsets in 32 bits (8 * len < 2^27: a tile list holds at
            // most one entry per Gaussian of the view, and N < 2^24)
            const int4 pa = *reinterpret_cast<const int4 *>(&s_pre[wid][0]), pb = *reinterpret_cast<const int4 *>(&s_pre[wid][4]);
            const int pre[8] = { pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w };
            const uint32_t below = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((m8 >> k) & 1u) {
                    const uint32_t o = (uint32_t)k * (uint32_t)len + (uint32_t)pre[k] + (uint32_t)__popc(bal[k] & below);
                    out[o] = id_cur;
                    if (BPOS) outp[o] = (uint32_t)j;
                }
            }
        }
        if (threadIdx.x < 8) s_run[threadIdx.x] = s_tot[threadIdx.x];
        // s_cnt is rewritten by the next round only after every thread has passed the second barrier above, and
        // s_pre / s_run are read by it only after its own first barrier
    }
    __syncthreads();
    if (threadIdx.x < 8) bcount[item * 8 + threadIdx.x] = s_run[threadIdx.x];
}

} // namespace

int ps_launch_block_lists(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const uint8_t *m8s, cudaStream_t s)
{
    if (n_work <= 0) return 0;
#define PS_BL(MODE, BPOS) block_lists_kernel<MODE, BPOS><<<n_work, 256, 0, s>>>(g, t, l.vals, l.offsets, l.worklist, l.blist, l.bpos, l.bcount, m8s, l.n_lists)
    if (g.mode == PS_MODE_3D) { if (l.bpos) PS_BL(PS_MODE_3D, true); else PS_BL(PS_MODE_3D, false); }
    else { if (l.bpos) PS_BL(PS_MODE_2D, true); else PS_BL(PS_MODE_2D, false); }
#undef PS_BL
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_raster_fwd(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const float *background,
                         float *rgb, float *alpha, int32_t *n_contrib, int32_t *last, int32_t *blast, float *t_pen,
                         uint32_t *rgba8, unsigned long long *stats, cudaStream_t s)
{
    if (n_work <= 0) return 0;
    const unsigned grid = (unsigned)n_work * TASKS_PER_TILE;
    // 3D: v6 (candidate masks + per-pixel walk; 3.22 -> 3.04 ms at c2).  2D footprints cover half a block, where the
    // all-lanes walk of v4 is faster (5.25 vs 5.44 ms at c3).  With `stats` the SAME kernel runs with its pair
    // counters compiled in (bench.py's roofline numerator comes from the kernel it times).
    static const bool force_v4 = getenv("PS_FWD_V4") != nullptr; // A/B switch for measurements
#define PS_FWD(K, MODE, ST) K<MODE, ST, WPC><<<grid, RT_THREADS, 0, s>>>(g, t, l.vals, l.offsets, l.worklist, background, rgb, alpha, n_contrib, last, blast, t_pen, l.blist, l.bpos, l.bcount, rgba8, stats, l.n_lists, l.cmask, l.cids, l.ccount)
    if (g.mode == PS_MODE_3D && !force_v4) {
        if (stats) PS_FWD(raster_fwd6_kernel, PS_MODE_3D, true); else PS_FWD(raster_fwd6_kernel, PS_MODE_3D, false);
    } else if (g.mode == PS_MODE_3D) {
        if (stats) PS_FWD(raster_fwd_kernel, PS_MODE_3D, true); else PS_FWD(raster_fwd_kernel, PS_MODE_3D, false);
    } else {
        if (stats) PS_FWD(raster_fwd_kernel, PS_MODE_2D, true); else PS_FWD(raster_fwd_kernel, PS_MODE_2D, false);
    }
#undef PS_FWD
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

namespace {
// 8 independent FFMA chains per thread; 2 flops per FFMA
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *sink, int iters)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + 1e-3f * (float)(threadIdx.x + i);
    const float m = 0.999f, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], m, c);
        }
    }
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456f) sink[0] = r;
}
} // namespace

int ps_launch_fp32_probe(float *sink, int iters, cudaStream_t s)
{
    fp32_probe_kernel<<<148 * 8, 256, 0, s>>>(sink, iters);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

namespace {
// persistent grid of a backward kernel on the current device: resident CTAs per SM x SMs (full shared-memory carve-out)
template <typename K>
unsigned persistent_ctas(K kernel, int slot)
{
    constexpr int MAX_DEV = 64;
    static int cached[MAX_DEV][10] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = dev < MAX_DEV ? dev : MAX_DEV - 1;
    if (!cached[di][slot]) {
        int per = 0, n_sm = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, RT_THREADS, 0);
        cached[di][slot] = n_sm * (per > 0 ? per : 1);
    }
    return (unsigned)cached[di][slot];
}
} // namespace

int ps_launch_raster_bwd(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const float *background,
                         const int32_t *last, const float *t_pen, const float *d_rgb, const float *d_alpha, float *acc,
                         unsigned *next_task, unsigned long long *stats, cudaStream_t s)
{
    if (n_work <= 0) return 0;
    const unsigned n_tasks = (unsigned)n_work * 8u; // n_work may be an upper bound: the kernels read the exact count
    const unsigned want = (n_tasks + WPC - 1) / WPC;
    static const bool use_v5 = getenv("PS_BWD_V5") != nullptr; // A/B switch for measurements
    const int mi = g.mode == PS_MODE_3D ? 0 : 1;
    if (l.cmask && !use_v5) { // v6: replay of the contributing pairs the forward recorded
#define PS_BWD3(MODE, ST, BK, SLOT)                                                                                          \
        do {                                                                                                                 \
            const unsigned cap = persistent_ctas(raster_bwd3_kernel<MODE, WPC, ST, BK>, SLOT);                               \
            raster_bwd3_kernel<MODE, WPC, ST, BK><<<want < cap ? want : cap, RT_THREADS, 0, s>>>(g, t, l.offsets, l.worklist, background, last, t_pen, d_rgb, d_alpha, l.blist, l.cids, l.cmask, l.ccount, l.bcount, acc, l.n_lists, next_task, stats); \
        } while (0)
        static const bool bulk = getenv("PS_BWD_BULK") != nullptr; // A/B switch: records staged by cp.async.bulk (TMA 1-D)
        if (bulk && !stats) { if (mi == 0) PS_BWD3(PS_MODE_3D, false, true, 8); else PS_BWD3(PS_MODE_2D, false, true, 9); }
        else if (mi == 0) { if (stats) PS_BWD3(PS_MODE_3D, true, false, 4); else PS_BWD3(PS_MODE_3D, false, false, 5); }
        else { if (stats) PS_BWD3(PS_MODE_2D, true, false, 6); else PS_BWD3(PS_MODE_2D, false, false, 7); }
#undef PS_BWD3
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
#define PS_BWD(MODE, ST, SLOT)                                                                                               \
    do {                                                                                                                     \
        const unsigned cap = persistent_ctas(raster_bwd2_kernel<MODE, WPC, ST>, SLOT);                                       \
        raster_bwd2_kernel<MODE, WPC, ST><<<want < cap ? want : cap, RT_THREADS, 0, s>>>(g, t, l.vals, l.offsets, l.worklist, background, last, t_pen, d_rgb, d_alpha, l.blist, l.bcount, acc, l.n_lists, next_task, stats); \
    } while (0)
    if (mi == 0) { if (stats) PS_BWD(PS_MODE_3D, true, 0); else PS_BWD(PS_MODE_3D, false, 1); }
    else { if (stats) PS_BWD(PS_MODE_2D, true, 2); else PS_BWD(PS_MODE_2D, false, 3); }
#undef PS_BWD
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Pixels of tiles whose list is empty: rgb = background, alpha = 0 (what the compositing loop yields for zero
// contributors: fma(1, bg, 0) and 1 - 1).  One thread per 4 horizontally adjacent pixels (same tile).
namespace {
__global__ void __launch_bounds__(256)
fill_empty_kernel(PsGeometry g, const int32_t *__restrict__ offsets, const float *__restrict__ background,
                  float *__restrict__ rgb, float *__restrict__ alpha, int32_t *__restrict__ n_contrib,
                  int32_t *__restrict__ last, uint32_t *__restrict__ rgba8, int groups_x, long long n_groups, int vec_ok)
{
    const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= n_groups) return;
    const int gx = (int)(gidx % groups_x);
    const long long row = gidx / groups_x; // view * H + y
    const int y = (int)(row % g.H), v = (int)(row / g.H);
    const int x0 = gx * 4;
    const int lin = v * g.n_tiles + (y / PS_TILE) * g.tiles_x + x0 / PS_TILE;
    const int start = offsets[lin];
    if (offsets[lin + 1] != start) return;
    const float b0 = __ldg(background), b1 = __ldg(background + 1), b2 = __ldg(background + 2);
    const size_t p = (size_t)row * g.W + x0;
    const int n = min(4, g.W - x0);
    const uint32_t q8 = quantise_rgba8(b0, b1, b2, 0.0f);
    if (n == 4 && vec_ok) {
        if (rgb) {
            float4 *d = reinterpret_cast<float4 *>(rgb + 3 * p);
            d[0] = make_float4(b0, b1, b2, b0);
            d[1] = make_float4(b1, b2, b0, b1);
            d[2] = make_float4(b2, b0, b1, b2);
        }
        if (alpha) *reinterpret_cast<float4 *>(alpha + p) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rgba8) *reinterpret_cast<uint4 *>(rgba8 + p) = make_uint4(q8, q8, q8, q8);
        if (n_contrib) *reinterpret_cast<int4 *>(n_contrib + p) = make_int4(0, 0, 0, 0);
        if (last) *reinterpret_cast<int4 *>(last + p) = make_int4(start, start, start, start);
    } else {
        for (int k = 0; k < n; ++k) {
            if (rgb) { rgb[3 * (p + k)] = b0; rgb[3 * (p + k) + 1] = b1; rgb[3 * (p + k) + 2] = b2; }
            if (alpha) alpha[p + k] = 0.0f;
            if (rgba8) rgba8[p + k] = q8;
            if (n_contrib) n_contrib[p + k] = 0;
            if (last) last[p + k] = start;
        }
    }
}
} // namespace

int ps_launch_fill_empty(const PsGeometry &g, const int32_t *offsets, const float *background, float *rgb, float *alpha,
                         int32_t *n_contrib, int32_t *last, uint32_t *rgba8, cudaStream_t s)
{
    const int groups_x = (g.W + 3) / 4;
    const long long n_groups = (long long)g.V * g.H * groups_x;
    if (n_groups == 0) return 0;
    const uintptr_t align = (uintptr_t)rgb | (uintptr_t)alpha | (uintptr_t)n_contrib | (uintptr_t)last | (uintptr_t)rgba8;
    const int vec_ok = (g.W & 3) == 0 && (align & 15u) == 0;
    fill_empty_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(g, offsets, background, rgb, alpha, n_contrib, last,
                                                                      rgba8, groups_x, n_groups, vec_ok);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
