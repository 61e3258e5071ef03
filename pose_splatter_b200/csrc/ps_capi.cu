// ps_capi.cu -- C ABI of libpsplat.so (include/psplat.h): context, cached device scratch arena,
// and the stage orchestration  project(+tile histogram) -> depth rank -> scan [M to host] ->
// partition -> per-list bitmap sort -> fill empty tiles -> rasterize  (forward)  /
// rasterize-backward -> projection-backward.
// No torch types, no CPU fallback: every entry point needs a CUDA device.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include "ps_contract.cuh"
#include "ps_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define PS_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define PS_LAUNCH(ctx, call)                                                   \
    do {                                                                       \
        int n_ = (call);                                                       \
        if (n_ < 0) return fail(3, "kernel launch failed in %s: %s", #call, cudaGetErrorString(cudaGetLastError())); \
        (ctx)->launches += n_;                                                 \
    } while (0)

// Device scratch arena: a per-device cache of cudaMalloc'ed blocks in geometric size classes (8 per octave), so the
// steady state of a render loop allocates nothing (the list sizes change a little from step to step; the
// driver's stream-ordered pool turned out to re-map memory on many calls, costing milliseconds of host time).
// A cached block is handed out again only for work on the stream it was last used on (stream order makes the
// reuse safe); a block last used on another stream is reused after synchronising that stream.
struct ArenaBlock { void *p; size_t cap; cudaStream_t stream; };
struct Arena {
    std::mutex mu;
    std::vector<ArenaBlock> free_blocks;
    std::unordered_map<void *, size_t> live; // pointer -> capacity
    size_t cached_bytes = 0;
};
constexpr int PS_MAX_DEVICES = 64;
Arena g_arena[PS_MAX_DEVICES];
constexpr size_t ARENA_CACHE_LIMIT = (size_t)48 << 30; // beyond this many cached bytes everything cached is released

size_t arena_class(size_t bytes)
{
    if (bytes < 4096) return 4096;
    int hb = 63 - __builtin_clzll((unsigned long long)bytes);
    const size_t step = (size_t)1 << (hb - 3); // 8 classes per octave: at most 12.5 % slack
    return (bytes + step - 1) & ~(step - 1);
}

void arena_release_cached(Arena &a)
{
    for (auto &b : a.free_blocks) { cudaStreamSynchronize(b.stream); cudaFree(b.p); }
    a.free_blocks.clear();
    a.cached_bytes = 0;
}

cudaError_t arena_alloc(void **p, size_t bytes, cudaStream_t s)
{
    *p = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= PS_MAX_DEVICES) return cudaErrorInvalidDevice;
    Arena &a = g_arena[dev];
    const size_t cap = arena_class(bytes);
    std::lock_guard<std::mutex> lock(a.mu);
    int pick = -1;
    for (int i = (int)a.free_blocks.size() - 1; i >= 0; --i) {
        if (a.free_blocks[i].cap != cap) continue;
        if (a.free_blocks[i].stream == s) { pick = i; break; }
        if (pick < 0) pick = i;
    }
    if (pick >= 0) {
        ArenaBlock b = a.free_blocks[pick];
        a.free_blocks.erase(a.free_blocks.begin() + pick);
        a.cached_bytes -= b.cap;
        if (b.stream != s) cudaStreamSynchronize(b.stream);
        a.live[b.p] = b.cap;
        *p = b.p;
        return cudaSuccess;
    }
    e = cudaMalloc(p, cap);
    if (e != cudaSuccess) { // out of memory: drop the cache and retry once
        cudaGetLastError();
        arena_release_cached(a);
        e = cudaMalloc(p, cap);
        if (e != cudaSuccess) return e;
    }
    a.live[*p] = cap;
    return cudaSuccess;
}

void arena_free(void *p, cudaStream_t s)
{
    if (!p) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PS_MAX_DEVICES) return;
    Arena &a = g_arena[dev];
    std::lock_guard<std::mutex> lock(a.mu);
    auto it = a.live.find(p);
    if (it == a.live.end()) return;
    const size_t cap = it->second;
    a.live.erase(it);
    a.free_blocks.push_back({ p, cap, s });
    a.cached_bytes += cap;
    if (a.cached_bytes > ARENA_CACHE_LIMIT) arena_release_cached(a);
}

template <typename T>
cudaError_t dev_alloc(T **p, size_t count, cudaStream_t s)
{
    if (count == 0) count = 1;
    return arena_alloc((void **)p, count * sizeof(T), s);
}

template <typename T>
void dev_free(T *&p, cudaStream_t s)
{
    arena_free((void *)p, s);
    p = nullptr;
}

} // namespace

struct PsSpan { int stage; cudaEvent_t a, b; };

// One forward's mailbox: {M, number of non-empty lists} stored by the scan kernel straight into mapped pinned host
// memory (h = host view, d = device view of the same words).  No copy-engine transfer: a 16-byte memcpy would queue
// behind whatever bulk device->host copy the application has in flight on another stream.  Every forward takes its
// own slot (concurrent calls on one context never share one); like the arena blocks, a returned slot is reused on the
// stream it was last written on, or after that stream has drained.
struct PsMailbox { volatile int64_t *h; int64_t *d; cudaStream_t stream; };

struct ps_ctx {
    int device;
    std::atomic<int64_t> launches;
    std::mutex mu;                       // guards everything below
    std::vector<PsMailbox> mail_free;
    std::vector<void *> mail_chunks;     // pinned allocations the slots live in
    unsigned long long *d_stats;         // [8] pair counters (PS_FLAG_RASTER_STATS): forward [0..3], backward [4..7]
    cudaStream_t side;                   // the background fill of large forwards runs here, beside the binning kernels
    bool profiling;
    std::vector<PsSpan> pending;
    std::vector<cudaEvent_t> spare;
    double stage_ms[PS_N_STAGES];
    int64_t stage_calls[PS_N_STAGES];
};

namespace {
constexpr int MAIL_WORDS = 4, MAIL_CHUNK = 64;

cudaError_t mail_take(ps_ctx *ctx, cudaStream_t s, PsMailbox *out)
{
    std::unique_lock<std::mutex> lock(ctx->mu);
    if (ctx->mail_free.empty()) {
        int64_t *h = nullptr, *d = nullptr;
        cudaError_t e = cudaHostAlloc((void **)&h, sizeof(int64_t) * MAIL_WORDS * MAIL_CHUNK, cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        e = cudaHostGetDevicePointer((void **)&d, (void *)h, 0);
        if (e != cudaSuccess) { cudaFreeHost(h); return e; }
        memset(h, 0, sizeof(int64_t) * MAIL_WORDS * MAIL_CHUNK);
        ctx->mail_chunks.push_back(h);
        for (int i = 0; i < MAIL_CHUNK; ++i) ctx->mail_free.push_back({ h + MAIL_WORDS * i, d + MAIL_WORDS * i, s });
    }
    int pick = (int)ctx->mail_free.size() - 1;
    for (int i = pick; i >= 0; --i)
        if (ctx->mail_free[i].stream == s) { pick = i; break; }
    PsMailbox mb = ctx->mail_free[pick];
    ctx->mail_free.erase(ctx->mail_free.begin() + pick);
    lock.unlock();
    if (mb.stream != s) cudaStreamSynchronize(mb.stream); // a write queued by its previous user cannot land in our call
    mb.stream = s;
    *out = mb;
    return cudaSuccess;
}

void mail_give(ps_ctx *ctx, const PsMailbox &mb)
{
    if (!mb.h) return;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->mail_free.push_back(mb);
}

// brackets one stage with CUDA events on the launching stream when profiling is on
struct StageTimer {
    ps_ctx *ctx; cudaStream_t s; PsSpan span; bool on;
    StageTimer(ps_ctx *c, int stage, cudaStream_t st) : ctx(c), s(st), on(c->profiling)
    {
        if (!on) return;
        span.stage = stage;
        {
            std::lock_guard<std::mutex> lock(ctx->mu);
            for (cudaEvent_t *e : { &span.a, &span.b }) {
                if (!ctx->spare.empty()) { *e = ctx->spare.back(); ctx->spare.pop_back(); }
                else if (cudaEventCreate(e) != cudaSuccess) { on = false; return; }
            }
        }
        cudaEventRecord(span.a, s);
    }
    ~StageTimer()
    {
        if (!on) return;
        cudaEventRecord(span.b, s);
        std::lock_guard<std::mutex> lock(ctx->mu);
        ctx->pending.push_back(span);
    }
};
} // namespace

struct ps_saved {
    PsGeometry g;
    PsTable t;
    PsLists l;
    ps_ctx *ctx;
    cudaStream_t stream;  // the forward's stream
    cudaStream_t last_stream; // the stream of the last call that touched the saved buffers (forward, backward, tap)
    PsMailbox mail;
    bool resolved;        // M / n_work below are known on the host (sync-free small calls resolve them lazily)
    bool stats;           // the forward ran with PS_FLAG_RASTER_STATS: the backward counts its pairs too
    int64_t M;
    int n_work;           // non-empty (view, tile) lists
    int64_t cap_M;        // capacity of the list arrays (= M, or the worst case V * N * n_tiles of a sync-free call)
    int cap_work;         // grid bound for the per-list kernels (= n_work, or V * n_tiles)
    uint64_t *keys; // sorted int64 keys, materialised only with PS_FLAG_KEEP_BINNING
    int32_t *last;  // [V,H,W] tile-list position + 1 of the last contributor (PS_FLAG_KEEP_BINNING: tap only)
    int32_t *blast; // [V,H,W] block-list index + 1 of the last contributor (what the backward starts from)
    int32_t *frame_off, *frame_views; // CSR: the views of every frame (projection backward sums them per row)
    float *bg;      // [3] background colour of the forward (lives in the offsets allocation)
    float *t_pen;   // [V,H,W]
};

namespace {
// A later call on another stream (a backward issued from a different torch stream, a release from whichever thread
// drops the last reference) is ordered behind the work already queued on the saved buffers: without this the arena
// could hand a released block to that stream while the previous stream's kernels still use it.
int saved_order_after(ps_saved *sv, cudaStream_t s)
{
    if (sv->last_stream == s) return 0;
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return fail(2, "cudaEventCreate failed (cross-stream use of a saved forward)");
    cudaError_t e = cudaEventRecord(ev, sv->last_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, ev, 0);
    cudaEventDestroy(ev); // released by the runtime once the wait has been satisfied
    if (e != cudaSuccess) return fail(2, "cross-stream ordering of a saved forward failed: %s", cudaGetErrorString(e));
    sv->last_stream = s;
    return 0;
}

// M and the list count of a sync-free forward, read once its stream has drained
int saved_resolve(ps_saved *sv)
{
    if (sv->resolved) return 0;
    if (cudaStreamSynchronize(sv->stream) != cudaSuccess) return fail(2, "cudaStreamSynchronize failed while reading the forward's mailbox");
    sv->M = sv->mail.h[0];
    sv->n_work = (int)sv->mail.h[1];
    sv->resolved = true;
    return 0;
}
} // namespace

extern "C" {

int ps_abi_version(void) { return PS_ABI_VERSION; }
const char *ps_last_error(void) { return g_err; }

int ps_ctx_create(int device, ps_ctx **out)
{
    if (!out) return fail(1, "ps_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(2, "ps_ctx_create: no CUDA device (%s); libpsplat has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(1, "ps_ctx_create: device %d out of range (%d devices)", device, n);
    PS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(2, "ps_ctx_create: device %d is sm_%d%d; libpsplat is built for sm_100a only", device, prop.major, prop.minor);
    ps_ctx *c = new (std::nothrow) ps_ctx();
    if (!c) return fail(4, "ps_ctx_create: out of host memory");
    c->device = device;
    c->launches = 0;
    c->profiling = false;
    c->side = nullptr;
    for (int i = 0; i < PS_N_STAGES; ++i) { c->stage_ms[i] = 0.0; c->stage_calls[i] = 0; }
    PS_CUDA(cudaMalloc((void **)&c->d_stats, 8 * sizeof(unsigned long long)));
    PS_CUDA(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
    *out = c;
    return 0;
}

int ps_ctx_destroy(ps_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void *h : ctx->mail_chunks) cudaFreeHost(h);
    cudaFree(ctx->d_stats);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    for (auto &sp : ctx->pending) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : ctx->spare) cudaEventDestroy(e);
    if (ctx->device >= 0 && ctx->device < PS_MAX_DEVICES) {
        std::lock_guard<std::mutex> lock(g_arena[ctx->device].mu);
        arena_release_cached(g_arena[ctx->device]);
    }
    delete ctx;
    return 0;
}

int64_t ps_ctx_launch_count(const ps_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

static void saved_free(ps_saved *sv, cudaStream_t s)
{
    dev_free(sv->t.rec, s);
    dev_free(sv->t.tile_rect, s); dev_free(sv->t.tiles_touched, s); dev_free(sv->t.order, s); dev_free(sv->t.rank, s);
    dev_free(sv->t.depth, s);
    dev_free(sv->l.offsets, s); dev_free(sv->l.fill, s); dev_free(sv->l.worklist, s); dev_free(sv->l.cls, s);
    dev_free(sv->l.slots, s); dev_free(sv->l.vals, s); dev_free(sv->l.blist, s); dev_free(sv->l.bpos, s); dev_free(sv->l.cmask, s); dev_free(sv->l.cids, s); dev_free(sv->l.ccount, s); dev_free(sv->l.bcount, s);
    dev_free(sv->keys, s); dev_free(sv->last, s); dev_free(sv->blast, s); dev_free(sv->t_pen, s);
    sv->frame_off = sv->frame_views = nullptr; // live inside the offsets allocation
    sv->bg = nullptr;
    sv->l.n_lists = nullptr;                   // lives inside cls
    if (sv->ctx) mail_give(sv->ctx, sv->mail);
    sv->mail.h = nullptr;
}

static int forward_impl(ps_ctx *ctx, const ps_render_desc *d, const float *params, const int32_t *view_frame,
                        const float *viewmats, const float *Ks, const float *background, float *rgb, float *alpha,
                        uint32_t *rgba8, int32_t *n_contrib, ps_saved **saved, void *stream)
{
    if (saved) *saved = nullptr;
    if (!ctx || !d) return fail(1, "ps_forward: NULL context or descriptor");
    if (d->mode != PS_MODE_2D && d->mode != PS_MODE_3D) return fail(1, "ps_forward: unknown mode %d", d->mode);
    if (d->width <= 0 || d->height <= 0 || d->width > 32767 || d->height > 32767)
        return fail(1, "ps_forward: image size %dx%d unsupported", d->width, d->height);
    if (d->n_views < 0 || d->n_gauss < 0 || d->n_frames < 0) return fail(1, "ps_forward: negative size");
    if (d->n_views > 65535) return fail(1, "ps_forward: at most 65535 views per call (got %d)", d->n_views);
    if (d->n_views > 0 && (!view_frame || !background || (!rgba8 && (!rgb || !alpha)))) return fail(1, "ps_forward: NULL buffer");
    if (d->n_views > 0 && d->n_gauss > 0 && !params) return fail(1, "ps_forward: NULL params");
    if (d->mode == PS_MODE_3D && d->n_views > 0 && (!viewmats || !Ks)) return fail(1, "ps_forward: 3D needs viewmats and Ks");
    if ((int64_t)d->n_views * d->n_gauss > 0x7fffffffLL) return fail(1, "ps_forward: V*N exceeds 2^31");
    const bool save = (d->flags & PS_FLAG_SAVE_FOR_BACKWARD) != 0;
    const bool keep = (d->flags & PS_FLAG_KEEP_BINNING) != 0;
    if ((save || keep) && !saved) return fail(1, "ps_forward: saved is NULL but a SAVE/KEEP flag is set");
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));

    ps_saved *sv = new (std::nothrow) ps_saved();
    if (!sv) return fail(4, "ps_forward: out of host memory");
    memset(sv, 0, sizeof *sv);
    sv->ctx = ctx;
    sv->stream = s;
    sv->last_stream = s;
    sv->stats = (d->flags & PS_FLAG_RASTER_STATS) != 0;
    PsGeometry &g = sv->g;
    g.mode = d->mode; g.W = d->width; g.H = d->height; g.F = d->n_frames; g.N = d->n_gauss; g.V = d->n_views;
    g.tiles_x = (g.W + PS_TILE - 1) / PS_TILE; g.tiles_y = (g.H + PS_TILE - 1) / PS_TILE;
    g.n_tiles = g.tiles_x * g.tiles_y;
    g.tile_bits = ps_tile_bits(g.n_tiles);
    g.view_bits = 0;
    while ((1 << g.view_bits) < g.V) ++g.view_bits;
    g.near_plane = d->near_plane; g.far_plane = d->far_plane; g.radius_clip = d->radius_clip; g.eps2d = d->eps2d;
    g.activated = (d->flags & PS_FLAG_ACTIVATED_INPUTS) != 0 && d->mode == PS_MODE_3D;

    int rc = 0;
    cudaEvent_t ev_scan = nullptr, ev_fill = nullptr;
    bool filled = false; // the background fill already runs on the side stream
    static const bool no_fork = getenv("PS_NO_FILL_FORK") != nullptr; // A/B switch: everything on the caller's stream
    const bool fork_fill = !no_fork && !keep && (size_t)d->n_views * d->height * d->width > 0;
    static const bool late_fork = getenv("PS_FILL_FORK_LATE") != nullptr; // A/B switch: the fill beside the rasterizer instead of the binning
    uint32_t *rank_scratch = nullptr;
    int32_t *rank_flags = nullptr;
    long long *scan_scratch = nullptr;
    int32_t *csr_cursor = nullptr;
    uint8_t *mask_bytes = nullptr;
    // everything below jumps to `out` on error so scratch is always returned to the pool
#define PS_TRY_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { rc = fail(2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); goto out; } } while (0)
#define PS_TRY_LAUNCH(call) do { int n_ = (call); if (n_ < 0) { rc = fail(3, "kernel launch failed in %s: %s", #call, cudaGetErrorString(cudaGetLastError())); goto out; } ctx->launches += n_; } while (0)
    {
        const size_t VN = (size_t)g.V * g.N;
        const size_t T = (size_t)g.V * g.n_tiles;
        const size_t npix = (size_t)g.V * g.H * g.W;
        if (T > 0x7ffffff0ULL) { rc = fail(1, "ps_forward: %zu (view, tile) lists exceed 2^31", T); goto out; }
        if (g.N > (1 << 20)) { rc = fail(1, "ps_forward: at most 2^20 Gaussians per frame (got %d)", g.N); goto out; }
        // How the eight block lists of a tile are built (A/B switch PS_BIN_MODE for measurements):
        //   "bytes"  partition computes block-rectangle masks, the list sort carries them along as bytes, the split streams
        //   "split"  partition computes masks, one kernel sorts and splits from shared-memory bitmaps
        //   "gather" plain keys; the split gathers every record and tests it exactly
        // Measured at c2 / c3 (partition + sort + split + both rasterizers, ms): gather 10.69 / 17.29, bytes 10.80 / 17.64,
        // split 10.94 / 18.72 -- the mask variants build the lists faster (2.22 vs 2.44 ms at c2) but their cheap masks are a
        // superset of the exact test and the extra block-list entries cost the rasterizers more than that; exact masks in the
        // partition kernel cost 1.1 ms there.
        static const char *bin_env = getenv("PS_BIN_MODE");
        static const int mask_env = getenv("PS_EXACT_BLOCK_MASKS") ? 2 : 1;
        int bin_mode = 0; // 0 gather (default: fastest end to end, measured), 1 bytes, 2 split
        if (bin_env) bin_mode = !strcmp(bin_env, "gather") ? 0 : !strcmp(bin_env, "split") ? 2 : 1;
        if (bin_mode == 2 && !ps_split_fits_smem(g)) bin_mode = 1;
        if (bin_mode == 1 && !ps_mask_bytes_fit_smem(g)) bin_mode = 0;
        const bool split = bin_mode == 2;
        // Small calls (the reference's own call shape: one view per render()) never wait for the host: the list arrays
        // are sized for the worst case M = V * N * n_tiles and the per-list kernels are launched over all V * n_tiles
        // lists, reading the exact counts on the device.  Large batches size everything exactly from the mailbox
        // (one stream synchronisation): the worst case would not fit, and empty CTAs would cost more than the wait.
        static const bool force_sync = getenv("PS_FORCE_SYNC") != nullptr; // A/B switch for measurements
        const size_t worst = VN * (size_t)g.n_tiles;
        const bool sync_free = !keep && !force_sync && VN > 0 && T <= 16384 && worst <= ((size_t)1 << 31) / 40;
        sv->M = 0;
        sv->n_work = 0;
        sv->resolved = true;
        // offsets [T+1] | frame_off [F+1] | frame_views [V] | csr cursor [F] share one allocation: tiny pool
        // allocations split the large free blocks the next call wants to reuse
        PS_TRY_CUDA(dev_alloc(&sv->l.offsets, T + 1 + (size_t)g.F + 1 + (size_t)g.V + (size_t)g.F + 4, s));
        sv->frame_off = sv->l.offsets + T + 1;
        sv->frame_views = sv->frame_off + g.F + 1;
        csr_cursor = sv->frame_views + g.V;
        sv->bg = reinterpret_cast<float *>(csr_cursor + g.F); // the forward's background colour, kept for the backward
        PS_TRY_CUDA(dev_alloc(&sv->l.cls, (size_t)PS_CLS_WORDS, s));
        sv->l.n_lists = sv->l.cls + 3 * PS_N_CLASSES;
        PS_TRY_CUDA(dev_alloc(&scan_scratch, ps_scan_scratch_elems(g), s));
        PS_TRY_CUDA(cudaMemsetAsync(sv->l.offsets, 0, (T + 1) * sizeof(int32_t), s));
        if (VN > 0) {
            PS_TRY_CUDA(mail_take(ctx, s, &sv->mail));
            PS_TRY_CUDA(dev_alloc(&sv->t.rec, 4 * VN, s));
            PS_TRY_CUDA(dev_alloc(&sv->t.tile_rect, VN, s));
            PS_TRY_CUDA(dev_alloc(&sv->t.tiles_touched, VN, s));
            if (g.mode == PS_MODE_3D) PS_TRY_CUDA(dev_alloc(&sv->t.depth, VN, s));
            // the views of every frame (CSR): the 3D projection and the projection backward loop over them per Gaussian
            PS_TRY_LAUNCH(ps_launch_frame_csr(g, view_frame, sv->frame_off, csr_cursor, sv->frame_views, background, sv->bg, s));
            { StageTimer tm(ctx, PS_STAGE_PROJECT, s); PS_TRY_LAUNCH(ps_launch_project(g, params, view_frame, viewmats, Ks, sv->t, sv->l.offsets, sv->frame_off, sv->frame_views, s)); }
            // The list lengths are scanned first: their total M (which sizes the lists) reaches the host while the GPU is
            // still ranking depths, and the background fill of the empty tiles -- which needs nothing but the scanned
            // offsets and only writes memory -- runs on a side stream beside the latency-bound binning kernels.
            { StageTimer tm(ctx, PS_STAGE_SCAN, s); PS_TRY_LAUNCH(ps_launch_scan_lists(g, sv->l, scan_scratch, sv->mail.d, s)); }
            if (!sync_free) {
                PS_TRY_CUDA(cudaEventCreateWithFlags(&ev_scan, cudaEventDisableTiming));
                PS_TRY_CUDA(cudaEventRecord(ev_scan, s));
                if (fork_fill && !late_fork) {
                    {
                        std::lock_guard<std::mutex> lock(ctx->mu);
                        if (!ctx->side && cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) ctx->side = nullptr;
                    }
                    if (ctx->side) {
                        PS_TRY_CUDA(cudaEventCreateWithFlags(&ev_fill, cudaEventDisableTiming));
                        PS_TRY_CUDA(cudaStreamWaitEvent(ctx->side, ev_scan, 0));
                        PS_TRY_LAUNCH(ps_launch_fill_empty(g, sv->l.offsets, background, rgb, alpha, n_contrib, nullptr, rgba8, ctx->side));
                        PS_TRY_CUDA(cudaEventRecord(ev_fill, ctx->side));
                        filled = true;
                    }
                }
            }
            if (g.mode == PS_MODE_3D) {
                PS_TRY_CUDA(dev_alloc(&sv->t.order, VN, s));
                PS_TRY_CUDA(dev_alloc(&sv->t.rank, VN, s));
                const size_t ns = ps_rank_scratch_elems(g);
                if (ns) PS_TRY_CUDA(dev_alloc(&rank_scratch, ns, s));
                PS_TRY_CUDA(dev_alloc(&rank_flags, (size_t)g.V, s));
                StageTimer tm(ctx, PS_STAGE_RANK, s);
                PS_TRY_LAUNCH(ps_launch_depth_rank(g, sv->t, rank_scratch, rank_flags, s));
            }
            if (sync_free) {
                sv->resolved = false;
                sv->cap_M = (int64_t)worst;
                sv->cap_work = (int)T;
            } else {
                PS_TRY_CUDA(cudaEventSynchronize(ev_scan)); // the one host wait of a large forward: M sizes the lists
                sv->M = sv->mail.h[0];
                sv->n_work = (int)sv->mail.h[1];
                if (sv->M > 0x7fffffffLL) { rc = fail(1, "ps_forward: %lld tile intersections exceed 2^31", (long long)sv->M); goto out; }
                sv->cap_M = sv->M;
                sv->cap_work = sv->n_work;
            }
        }
        const size_t capM = (size_t)sv->cap_M, capW = (size_t)sv->cap_work;
        if (capM > 0) {
            PS_TRY_CUDA(dev_alloc(&sv->l.fill, T, s));
            PS_TRY_CUDA(dev_alloc(&sv->l.worklist, capW, s));
            PS_TRY_CUDA(dev_alloc(&sv->l.slots, capM, s));
            if (keep || !split) PS_TRY_CUDA(dev_alloc(&sv->l.vals, capM, s)); // the tile list: taps / the gather fallback
            PS_TRY_CUDA(dev_alloc(&sv->l.blist, 8 * capM, s));
            PS_TRY_CUDA(dev_alloc(&sv->l.bcount, 8 * capW, s));
            if (keep) PS_TRY_CUDA(dev_alloc(&sv->l.bpos, 8 * capM, s));
            if (save) { // contributor lists, written by the forward rasterizer
                PS_TRY_CUDA(dev_alloc(&sv->l.cids, 8 * capM, s));
                PS_TRY_CUDA(dev_alloc(&sv->l.cmask, 8 * capM, s));
                PS_TRY_CUDA(dev_alloc(&sv->l.ccount, 8 * capW, s));
            }
            PS_TRY_CUDA(cudaMemsetAsync(sv->l.fill, 0, T * sizeof(int32_t), s));
            { StageTimer tm(ctx, PS_STAGE_PARTITION, s); PS_TRY_LAUNCH(ps_launch_partition(g, sv->t, sv->l, bin_mode ? mask_env : 0, s)); }
            if (split) {
                StageTimer tm(ctx, PS_STAGE_SORT, s);
                PS_TRY_LAUNCH(ps_launch_build_worklist(g, sv->l, s));
                PS_TRY_LAUNCH(ps_launch_sort_split(g, sv->t, sv->l, sv->cap_work, s));
            } else {
                if (bin_mode == 1) PS_TRY_CUDA(dev_alloc(&mask_bytes, capM, s));
                { StageTimer tm(ctx, PS_STAGE_SORT, s);
                  PS_TRY_LAUNCH(ps_launch_build_worklist(g, sv->l, s));
                  PS_TRY_LAUNCH(ps_launch_sort_lists(g, sv->t, sv->l, sv->cap_work, mask_bytes, s)); }
                { StageTimer tm(ctx, PS_STAGE_BLOCKS, s); PS_TRY_LAUNCH(ps_launch_block_lists(g, sv->t, sv->l, sv->cap_work, mask_bytes, s)); }
            }
            if (keep) {
                PS_TRY_CUDA(dev_alloc(&sv->keys, capM, s));
                PS_TRY_LAUNCH(ps_launch_debug_keys(g, sv->t, sv->l, sv->n_work, sv->keys, s));
            }
        }
        if (save && VN == 0) PS_TRY_LAUNCH(ps_launch_frame_csr(g, view_frame, sv->frame_off, csr_cursor, sv->frame_views, background, sv->bg, s));
        if (npix > 0) {
            if (save) {
                PS_TRY_CUDA(dev_alloc(&sv->blast, npix, s));
                PS_TRY_CUDA(dev_alloc(&sv->t_pen, npix, s));
            }
            if (keep) PS_TRY_CUDA(dev_alloc(&sv->last, npix, s));
            StageTimer tm(ctx, PS_STAGE_RASTER_FWD, s);
            if (!filled && fork_fill && late_fork && ev_scan) {
                {
                    std::lock_guard<std::mutex> lock(ctx->mu);
                    if (!ctx->side && cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) ctx->side = nullptr;
                }
                if (ctx->side) {
                    PS_TRY_CUDA(cudaEventCreateWithFlags(&ev_fill, cudaEventDisableTiming));
                    PS_TRY_CUDA(cudaEventRecord(ev_scan, s)); // re-used: everything the caller's stream holds so far
                    PS_TRY_CUDA(cudaStreamWaitEvent(ctx->side, ev_scan, 0));
                    PS_TRY_LAUNCH(ps_launch_fill_empty(g, sv->l.offsets, background, rgb, alpha, n_contrib, nullptr, rgba8, ctx->side));
                    PS_TRY_CUDA(cudaEventRecord(ev_fill, ctx->side));
                    filled = true;
                }
            }
            if (!filled) PS_TRY_LAUNCH(ps_launch_fill_empty(g, sv->l.offsets, background, rgb, alpha, n_contrib, sv->last, rgba8, s));
            PS_TRY_LAUNCH(ps_launch_raster_fwd(g, sv->t, sv->l, sv->cap_work, background, rgb, alpha, n_contrib,
                                               sv->last, sv->blast, sv->t_pen, rgba8, sv->stats ? ctx->d_stats : nullptr, s));
        }
    }
out:
    if (ev_fill) { // join: the caller's stream completes only after the side stream's fill
        if (filled && cudaStreamWaitEvent(s, ev_fill, 0) != cudaSuccess && rc == 0) rc = fail(2, "ps_forward: cudaStreamWaitEvent failed");
        cudaEventDestroy(ev_fill);
    }
    if (ev_scan) cudaEventDestroy(ev_scan);
    dev_free(rank_scratch, s);
    dev_free(rank_flags, s);
    dev_free(scan_scratch, s);
    dev_free(mask_bytes, s);
    dev_free(sv->l.fill, s); dev_free(sv->l.slots, s);
    dev_free(sv->t.order, s); dev_free(sv->t.rank, s);
    if (rc == 0 && !keep) { dev_free(sv->t.tile_rect, s); dev_free(sv->t.depth, s); dev_free(sv->l.vals, s); }
    if (rc != 0 || !(save || keep)) {
        saved_free(sv, s);
        delete sv;
        sv = nullptr;
    }
    if (saved) *saved = sv;
    return rc;
#undef PS_TRY_CUDA
#undef PS_TRY_LAUNCH
}

int ps_forward(ps_ctx *ctx, const ps_render_desc *d, const float *params, const int32_t *view_frame,
               const float *viewmats, const float *Ks, const float *background, float *rgb, float *alpha,
               int32_t *n_contrib, ps_saved **saved, void *stream)
{
    return forward_impl(ctx, d, params, view_frame, viewmats, Ks, background, rgb, alpha, nullptr, n_contrib, saved, stream);
}

int ps_forward_rgba8(ps_ctx *ctx, const ps_render_desc *d, const float *params, const int32_t *view_frame,
                     const float *viewmats, const float *Ks, const float *background, uint8_t *rgba8, void *stream)
{
    if (!rgba8 && d && d->n_views > 0) return fail(1, "ps_forward_rgba8: NULL output");
    if (((uintptr_t)rgba8 & 3u) != 0) return fail(1, "ps_forward_rgba8: output must be 4-byte aligned");
    if (d && (d->flags & (PS_FLAG_SAVE_FOR_BACKWARD | PS_FLAG_KEEP_BINNING))) return fail(1, "ps_forward_rgba8 is inference only");
    return forward_impl(ctx, d, params, view_frame, viewmats, Ks, background, nullptr, nullptr, reinterpret_cast<uint32_t *>(rgba8),
                        nullptr, nullptr, stream);
}

static int backward_impl(ps_ctx *ctx, ps_saved *sv, const float *params, const float *viewmats, const float *Ks,
                         const float *background, const float *d_rgb, const float *d_alpha, float *d_params,
                         float *const *peers, int my_rank, int world, void *stream)
{
    if (!ctx || !sv) return fail(1, "ps_backward: NULL context or saved state");
    const PsGeometry &g = sv->g;
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    const int P = g.mode == PS_MODE_3D ? 14 : 9;
    const size_t n_out = (size_t)g.F * g.N * P;
    if (n_out == 0) return 0;
    if (saved_order_after(sv, s)) return 2;
    if (!peers && !d_params) return fail(1, "ps_backward: NULL d_params");
    if (peers && (world < 1 || my_rank < 0 || my_rank >= world)) return fail(1, "ps_backward_peer: rank %d of %d", my_rank, world);
    const size_t VN = (size_t)g.V * g.N;
    // nothing was rendered (known on the host): the gradient is zero.  In peer mode the zeros are still PUSHED: every
    // staging slot is rewritten every step, so the owners never add a stale slot.
    const bool nothing = VN == 0 || (sv->resolved && sv->M == 0) || (size_t)g.H * g.W == 0;
    if (nothing && !peers) {
        PS_CUDA(cudaMemsetAsync(d_params, 0, n_out * sizeof(float), s));
        return 0;
    }
    if (!nothing && (!sv->blast || !sv->t_pen)) return fail(1, "ps_backward: forward was not run with PS_FLAG_SAVE_FOR_BACKWARD");
    if (!sv->frame_off) return fail(1, "ps_backward: forward was not run with PS_FLAG_SAVE_FOR_BACKWARD");
    if (!params || (!nothing && (!d_rgb || !d_alpha))) return fail(1, "ps_backward: NULL buffer");
    background = sv->bg; // the colour the forward composited against (the caller's buffer may have changed since)
    float *acc = nullptr;
    const size_t acc_rows = VN > 0 ? VN : 1;
    PS_CUDA(dev_alloc(&acc, acc_rows * PS_ACC_STRIDE + 4, s)); // + the rasterizer's task counter, zeroed with the rows
    int rc = 0;
    do {
        if (cudaMemsetAsync(acc, 0, (acc_rows * PS_ACC_STRIDE + 4) * sizeof(float), s) != cudaSuccess) { rc = fail(2, "ps_backward: memset failed"); break; }
        int n = 0;
        if (!nothing) {
            StageTimer tm(ctx, PS_STAGE_RASTER_BWD, s);
            n = ps_launch_raster_bwd(g, sv->t, sv->l, sv->cap_work, background, sv->blast, sv->t_pen, d_rgb, d_alpha, acc,
                                     reinterpret_cast<unsigned *>(acc + acc_rows * PS_ACC_STRIDE), sv->stats ? ctx->d_stats : nullptr, s);
        }
        if (n < 0) { rc = fail(3, "ps_backward: raster_bwd launch failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        ctx->launches += n;
        { StageTimer tm(ctx, PS_STAGE_PROJECT_BWD, s); n = ps_launch_project_bwd(g, params, sv->frame_off, sv->frame_views, viewmats, Ks, sv->t, acc, d_params, peers, my_rank, world, s); }
        if (n < 0) { rc = fail(3, "ps_backward: project_bwd launch failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        ctx->launches += n;
    } while (0);
    dev_free(acc, s);
    return rc;
}

int ps_backward(ps_ctx *ctx, ps_saved *sv, const float *params, const int32_t *view_frame, const float *viewmats,
                const float *Ks, const float *background, const float *d_rgb, const float *d_alpha, float *d_params,
                void *stream)
{
    (void)view_frame; // the forward saved the views of every frame
    return backward_impl(ctx, sv, params, viewmats, Ks, background, d_rgb, d_alpha, d_params, nullptr, 0, 1, stream);
}

int ps_backward_peer(ps_ctx *ctx, ps_saved *sv, const float *params, const float *viewmats, const float *Ks,
                     const float *background, const float *d_rgb, const float *d_alpha, float *const *stage_ranks,
                     int my_rank, int world, void *stream)
{
    if (!stage_ranks) return fail(1, "ps_backward_peer: NULL stage_ranks");
    return backward_impl(ctx, sv, params, viewmats, Ks, background, d_rgb, d_alpha, nullptr, stage_ranks, my_rank, world, stream);
}

int ps_peer_sum(ps_ctx *ctx, const float *stage_local, int world, size_t n_per_slot, float *out, void *stream)
{
    if (!ctx || !stage_local || !out || world < 1) return fail(1, "ps_peer_sum: bad argument");
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_LAUNCH(ctx, ps_launch_peer_sum(stage_local, world, n_per_slot, out, (cudaStream_t)stream));
    return 0;
}

int ps_saved_info_get(const ps_saved *sv, ps_saved_info *out)
{
    if (!sv || !out) return fail(1, "ps_saved_info_get: NULL argument");
    if (int rc = saved_resolve(const_cast<ps_saved *>(sv))) return rc;
    memset(out, 0, sizeof *out);
    out->n_isect = sv->M;
    out->tile_bits = sv->g.tile_bits; out->view_bits = sv->g.view_bits;
    out->tiles_x = sv->g.tiles_x; out->tiles_y = sv->g.tiles_y;
    out->n_views = sv->g.V; out->n_gauss = sv->g.N; out->n_frames = sv->g.F;
    out->mode = sv->g.mode; out->width = sv->g.W; out->height = sv->g.H;
    out->n_lists = sv->n_work;
    return 0;
}

int ps_saved_copy(ps_ctx *ctx, const ps_saved *sv, int what, void *dst, size_t bytes, void *stream)
{
    if (!ctx || !sv || !dst) return fail(1, "ps_saved_copy: NULL argument");
    if (int rc = saved_resolve(const_cast<ps_saved *>(sv))) return rc;
    PS_CUDA(cudaSetDevice(ctx->device));
    if (saved_order_after(const_cast<ps_saved *>(sv), (cudaStream_t)stream)) return 2;
    const PsGeometry &g = sv->g;
    const size_t VN = (size_t)g.V * g.N, npix = (size_t)g.V * g.H * g.W;
    const void *src = nullptr;
    size_t have = 0;
    switch (what) {
        case PS_TAP_ISECT_KEYS: src = sv->keys; have = sizeof(uint64_t) * (size_t)sv->M; break;
        case PS_TAP_FLATTEN_IDS: src = sv->l.vals; have = sizeof(uint32_t) * (size_t)sv->M; break;
        case PS_TAP_TILE_OFFSETS: src = sv->l.offsets; have = sizeof(int32_t) * ((size_t)g.V * g.n_tiles + 1); break;
        case PS_TAP_LAST_IDS: src = sv->last; have = sizeof(int32_t) * npix; break;
        case PS_TAP_TILES_TOUCHED: src = sv->t.tiles_touched; have = sizeof(int32_t) * VN; break;
        case PS_TAP_REC0: case PS_TAP_REC1: case PS_TAP_REC2: { // one 16-byte word out of every 64-byte record
            if (VN == 0) return 0;
            if (!sv->t.rec) return fail(1, "ps_saved_copy: records were not kept");
            if (bytes < sizeof(float4) * VN) return fail(1, "ps_saved_copy: destination holds %zu bytes, tap %d needs %zu", bytes, what, sizeof(float4) * VN);
            cudaStream_t s2 = (cudaStream_t)stream;
            PS_CUDA(cudaSetDevice(ctx->device));
            PS_CUDA(cudaMemcpy2DAsync(dst, sizeof(float4), sv->t.rec + (what - PS_TAP_REC0), 4 * sizeof(float4), sizeof(float4), VN,
                                      cudaMemcpyDefault, s2));
            PS_CUDA(cudaStreamSynchronize(s2));
            return 0;
        }
        case PS_TAP_DEPTH: src = sv->t.depth; have = sv->g.mode == PS_MODE_3D ? sizeof(uint32_t) * VN : 0; break;
        default: return fail(1, "ps_saved_copy: unknown tap %d", what);
    }
    if (have == 0) return 0;
    if (!src) return fail(1, "ps_saved_copy: tap %d was not kept (PS_FLAG_KEEP_BINNING / SAVE_FOR_BACKWARD)", what);
    if (bytes < have) return fail(1, "ps_saved_copy: destination holds %zu bytes, tap %d needs %zu", bytes, what, have);
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_CUDA(cudaMemcpyAsync(dst, src, have, cudaMemcpyDefault, s));
    PS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int ps_saved_release(ps_ctx *ctx, ps_saved *sv, void *stream)
{
    if (!sv) return 0;
    if (ctx) cudaSetDevice(ctx->device);
    if (saved_order_after(sv, (cudaStream_t)stream)) cudaStreamSynchronize(sv->last_stream); // cannot order: drain instead
    saved_free(sv, (cudaStream_t)stream);
    delete sv;
    return 0;
}

int ps_ctx_set_profiling(ps_ctx *ctx, int on)
{
    if (!ctx) return fail(1, "ps_ctx_set_profiling: NULL context");
    ctx->profiling = on != 0;
    return 0;
}

int ps_ctx_stage_times(ps_ctx *ctx, double *ms, int64_t *calls, int reset)
{
    if (!ctx) return fail(1, "ps_ctx_stage_times: NULL context");
    PS_CUDA(cudaSetDevice(ctx->device));
    std::vector<PsSpan> done;
    { std::lock_guard<std::mutex> lock(ctx->mu); done.swap(ctx->pending); }
    for (auto &sp : done) {
        PS_CUDA(cudaEventSynchronize(sp.b));
        float t = 0.0f;
        PS_CUDA(cudaEventElapsedTime(&t, sp.a, sp.b));
        std::lock_guard<std::mutex> lock(ctx->mu);
        ctx->stage_ms[sp.stage] += t;
        ctx->stage_calls[sp.stage] += 1;
        ctx->spare.push_back(sp.a);
        ctx->spare.push_back(sp.b);
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    for (int i = 0; i < PS_N_STAGES; ++i) {
        if (ms) ms[i] = ctx->stage_ms[i];
        if (calls) calls[i] = ctx->stage_calls[i];
        if (reset) { ctx->stage_ms[i] = 0.0; ctx->stage_calls[i] = 0; }
    }
    return 0;
}

int ps_ctx_raster_stats(ps_ctx *ctx, uint64_t *pairs, int reset, void *stream)
{
    if (!ctx || !pairs) return fail(1, "ps_ctx_raster_stats: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_CUDA(cudaMemcpyAsync(pairs, ctx->d_stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (reset) PS_CUDA(cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(uint64_t), s));
    PS_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int ps_fp32_peak_probe(ps_ctx *ctx, double *tflops, void *stream)
{
    if (!ctx || !tflops) return fail(1, "ps_fp32_peak_probe: NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    float *sink = nullptr;
    PS_CUDA(cudaMalloc((void **)&sink, sizeof(float)));
    cudaEvent_t a, b;
    PS_CUDA(cudaEventCreate(&a));
    PS_CUDA(cudaEventCreate(&b));
    const int iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) { // first reps warm the clocks
        PS_CUDA(cudaEventRecord(a, s));
        PS_LAUNCH(ctx, ps_launch_fp32_probe(sink, iters, s));
        PS_CUDA(cudaEventRecord(b, s));
        PS_CUDA(cudaEventSynchronize(b));
        float ms = 0.0f;
        PS_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 64.0 * iters * 256.0 * (148.0 * 8.0);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    *tflops = best;
    return 0;
}

int ps_math_probe(ps_ctx *ctx, const float *x, int n, float *y, void *stream)
{
    if (!ctx) return fail(1, "ps_math_probe: NULL context");
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_LAUNCH(ctx, ps_launch_math_probe(x, n, y, (cudaStream_t)stream));
    return 0;
}

int ps_adapter3d_probe(ps_ctx *ctx, const float *rows, int n, const float *v_act, float *act, float *d_rows, void *stream)
{
    if (!ctx) return fail(1, "ps_adapter3d_probe: NULL context");
    if (n > 0 && (!rows || !act)) return fail(1, "ps_adapter3d_probe: NULL buffer");
    if ((v_act == nullptr) != (d_rows == nullptr)) return fail(1, "ps_adapter3d_probe: v_act and d_rows go together");
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_LAUNCH(ctx, ps_launch_adapter3d_probe(rows, n, v_act, act, d_rows, (cudaStream_t)stream));
    return 0;
}

int ps_view_loss(ps_ctx *ctx, int n_views, int height, int width, const float *rgb, const float *alpha,
                 const float *target_img, const float *target_mask, float ssim_lambda, float img_lambda,
                 float *losses, float *d_rgb, float *d_alpha, void *stream)
{
    if (!ctx) return fail(1, "ps_view_loss: NULL context");
    if (n_views < 0 || height < 0 || width < 0) return fail(1, "ps_view_loss: negative size");
    if (n_views == 0) return 0;
    if (height < 11 || width < 11) return fail(1, "ps_view_loss: the 11x11 SSIM window needs images of at least 11x11 pixels (got %dx%d)", width, height);
    if (!rgb || !alpha || !target_img || !target_mask || !losses) return fail(1, "ps_view_loss: NULL buffer");
    if ((d_rgb == nullptr) != (d_alpha == nullptr)) return fail(1, "ps_view_loss: d_rgb and d_alpha must be given together");
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    double *stats = nullptr;
    float *adj = nullptr;
    PS_CUDA(dev_alloc(&stats, (size_t)n_views * 8, s));
    cudaError_t e = dev_alloc(&adj, (size_t)n_views * 9 * (size_t)height * width, s);
    if (e != cudaSuccess) { dev_free(stats, s); return fail(2, "ps_view_loss: scratch allocation failed: %s", cudaGetErrorString(e)); }
    const int n = ps_launch_view_loss(n_views, height, width, rgb, alpha, target_img, target_mask, ssim_lambda, img_lambda,
                                      stats, adj, losses, d_rgb, d_alpha, s);
    dev_free(adj, s);
    dev_free(stats, s);
    if (n < 0) return fail(3, "ps_view_loss: kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches += n;
    return 0;
}

int ps_iou_loss(ps_ctx *ctx, int n_views, int height, int width, const float *alpha, const float *target_mask, float *losses,
                float *d_alpha, void *stream)
{
    if (!ctx) return fail(1, "ps_iou_loss: NULL context");
    if (n_views < 0 || height < 0 || width < 0) return fail(1, "ps_iou_loss: negative size");
    if (n_views == 0) return 0;
    if (!alpha || !target_mask || !losses) return fail(1, "ps_iou_loss: NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    PS_CUDA(cudaSetDevice(ctx->device));
    double *stats = nullptr;
    PS_CUDA(dev_alloc(&stats, (size_t)n_views * 8, s));
    const int n = ps_launch_iou_loss(n_views, height, width, alpha, target_mask, stats, losses, d_alpha, s);
    dev_free(stats, s);
    if (n < 0) return fail(3, "ps_iou_loss: kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches += n;
    return 0;
}

static int head_check(ps_ctx *ctx, int mode, int n, const void *a, const void *b, const char *what)
{
    if (!ctx) return fail(1, "%s: NULL context", what);
    if (mode != PS_MODE_2D && mode != PS_MODE_3D) return fail(1, "%s: unknown mode %d", what, mode);
    if (n < 0) return fail(1, "%s: negative row count", what);
    if (n > 0 && (!a || !b)) return fail(1, "%s: NULL buffer", what);
    return 0;
}

int ps_param_head_forward(ps_ctx *ctx, int mode, int n, const float *net_out, const float *probs_sel, const float *grid_sel,
                          const float *scale0, float voxel_size, float prob_threshold, float clip_lo, float clip_hi,
                          int pose, double angle, const float *p_3d_host, const float *poses, const int32_t *row_frame,
                          int n_frames, float *rows, void *stream)
{
    if (int rc = head_check(ctx, mode, n, net_out, probs_sel, "ps_param_head_forward")) return rc;
    if (n > 0 && (!rows || !scale0 || (mode == PS_MODE_3D && !grid_sel))) return fail(1, "ps_param_head_forward: NULL buffer");
    if (pose && mode == PS_MODE_3D && !p_3d_host && !poses) return fail(1, "ps_param_head_forward: pose requested without p_3d");
    if ((poses == nullptr) != (row_frame == nullptr)) return fail(1, "ps_param_head_forward: poses and row_frame go together");
    PS_CUDA(cudaSetDevice(ctx->device));
    PS_LAUNCH(ctx, ps_launch_head_fwd(mode, n, net_out, probs_sel, grid_sel, scale0, voxel_size, prob_threshold, clip_lo, clip_hi,
                                      pose, angle, p_3d_host, poses, row_frame, n_frames, rows, (cudaStream_t)stream));
    return 0;
}

int ps_param_head_backward(ps_ctx *ctx, int mode, int n, const float *net_out, const float *probs_sel, float voxel_size,
                           float prob_threshold, float clip_lo, float clip_hi, int pose, double angle, const float *poses,
                           const int32_t *row_frame, int n_frames, const float *d_rows, float *d_net_out, float *d_probs_sel, float *d_scale0,
                           void *stream)
{
    if (int rc = head_check(ctx, mode, n, net_out, probs_sel, "ps_param_head_backward")) return rc;
    if (!d_scale0 || (n > 0 && (!d_rows || !d_net_out || !d_probs_sel))) return fail(1, "ps_param_head_backward: NULL buffer");
    PS_CUDA(cudaSetDevice(ctx->device));
    if ((poses == nullptr) != (row_frame == nullptr)) return fail(1, "ps_param_head_backward: poses and row_frame go together");
    PS_LAUNCH(ctx, ps_launch_head_bwd(mode, n, net_out, probs_sel, voxel_size, prob_threshold, clip_lo, clip_hi, pose, angle,
                                      poses, row_frame, n_frames, d_rows, d_net_out, d_probs_sel, d_scale0, (cudaStream_t)stream));
    return 0;
}

} // extern "C"
