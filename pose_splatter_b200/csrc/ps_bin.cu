// ps_bin.cu -- tile binning (SURVEY.md 2.2 K2'-K4'): builds, for every (view, tile), the list of
// Gaussians sorted by (depth bits | row index), i.e. exactly the order gsplat's
// isect_tiles + cub::DeviceRadixSort(int64 key = view|tile|depth) + isect_offset_encode produce --
// without ever sorting 12-byte (key, value) pairs over M = sum(tiles touched):
//
//   depth_rank   (3D) one CTA per view: stable LSD radix sort of the view's N depth words in shared
//                memory (skipping the digits above the highest differing bit; every warp owns a contiguous
//                segment, so a pass needs three CTA barriers) -> order[], rank[]
//   scan_lists   exclusive scan of the per-(view,tile) counts the projection kernel histogrammed
//                -> tile ranges (offsets), M, list size classes            [this IS isect_offsets]
//   partition    every Gaussian drops its depth rank (3D) / row index (2D) into the lists of the
//                tiles it touches (CTA-aggregated reservations, order inside a list arbitrary)
//   sort_lists   per non-empty list: the ranks are unique integers < N, so a bitmap in shared
//                memory + popcount prefix sorts them; writes vals = view*N + gaussian
//   worklist     non-empty lists ordered longest class first (launch order of the rasterizers)
//
// Equal depth words keep Gaussian order (stable ranking) like the stable radix sort of the reference
// path.  HBM traffic: 4 B written + 4 B read + 4 B written per list entry instead of 12 B x 2 x 7 passes.
// The sorted int64 keys are implied by (list id, depth word of vals[i]); ps_launch_debug_keys
// materialises them for the bit-exact parity taps only.
#include "ps_contract.cuh"
#include "ps_cull.cuh"
#include "ps_internal.h"
#include <cstdlib>

namespace {

constexpr int RT = PS_RANK_THREADS;
constexpr int RW = RT / 32;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t PS_SLOT_KEY_MASK = (1u << PS_SLOT_MASK_SHIFT) - 1u;

// exclusive scan of s[0..255] in place, result total returned to every thread; all RT threads call
__device__ __forceinline__ uint32_t scan256_exclusive(uint32_t *s, uint32_t *s_tmp)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t v = 0, incl = 0;
    if (tid < 256) {
        v = s[tid];
        incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_tmp[wid] = incl;
    }
    __syncthreads();
    if (tid < 256) {
        uint32_t pre = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) pre += (w < wid) ? s_tmp[w] : 0u;
        s[tid] = pre + incl - v;
    }
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) total += s_tmp[w];
    __syncthreads();
    return total;
}

// One CTA per view.  IdT = uint16_t with the ping-pong buffers in dynamic shared memory,
// uint32_t with them in global scratch (N too large for shared memory).
template <typename IdT>
__global__ void __launch_bounds__(RT)
depth_rank_kernel(int N, const uint32_t *__restrict__ depth, const int32_t *__restrict__ touched,
                  uint32_t *__restrict__ order, uint32_t *__restrict__ rank, uint32_t *gscratch,
                  const int32_t *__restrict__ only_flagged)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ uint32_t s_hist[256], s_tmp[8], s_minmax[2];
    if (only_flagged && only_flagged[blockIdx.x] == 0) return; // this view was ranked by the bucket kernel
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)blockIdx.x * N;
    // per-(warp, digit) counts, then output cursors (< N: fits IdT); two matrices: this pass's and the next's
    IdT (*s_wcnt)[RW][256] = reinterpret_cast<IdT (*)[RW][256]>(dyn);
    unsigned char *dyn_keys = dyn + 2 * RW * 256 * sizeof(IdT);
    uint32_t *k[2];
    IdT *id[2];
    if (gscratch) {
        uint32_t *p = gscratch + (size_t)blockIdx.x * 4 * N;
        k[0] = p; k[1] = p + N;
        id[0] = reinterpret_cast<IdT *>(p + 2 * (size_t)N); id[1] = reinterpret_cast<IdT *>(p + 3 * (size_t)N);
    } else {
        k[0] = reinterpret_cast<uint32_t *>(dyn_keys); k[1] = k[0] + N;
        id[0] = reinterpret_cast<IdT *>(k[1] + N); id[1] = id[0] + N;
    }
    if (tid == 0) { s_minmax[0] = 0xffffffffu; s_minmax[1] = 0u; }
    __syncthreads();
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (int i = tid; i < N; i += RT) {
        const bool listed = touched[base + i] > 0;
        const uint32_t key = listed ? depth[base + i] : 0xffffffffu;
        k[0][i] = key;
        id[0][i] = (IdT)i;
        if (listed) { mn = min(mn, key); mx = max(mx, key); }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(FULL, mn, d));
        mx = max(mx, __shfl_xor_sync(FULL, mx, d));
    }
    if (lane == 0) { atomicMin(&s_minmax[0], mn); atomicMax(&s_minmax[1], mx); }
    __syncthreads();
    mn = s_minmax[0]; mx = s_minmax[1];
    // digits above the highest bit in which two listed depth words differ cannot change the order
    const uint32_t diff = (mx >= mn) ? (mn ^ mx) : 0u;
    const int npass = diff ? (32 - __clz(diff) + 7) / 8 : 0;
    // Every warp owns one contiguous segment of the array for the whole pass, so one scan of the per-(warp, digit)
    // counts in (digit, warp) order gives each warp a private output cursor for every digit: ties keep their input
    // order with a handful of CTA barriers per pass.  The counts of pass p+1 are taken while pass p scatters (the
    // destination segment of an element is known from its output position), so the keys are swept once per pass.
    const int seg = ((N + RW - 1) / RW + 31) & ~31;
    const float inv_seg = 1.0f / (float)seg;
    const int seg_lo = min(N, wid * seg), seg_hi = min(N, seg_lo + seg);
    auto cnt_inc = [&](IdT *row, uint32_t digit) { // ++row[digit]; 16-bit counters are bumped in pairs (no 16-bit atomics)
        if (sizeof(IdT) == 2) atomicAdd(reinterpret_cast<uint32_t *>(row) + (digit >> 1), 1u << (16u * (digit & 1u)));
        else atomicAdd(reinterpret_cast<uint32_t *>(row) + digit, 1u);
    };
    int cur = 0;
    if (npass > 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s_wcnt[0][wid][lane + 32 * q] = (IdT)0;
        __syncwarp();
        for (int i = seg_lo + lane; i < seg_hi; i += 32) cnt_inc(s_wcnt[0][wid], k[0][i] & 255u);
    }
    for (int p = 0; p < npass; ++p) {
        const int shift = 8 * p;
        const uint32_t *kin = k[p & 1];
        const IdT *iin = id[p & 1];
        uint32_t *kout = k[(p + 1) & 1];
        IdT *iout = id[(p + 1) & 1];
        const bool more = p + 1 < npass;
        __syncthreads(); // the counts of this pass are complete
        {   // exclusive scan of the 256 x 32 counts in (digit, warp) order: four threads per digit, eight warps
            // each; digit totals are scanned across the CTA
            const int d = tid >> 2, part = tid & 3;
            uint32_t loc[8], run = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t cc = s_wcnt[cur][part * 8 + i][d];
                loc[i] = run;
                run += cc;
                s_wcnt[cur ^ 1][part * 8 + i][d] = (IdT)0; // next pass's counters
            }
            uint32_t incl = run;
            uint32_t n = __shfl_up_sync(FULL, incl, 1, 4);
            if (part >= 1) incl += n;
            n = __shfl_up_sync(FULL, incl, 2, 4);
            if (part >= 2) incl += n;
            const uint32_t off = incl - run;
            if (part == 3) s_hist[d] = incl; // digit total
            __syncthreads();
            scan256_exclusive(s_hist, s_tmp);
            const uint32_t dbase = s_hist[d];
#pragma unroll
            for (int i = 0; i < 8; ++i) s_wcnt[cur][part * 8 + i][d] = (IdT)(dbase + loc[i] + off);
        }
        __syncthreads();
        for (int i0 = seg_lo; i0 < seg_hi; i0 += 32) {
            const int i = i0 + lane;
            const bool live = i < seg_hi;
            const uint32_t live_mask = __ballot_sync(FULL, live);
            if (live) {
                const uint32_t key = kin[i];
                const IdT idv = iin[i];
                const uint32_t digit = (key >> shift) & 255u;
                const uint32_t peers = __match_any_sync(live_mask, digit);
                const uint32_t r = __popc(peers & ((1u << lane) - 1u));
                const uint32_t pos = (uint32_t)s_wcnt[cur][wid][digit] + r;
                kout[pos] = key;
                iout[pos] = idv;
                if (more) cnt_inc(s_wcnt[cur ^ 1][min(RW - 1, (int)(((float)pos + 0.5f) * inv_seg))], (key >> (shift + 8)) & 255u);
                __syncwarp(live_mask);
                if (r == 0) s_wcnt[cur][wid][digit] += (IdT)__popc(peers);
            }
            __syncwarp();
        }
        cur ^= 1;
    }
    __syncthreads();
    const IdT *fin = id[npass & 1];
    for (int r = tid; r < N; r += RT) {
        const uint32_t gid = (uint32_t)fin[r];
        order[base + r] = gid;
        rank[base + gid] = (uint32_t)r;
    }
}

// Depth ranking by one bucket pass (the common case).  The depth words of a view lie in a narrow range, so a
// monotonic linear map onto NB >= N buckets leaves ~1 key per bucket: count (shared-memory atomics, which also give
// every key a slot inside its bucket), scan, place, and repair the order inside the few buckets that hold more than
// one key with an insertion sort on (depth word, Gaussian index) -- the order a stable sort of the depth words gives.
// Four sweeps over the keys and five CTA barriers instead of three radix passes of scans and match-any scatters.
// A view with a crowded bucket (> RANK_BUCKET_LIMIT equal or nearly equal depths) sets its flag and is ranked by
// the radix kernel afterwards.
constexpr int RANK_BUCKET_LIMIT = 32;

__global__ void __launch_bounds__(RT)
depth_rank_bucket_kernel(int N, int NB, const uint32_t *__restrict__ depth, const int32_t *__restrict__ touched,
                         uint32_t *__restrict__ order, uint32_t *__restrict__ rank, int32_t *__restrict__ fallback)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ uint32_t s_minmax[2], s_wsum[RW], s_nunlisted, s_bad;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(dyn);             // [NB + 1] counts, then exclusive bases
    uint32_t *keys = cnt + NB + 1;                                 // [N] depth words by Gaussian (0xffffffff = not listed)
    uint16_t *tgid = reinterpret_cast<uint16_t *>(keys + N);       // [N] Gaussians in bucket order
    uint16_t *slot = tgid + ((N + 1) & ~1);                        // [N] slot of every key inside its bucket
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)blockIdx.x * N;
    for (int i = tid; i <= NB; i += RT) cnt[i] = 0u;
    if (tid == 0) { s_minmax[0] = 0xffffffffu; s_minmax[1] = 0u; s_nunlisted = 0u; s_bad = 0u; }
    // the only sweep over global memory: independent loads, four elements in flight per thread
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (int i0 = tid; i0 < N; i0 += 4 * RT) {
        uint32_t k4[4];
        int t4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * RT;
            k4[u] = i < N ? __ldg(depth + base + i) : 0u;
            t4[u] = i < N ? __ldg(touched + base + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * RT;
            if (i >= N) continue;
            const bool listed = t4[u] > 0;
            keys[i] = listed ? k4[u] : 0xffffffffu;
            if (listed) { mn = min(mn, k4[u]); mx = max(mx, k4[u]); }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(FULL, mn, d));
        mx = max(mx, __shfl_xor_sync(FULL, mx, d));
    }
    __syncthreads(); // s_minmax and cnt are initialised
    if (lane == 0) { atomicMin(&s_minmax[0], mn); atomicMax(&s_minmax[1], mx); }
    __syncthreads();
    mn = s_minmax[0]; mx = s_minmax[1];
    // a listed depth word is never 0xffffffff (that is a NaN): the marker cannot collide with a key
    const unsigned long long range = (mx >= mn) ? (unsigned long long)(mx - mn) : 0ull;
    const unsigned long long scale = ((unsigned long long)NB << 32) / (range + 1ull); // bucket = (key - mn) * scale >> 32 < NB
    auto bucket_of = [&](uint32_t key) -> uint32_t { return (uint32_t)(((unsigned long long)(key - mn) * scale) >> 32); };
    for (int i = tid; i < N; i += RT) {
        const uint32_t key = keys[i];
        const uint32_t sl = key != 0xffffffffu ? atomicAdd(&cnt[bucket_of(key)], 1u) : atomicAdd(&s_nunlisted, 1u);
        slot[i] = (uint16_t)sl;
    }
    __syncthreads();
    {   // exclusive scan of the NB counts (NB is a multiple of RT): a contiguous run per thread
        const int per = NB / RT;
        uint32_t sum = 0, big = 0;
        for (int k = 0; k < per; ++k) { const uint32_t c = cnt[tid * per + k]; sum += c; big = max(big, c); }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) s_wsum[wid] = incl;
        if (big > (uint32_t)RANK_BUCKET_LIMIT) s_bad = 1u;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < RW; ++w) run += (w < wid) ? s_wsum[w] : 0u;
        for (int k = 0; k < per; ++k) { const uint32_t c = cnt[tid * per + k]; cnt[tid * per + k] = run; run += c; }
        if (tid == RT - 1) cnt[NB] = run; // number of listed Gaussians
    }
    __syncthreads();
    if (s_bad) { // crowded bucket: leave this view to the radix kernel
        if (tid == 0) fallback[blockIdx.x] = 1;
        return;
    }
    const uint32_t n_listed = cnt[NB];
    for (int i = tid; i < N; i += RT) {
        const uint32_t key = keys[i];
        const uint32_t pos = key != 0xffffffffu ? cnt[bucket_of(key)] + slot[i] : n_listed + slot[i];
        tgid[pos] = (uint16_t)i;
    }
    __syncthreads();
    {   // buckets holding several keys: order by (depth word, Gaussian index)
        const int per = NB / RT;
        for (int k = 0; k < per; ++k) {
            const uint32_t b0 = cnt[tid * per + k], b1 = cnt[tid * per + k + 1];
            for (uint32_t a = b0 + 1; a < b1; ++a) {
                const uint16_t ga = tgid[a];
                const uint32_t ka = keys[ga];
                uint32_t j = a;
                while (j > b0) {
                    const uint16_t gp = tgid[j - 1];
                    const uint32_t kp = keys[gp];
                    if (!(kp > ka || (kp == ka && gp > ga))) break;
                    tgid[j] = gp;
                    --j;
                }
                tgid[j] = ga;
            }
        }
    }
    __syncthreads();
    for (int r = tid; r < N; r += RT) {
        const uint32_t gid = (uint32_t)tgid[r];
        order[base + r] = gid;
        rank[base + gid] = (uint32_t)r;
    }
    if (tid == 0) fallback[blockIdx.x] = 0;
}

// Exclusive scan of the T list lengths in place + size-class histogram of the non-empty lists, three launches:
//   scan_reduce : one CTA per 4096 lengths -> chunk total, non-empty count, class histogram (global atomics)
//   scan_spine  : one CTA scans the chunk totals; class bases (longest class first); mailbox = {M, non-empty}
//   scan_apply  : one CTA per chunk rescans it with its base and writes the offsets
constexpr int SCAN_CHUNK = 4096;

__device__ __forceinline__ long long block_exclusive_scan_1024(long long v, long long *s_warp, long long *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long n = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const long long w = s_warp[lane];
        long long wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long n = __shfl_up_sync(FULL, wi, d);
            if (lane >= d) wi += n;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    if (total) *total = s_warp[32];
    return s_warp[wid] + incl - v;
}

__global__ void __launch_bounds__(1024)
scan_reduce_kernel(const int32_t *__restrict__ counts, int T, long long *__restrict__ chunk_sum, int32_t *__restrict__ cls_count,
                   int32_t *__restrict__ nz_total)
{
    __shared__ long long s_warp[33];
    __shared__ int s_cls[PS_N_CLASSES];
    __shared__ int s_nz;
    const int tid = threadIdx.x;
    if (tid < PS_N_CLASSES) s_cls[tid] = 0;
    if (tid == 0) s_nz = 0;
    __syncthreads();
    const int base = blockIdx.x * SCAN_CHUNK;
    long long acc = 0;
    int nz = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / 1024; ++k) {
        const int i = base + k * 1024 + tid;
        const int c = i < T ? counts[i] : 0;
        acc += c;
        if (c > 0) { ++nz; atomicAdd(&s_cls[31 - __clz(c)], 1); }
    }
    if (nz) atomicAdd(&s_nz, nz);
    long long total;
    block_exclusive_scan_1024(acc, s_warp, &total);
    if (tid == 0) { chunk_sum[blockIdx.x] = total; if (s_nz) atomicAdd(nz_total, s_nz); }
    if (tid < PS_N_CLASSES && s_cls[tid]) atomicAdd(&cls_count[tid], s_cls[tid]);
}

// cls layout: [0, 32) class bases, [32, 64) fill counters (zeroed here), [64, 96) class counts, [96] non-empty total
__global__ void __launch_bounds__(1024)
scan_spine_kernel(long long *__restrict__ chunk_sum, int n_chunks, int32_t *__restrict__ cls, int32_t *__restrict__ offsets, int T,
                  int64_t *__restrict__ mailbox)
{
    __shared__ long long s_warp[33];
    __shared__ long long s_carry;
    const int tid = threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_chunks; c0 += 1024) {
        const int i = c0 + tid;
        const long long v = i < n_chunks ? chunk_sum[i] : 0;
        long long total;
        const long long ex = block_exclusive_scan_1024(v, s_warp, &total);
        if (i < n_chunks) chunk_sum[i] = s_carry + ex;
        __syncthreads();
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    if (tid == 0) {
        const long long total = s_carry;
        offsets[T] = (int32_t)(total > 0x7fffffffLL ? 0x7fffffffLL : total);
        mailbox[0] = total; // mapped host memory: visible to the host once the stream has drained
        mailbox[1] = cls[3 * PS_N_CLASSES];
        __threadfence_system();
        int r = 0;
        for (int c = PS_N_CLASSES - 1; c >= 0; --c) {
            cls[c] = r;
            r += cls[2 * PS_N_CLASSES + c];
            cls[PS_N_CLASSES + c] = 0;
        }
    }
}

__global__ void __launch_bounds__(1024)
scan_apply_kernel(int32_t *__restrict__ offsets, int T, const long long *__restrict__ chunk_base)
{
    __shared__ long long s_warp[33];
    const int tid = threadIdx.x;
    const int base = blockIdx.x * SCAN_CHUNK + tid * (SCAN_CHUNK / 1024);
    int c[SCAN_CHUNK / 1024];
    long long acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / 1024; ++k) {
        c[k] = base + k < T ? offsets[base + k] : 0;
        acc += c[k];
    }
    long long run = chunk_base[blockIdx.x] + block_exclusive_scan_1024(acc, s_warp, nullptr);
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / 1024; ++k) {
        if (base + k < T) offsets[base + k] = (int32_t)run;
        run += c[k];
    }
}

__global__ void __launch_bounds__(256)
build_worklist_kernel(const int32_t *__restrict__ offsets, int T, int32_t *__restrict__ cls, int32_t *__restrict__ worklist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int c = i < T ? offsets[i + 1] - offsets[i] : 0;
    const uint32_t mask = __ballot_sync(FULL, c > 0);
    if (c > 0) {
        const int kcls = 31 - __clz(c);
        const uint32_t peers = __match_any_sync(mask, kcls);
        const int leader = __ffs(peers) - 1;
        int b = 0;
        if (lane == leader) b = atomicAdd(&cls[PS_N_CLASSES + kcls], __popc(peers));
        b = __shfl_sync(peers, b, leader);
        worklist[cls[kcls] + b + __popc(peers & ((1u << lane) - 1u))] = i;
    }
}

// Every listed Gaussian drops one slot word into the list of every tile it touches: its depth rank (3D) / row index
// (2D) in the low 24 bits and, above them, the 8-bit mask of the tile's 8x4 pixel blocks its footprint can reach
// (from the record the thread already holds: the block lists are later built from these masks without touching the
// records again).  MASKS = 1: the blocks met by the footprint's bounding box (a few per cent more block-list entries
// than the exact ellipse-vs-block test of MASKS = 2, at a tenth of the instructions; the rasterizers' per-pixel tests
// decide anyway).  MASKS = 0: plain keys (the block split gathers the records itself).
template <int MODE, int MASKS> // MASKS: 0 = plain keys, 1 = block rectangle of the footprint's bounding box, 2 = exact test
__global__ void __launch_bounds__(PS_PROJ_BLOCK)
partition_kernel(PsGeometry g, PsTable t, const int32_t *__restrict__ offsets, int32_t *__restrict__ fill,
                 uint32_t *__restrict__ slots, int use_smem)
{
    extern __shared__ int s_dyn[]; // [2 * n_tiles] when use_smem
    const int v = blockIdx.y;
    const int gi = blockIdx.x * PS_PROJ_BLOCK + threadIdx.x;
    const bool live = gi < g.N;
    const size_t idx = (size_t)v * g.N + (live ? gi : 0);
    const int touched = live ? t.tiles_touched[idx] : 0;
    int tx0 = 0, ty0 = 0, tx1 = 0, ty1 = 0;
    uint32_t val = 0;
    float gx = 0.0f, gy = 0.0f, thr = 0.0f, hA = 0.0f, B = 0.0f, hC = 0.0f;
    const float half = (MODE == PS_MODE_3D) ? 0.5f : 0.0f;
    PsBlockRect brect = { 1, 0, 1, 0 };
    if (touched) {
        const uint2 tr = t.tile_rect[idx];
        tx0 = tr.x & 0xffff; ty0 = tr.x >> 16; tx1 = tr.y & 0xffff; ty1 = tr.y >> 16;
        val = (MODE == PS_MODE_3D) ? t.rank[idx] : (uint32_t)gi;
        if (MASKS) {
            const float4 r0 = __ldg(PS_REC(t, idx, 0)), r1 = __ldg(PS_REC(t, idx, 1));
            gx = r0.x; gy = r0.y; thr = r0.z;
            hA = r1.x; B = r1.y; hC = r1.z;
            if (MODE == PS_MODE_2D) ps_conic2d(r1, hA, B, hC);
            float ex, ey;
            ps_footprint_box(hA, B, hC, thr, ex, ey);
            brect = ps_block_rect(gx, gy, ex, ey, half, g.W, g.H);
        }
    }
    auto slot_word = [&](int tx, int ty) -> uint32_t {
        if (!MASKS) return val;
        const uint32_t m8 = MASKS == 2 ? (ps_block_mask8(gx, gy, hA, B, hC, thr, half, tx, ty) & ps_blocks_inside8(tx, ty, g.W, g.H))
                                       : ps_block_mask8_rect(brect, tx, ty);
        return val | (m8 << PS_SLOT_MASK_SHIFT);
    };
    const int32_t *off_v = offsets + (size_t)v * g.n_tiles;
    int32_t *fill_v = fill + (size_t)v * g.n_tiles;
    if (!use_smem) { // very large tile grids: reserve every slot with a global atomic
        for (int ty = ty0; ty < ty1; ++ty)
            for (int tx = tx0; tx < tx1; ++tx) {
                const int tl = ty * g.tiles_x + tx;
                slots[off_v[tl] + atomicAdd(&fill_v[tl], 1)] = slot_word(tx, ty);
            }
        return;
    }
    int *s_cnt = s_dyn, *s_base = s_dyn + g.n_tiles;
    for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) s_cnt[i] = 0;
    __syncthreads();
    for (int ty = ty0; ty < ty1; ++ty)
        for (int tx = tx0; tx < tx1; ++tx) atomicAdd(&s_cnt[ty * g.tiles_x + tx], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < g.n_tiles; i += PS_PROJ_BLOCK) {
        const int c = s_cnt[i];
        if (c) {
            s_base[i] = off_v[i] + atomicAdd(&fill_v[i], c); // one reservation per (CTA, tile)
            s_cnt[i] = 0;
        }
    }
    __syncthreads();
    for (int ty = ty0; ty < ty1; ++ty)
        for (int tx = tx0; tx < tx1; ++tx) {
            const int tl = ty * g.tiles_x + tx;
            slots[s_base[tl] + atomicAdd(&s_cnt[tl], 1)] = slot_word(tx, ty);
        }
}

__device__ __forceinline__ int block_exclusive_scan_256i(int v, int *s_warp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += n;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int pre = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) pre += (w < wid) ? s_warp[w] : 0;
    return pre + incl - v;
}

// One CTA per non-empty list: unique keys < N -> bitmap sort.  The sorted keys are enumerated into a shared-memory
// staging array (every thread owns a run of bitmap words and writes its run), then leave in one coalesced sweep that
// also maps depth rank -> Gaussian (3D), four gathers in flight: the list is written to HBM once, in order.
// With m8s != NULL the block masks packed into the slot words travel along: a byte table indexed by key in shared
// memory, read back in sorted order -> m8s [M], so that the block split streams (id, mask) pairs instead of gathering
// records.  STAGED = false: lists longer than the staging capacity (N > 65535) write the sorted keys to vals first.
template <int MODE, bool STAGED>
__global__ void __launch_bounds__(256)
sort_lists_kernel(PsGeometry g, const uint32_t *__restrict__ order, const int32_t *__restrict__ offsets,
                  const int32_t *__restrict__ worklist, const uint32_t *__restrict__ slots, uint32_t *__restrict__ vals,
                  uint8_t *__restrict__ m8s, const int32_t *__restrict__ n_lists)
{
    extern __shared__ uint32_t s_bm[]; // [(N + 31) / 32] | STAGED: sorted keys uint16 [N] | bytes [N] (m8s)
    __shared__ int s_warp[8];
    if ((int)blockIdx.x >= __ldg(n_lists)) return; // the grid may be an upper bound (sync-free small calls)
    const int lin = worklist[blockIdx.x];
    const int start = offsets[lin], end = offsets[lin + 1];
    const int view = lin / g.n_tiles;
    const int words = (g.N + 31) >> 5;
    uint16_t *stage = reinterpret_cast<uint16_t *>(s_bm + words);
    uint8_t *tab = reinterpret_cast<uint8_t *>(s_bm + words) + (STAGED ? 2 * (size_t)((g.N + 1) & ~1) : 0);
    for (int w = threadIdx.x; w < words; w += 256) s_bm[w] = 0u;
    __syncthreads();
    for (int i0 = start + threadIdx.x; i0 < end; i0 += 4 * 256) { // four independent loads in flight per thread
        uint32_t sw4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) sw4[u] = i0 + u * 256 < end ? __ldg(slots + i0 + u * 256) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * 256 >= end) continue;
            const uint32_t r = sw4[u] & PS_SLOT_KEY_MASK;
            atomicOr(&s_bm[r >> 5], 1u << (r & 31u));
            if (m8s) tab[r] = (uint8_t)(sw4[u] >> PS_SLOT_MASK_SHIFT);
        }
    }
    __syncthreads();
    const int wpt = (words + 255) / 256;
    const int w0 = min(words, (int)threadIdx.x * wpt), w1 = min(words, w0 + wpt);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(s_bm[w]);
    int out = block_exclusive_scan_256i(cnt, s_warp); // position inside the list
    const uint32_t vbase = (uint32_t)view * (uint32_t)g.N;
    for (int w = w0; w < w1; ++w) { // sorted ranks (3D) / row indices (2D), in order
        uint32_t bits = s_bm[w];
        while (bits) {
            const uint32_t r = (uint32_t)(w << 5) + (uint32_t)(__ffs(bits) - 1);
            bits &= bits - 1;
            if (STAGED) stage[out] = (uint16_t)r;
            else {
                if (m8s) m8s[start + out] = tab[r];
                vals[start + out] = (MODE == PS_MODE_3D) ? r : vbase + r;
            }
            ++out;
        }
    }
    if (!STAGED && MODE != PS_MODE_3D) return;
    __syncthreads();
    const int len = end - start;
    for (int j0 = threadIdx.x; j0 < len; j0 += 4 * 256) { // rank -> Gaussian, four gathers in flight, coalesced stores
        uint32_t r4[4], g4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 256;
            r4[u] = j < len ? (STAGED ? (uint32_t)stage[j] : vals[start + j]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) g4[u] = (MODE == PS_MODE_3D && j0 + u * 256 < len) ? __ldg(order + vbase + r4[u]) : r4[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * 256;
            if (j >= len) continue;
            vals[start + j] = vbase + g4[u];
            if (STAGED && m8s) m8s[start + j] = tab[r4[u]];
        }
    }
}

// One CTA per non-empty list: sorts it AND splits it into the lists of the tile's eight 8x4 pixel blocks, without
// reading a single splat record.  The keys of a list are unique integers < N, so nine bitmaps in shared memory (the
// whole list + one per block, filled from the slot words' block masks) hold the sorted lists implicitly.
//   phase 1  (thread = list entry, four in flight): set the entry's bit in the bitmaps of its blocks; 3D: look up the
//            Gaussian of this depth rank once and park it in a shared-memory table indexed by rank
//   phase 2  (warp k = block k): enumerate bitmap k 32 words at a time (popcount + warp scan), every lane writes the
//            ids of its word's set bits as one contiguous run: ordered output, neighbouring lanes write neighbouring runs
//   blist [8 start + k len + j]  = view * N + Gaussian   of the j-th entry (in depth / row order) of block k
//   KEEP (parity taps): vals [start + pos] = the sorted tile list itself (gsplat's flatten_ids) and
//   bpos [8 start + k len + j] = pos, the entry's position in it (for the last-id tap)
// Shared memory: 9 words per 32 keys (+ 4 N bytes for the rank -> id table in 3D when it fits, + 1 word per 32 keys KEEP).
constexpr int SPLIT_THREADS = 256;
constexpr int SPLIT_MAPS = 9;
constexpr int SPLIT_MLP = 4; // list entries in flight per thread in phase 1
constexpr int SPLIT_WPL = 4; // bitmap words per lane and step in phase 2

template <int MODE, bool KEEP>
__global__ void __launch_bounds__(SPLIT_THREADS)
sort_split_kernel(PsGeometry g, const uint32_t *__restrict__ order, const int32_t *__restrict__ offsets,
                  const int32_t *__restrict__ worklist, const uint32_t *__restrict__ slots, uint32_t *__restrict__ vals,
                  uint32_t *__restrict__ blist, uint32_t *__restrict__ bpos, int32_t *__restrict__ bcount,
                  const int32_t *__restrict__ n_lists, int use_table)
{
    extern __shared__ uint32_t s_dyn32[]; // bm [9][words] | pre [words] (KEEP) | ids [N] (3D, use_table)
    __shared__ int s_wsum[SPLIT_THREADS / 32];
    const int item = blockIdx.x;
    if (item >= __ldg(n_lists)) return; // the grid may be an upper bound (sync-free small calls)
    const int lin = worklist[item];
    const int start = offsets[lin], end = offsets[lin + 1], len = end - start;
    const int view = lin / g.n_tiles;
    const int words = (g.N + 31) >> 5;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t *bm = s_dyn32;
    uint32_t *pre = bm + SPLIT_MAPS * words;
    uint32_t *ids = (MODE == PS_MODE_3D && use_table) ? pre + (KEEP ? words : 0) : nullptr;
    const uint32_t vbase = (uint32_t)view * (uint32_t)g.N;
    for (int w = tid; w < SPLIT_MAPS * words; w += SPLIT_THREADS) bm[w] = 0u;
    __syncthreads();
    for (int i0 = start + tid; i0 < end; i0 += SPLIT_MLP * SPLIT_THREADS) {
        uint32_t sw[SPLIT_MLP], gid[SPLIT_MLP];
#pragma unroll
        for (int u = 0; u < SPLIT_MLP; ++u) {
            const int i = i0 + u * SPLIT_THREADS;
            sw[u] = i < end ? __ldg(slots + i) : 0xffffffffu;
        }
        if (ids) {
#pragma unroll
            for (int u = 0; u < SPLIT_MLP; ++u)
                gid[u] = sw[u] != 0xffffffffu ? __ldg(order + vbase + (sw[u] & PS_SLOT_KEY_MASK)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < SPLIT_MLP; ++u) {
            if (i0 + u * SPLIT_THREADS >= end) continue;
            const uint32_t r = sw[u] & PS_SLOT_KEY_MASK;
            uint32_t m8 = sw[u] >> PS_SLOT_MASK_SHIFT;
            const uint32_t bit = 1u << (r & 31u);
            atomicOr(&bm[r >> 5], bit);
            if (ids) ids[r] = vbase + gid[u];
            while (m8) {
                const int k = __ffs(m8) - 1;
                m8 &= m8 - 1;
                atomicOr(&bm[(1 + k) * words + (r >> 5)], bit);
            }
        }
    }
    __syncthreads();
    if (KEEP) { // the sorted tile list itself + the word prefixes that give every entry its position in it
        const int wpt = (words + SPLIT_THREADS - 1) / SPLIT_THREADS;
        const int w0 = min(words, tid * wpt), w1 = min(words, w0 + wpt);
        int c = 0;
        for (int w = w0; w < w1; ++w) c += __popc(bm[w]);
        int x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(FULL, x, d);
            if (lane >= d) x += n;
        }
        if (lane == 31) s_wsum[wid] = x;
        __syncthreads();
        int run = x - c;
#pragma unroll
        for (int w = 0; w < SPLIT_THREADS / 32; ++w) run += (w < wid) ? s_wsum[w] : 0;
        for (int w = w0; w < w1; ++w) {
            pre[w] = (uint32_t)run;
            uint32_t bits = bm[w];
            while (bits) {
                const uint32_t r = (uint32_t)(w << 5) + (uint32_t)(__ffs(bits) - 1);
                bits &= bits - 1;
                vals[start + run++] = (MODE == PS_MODE_3D) ? (ids ? ids[r] : vbase + __ldg(order + vbase + r)) : vbase + r;
            }
        }
        __syncthreads();
    }
    // phase 2: warp k enumerates block k's bitmap, 128 keys per lane and step
    {
        const int k = wid; // SPLIT_THREADS / 32 == 8 blocks
        const uint32_t *map = bm + (1 + k) * words;
        uint32_t *out = blist + 8 * (size_t)start + (size_t)k * len;
        uint32_t *outp = (KEEP && bpos) ? bpos + 8 * (size_t)start + (size_t)k * len : nullptr;
        int run = 0;
        for (int wb = 0; wb < words; wb += 32 * SPLIT_WPL) {
            const int w0 = wb + lane * SPLIT_WPL;
            uint32_t bits[SPLIT_WPL];
            int c = 0;
#pragma unroll
            for (int q = 0; q < SPLIT_WPL; ++q) {
                bits[q] = w0 + q < words ? map[w0 + q] : 0u;
                c += __popc(bits[q]);
            }
            int x = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(FULL, x, d);
                if (lane >= d) x += n;
            }
            const int total = __shfl_sync(FULL, x, 31);
            if (total == 0) continue;
            int pos = run + x - c;
#pragma unroll
            for (int q = 0; q < SPLIT_WPL; ++q) {
                uint32_t bq = bits[q];
                while (bq) {
                    const uint32_t b = (uint32_t)(__ffs(bq) - 1);
                    bq &= bq - 1;
                    const uint32_t r = (uint32_t)((w0 + q) << 5) + b;
                    out[pos] = (MODE == PS_MODE_3D) ? (ids ? ids[r] : vbase + __ldg(order + vbase + r)) : vbase + r;
                    if (KEEP && outp) outp[pos] = pre[w0 + q] + __popc(bm[w0 + q] & ((1u << b) - 1u));
                    ++pos;
                }
            }
            run += total;
        }
        if (lane == 0) bcount[item * 8 + k] = run;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
debug_keys_kernel(PsGeometry g, const uint32_t *__restrict__ depth, const int32_t *__restrict__ offsets,
                  const int32_t *__restrict__ worklist, const uint32_t *__restrict__ vals, uint64_t *__restrict__ keys)
{
    const int lin = worklist[blockIdx.x];
    const int start = offsets[lin], end = offsets[lin + 1];
    const int view = lin / g.n_tiles, tile = lin - view * g.n_tiles;
    const uint64_t hi = (((uint64_t)view << g.tile_bits) | (uint64_t)tile) << 32;
    for (int i = start + threadIdx.x; i < end; i += 256) {
        const uint32_t id = vals[i];
        const uint32_t low = (MODE == PS_MODE_3D) ? depth[id] : id - (uint32_t)view * (uint32_t)g.N;
        keys[i] = hi | low;
    }
}

size_t rank_smem_bytes(int N) { return (size_t)N * 12; }
constexpr size_t RANK_STATIC_SMEM = 256 * 4 + 8 * 4 + 8;
constexpr size_t RANK_CNT_SMEM16 = 2 * RW * 256 * 2, RANK_CNT_SMEM32 = 2 * RW * 256 * 4;
constexpr size_t SMEM_LIMIT = 227 * 1024;

} // namespace

size_t ps_rank_scratch_elems(const PsGeometry &g)
{
    if (g.mode != PS_MODE_3D || g.N == 0 || g.V == 0) return 0;
    if (g.N <= 65535 && rank_smem_bytes(g.N) + RANK_CNT_SMEM16 + RANK_STATIC_SMEM + 1024 <= SMEM_LIMIT) return 0;
    return (size_t)g.V * 4 * (size_t)g.N;
}

// buckets of the one-pass ranking: the power of two >= N, between RT and 16384; 0 = N too large for its shared memory
static int rank_buckets(int N)
{
    if (N > 65535) return 0;
    int nb = RT;
    while (nb < N && nb < 16384) nb <<= 1;
    const size_t dyn = (size_t)(nb + 1) * 4 + (size_t)N * 4 + (size_t)((N + 1) & ~1) * 2 + (size_t)N * 2;
    return dyn + 1024 <= SMEM_LIMIT ? nb : 0;
}

int ps_launch_depth_rank(const PsGeometry &g, const PsTable &t, uint32_t *scratch, int32_t *flags, cudaStream_t s)
{
    if (g.mode != PS_MODE_3D || g.N == 0 || g.V == 0) return 0;
    int launches = 0;
    static const bool radix_only = getenv("PS_RANK_RADIX") != nullptr; // A/B switch for measurements
    const int nb = (flags && !radix_only) ? rank_buckets(g.N) : 0;
    const int32_t *only_flagged = nullptr;
    if (nb) {
        const size_t dyn = (size_t)(nb + 1) * 4 + (size_t)g.N * 4 + (size_t)((g.N + 1) & ~1) * 2 + (size_t)g.N * 2;
        if (cudaFuncSetAttribute(depth_rank_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -1;
        depth_rank_bucket_kernel<<<g.V, RT, dyn, s>>>(g.N, nb, t.depth, t.tiles_touched, t.order, t.rank, flags);
        only_flagged = flags; // the radix kernel below only ranks the views the bucket pass gave up on
        ++launches;
    }
    if (ps_rank_scratch_elems(g) == 0) {
        const size_t dyn = rank_smem_bytes(g.N) + RANK_CNT_SMEM16;
        if (cudaFuncSetAttribute(depth_rank_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
            return -1;
        depth_rank_kernel<uint16_t><<<g.V, RT, dyn, s>>>(g.N, t.depth, t.tiles_touched, t.order, t.rank, nullptr, only_flagged);
    } else {
        if (!scratch) return -1;
        if (cudaFuncSetAttribute(depth_rank_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RANK_CNT_SMEM32) != cudaSuccess)
            return -1;
        depth_rank_kernel<uint32_t><<<g.V, RT, RANK_CNT_SMEM32, s>>>(g.N, t.depth, t.tiles_touched, t.order, t.rank, scratch, only_flagged);
    }
    return cudaGetLastError() == cudaSuccess ? launches + 1 : -1;
}

size_t ps_scan_scratch_elems(const PsGeometry &g) { return (size_t)(g.V * g.n_tiles + SCAN_CHUNK - 1) / SCAN_CHUNK + 1; }

int ps_launch_scan_lists(const PsGeometry &g, const PsLists &l, long long *chunk_scratch, int64_t *mailbox, cudaStream_t s)
{
    const int T = g.V * g.n_tiles;
    const int n_chunks = (T + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (cudaMemsetAsync(l.cls, 0, sizeof(int32_t) * PS_CLS_WORDS, s) != cudaSuccess) return -1;
    if (n_chunks > 0) scan_reduce_kernel<<<n_chunks, 1024, 0, s>>>(l.offsets, T, chunk_scratch, l.cls + 2 * PS_N_CLASSES, l.cls + 3 * PS_N_CLASSES);
    scan_spine_kernel<<<1, 1024, 0, s>>>(chunk_scratch, n_chunks, l.cls, l.offsets, T, mailbox);
    if (n_chunks > 0) scan_apply_kernel<<<n_chunks, 1024, 0, s>>>(l.offsets, T, chunk_scratch);
    return cudaGetLastError() == cudaSuccess ? (n_chunks > 0 ? 3 : 1) : -1;
}

int ps_launch_partition(const PsGeometry &g, const PsTable &t, const PsLists &l, int masks, cudaStream_t s)
{
    if (g.N == 0 || g.V == 0) return 0;
    dim3 grid((g.N + PS_PROJ_BLOCK - 1) / PS_PROJ_BLOCK, g.V);
    const int use_smem = g.n_tiles <= PS_HIST_SMEM_TILES;
    const size_t dyn = use_smem ? (size_t)2 * g.n_tiles * sizeof(int) : 0;
#define PS_PART(MODE, MK)                                                                                                     \
    do {                                                                                                                      \
        if (dyn > 48 * 1024 && cudaFuncSetAttribute(partition_kernel<MODE, MK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -1; \
        partition_kernel<MODE, MK><<<grid, PS_PROJ_BLOCK, dyn, s>>>(g, t, l.offsets, l.fill, l.slots, use_smem);                  \
    } while (0)
    if (g.mode == PS_MODE_3D) { if (masks == 2) PS_PART(PS_MODE_3D, 2); else if (masks == 1) PS_PART(PS_MODE_3D, 1); else PS_PART(PS_MODE_3D, 0); }
    else { if (masks == 2) PS_PART(PS_MODE_2D, 2); else if (masks == 1) PS_PART(PS_MODE_2D, 1); else PS_PART(PS_MODE_2D, 0); }
#undef PS_PART
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_build_worklist(const PsGeometry &g, const PsLists &l, cudaStream_t s)
{
    const int T = g.V * g.n_tiles;
    if (T == 0) return 0;
    build_worklist_kernel<<<(T + 255) / 256, 256, 0, s>>>(l.offsets, T, l.cls, l.worklist);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_sort_lists(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, uint8_t *m8s, cudaStream_t s)
{
    if (n_work <= 0) return 0;
    const size_t bm = (size_t)((g.N + 31) / 32) * sizeof(uint32_t);
    const size_t stage = 2 * (size_t)((g.N + 1) & ~1);
    static const bool no_stage = getenv("PS_SORT_UNSTAGED") != nullptr; // A/B switch for measurements
    const bool staged = !no_stage && g.N <= 65535 && bm + stage + (m8s ? (size_t)g.N : 0) <= 100 * 1024;
    const size_t dyn = bm + (staged ? stage : 0) + (m8s ? (size_t)g.N : 0);
#define PS_SORT(MODE, ST)                                                                                                     \
    do {                                                                                                                      \
        if (dyn > 48 * 1024 && cudaFuncSetAttribute(sort_lists_kernel<MODE, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -1; \
        sort_lists_kernel<MODE, ST><<<n_work, 256, dyn, s>>>(g, t.order, l.offsets, l.worklist, l.slots, l.vals, m8s, l.n_lists); \
    } while (0)
    if (g.mode == PS_MODE_3D) { if (staged) PS_SORT(PS_MODE_3D, true); else PS_SORT(PS_MODE_3D, false); }
    else { if (staged) PS_SORT(PS_MODE_2D, true); else PS_SORT(PS_MODE_2D, false); }
#undef PS_SORT
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// the mask-byte table of sort_lists needs N bytes of shared memory beside the bitmap
bool ps_mask_bytes_fit_smem(const PsGeometry &g) { return (size_t)((g.N + 31) / 32) * 4 + (size_t)g.N <= 200 * 1024; }

// the fused sort + block split needs 9 (+1) shared-memory words per 32 Gaussians for its bitmaps
static size_t split_map_bytes(const PsGeometry &g, bool keep) { return (size_t)((g.N + 31) / 32) * (SPLIT_MAPS + (keep ? 1 : 0)) * sizeof(uint32_t); }
bool ps_split_fits_smem(const PsGeometry &g) { return split_map_bytes(g, true) <= 160 * 1024; }

int ps_launch_sort_split(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, cudaStream_t s)
{
    if (n_work <= 0) return 0;
    const bool keep = l.vals != nullptr;
    size_t dyn = split_map_bytes(g, keep);
    // 3D: rank -> Gaussian table in shared memory when it leaves room for two CTAs per SM, else gathers in phase 2
    static const bool table_ok = getenv("PS_SPLIT_TABLE") != nullptr; // A/B switch for measurements
    const int use_table = table_ok && g.mode == PS_MODE_3D && dyn + (size_t)g.N * sizeof(uint32_t) <= 100 * 1024;
    if (use_table) dyn += (size_t)g.N * sizeof(uint32_t);
#define PS_SPLIT(MODE, KP)                                                                                                    \
    do {                                                                                                                      \
        if (dyn > 48 * 1024 && cudaFuncSetAttribute(sort_split_kernel<MODE, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -1; \
        sort_split_kernel<MODE, KP><<<n_work, SPLIT_THREADS, dyn, s>>>(g, t.order, l.offsets, l.worklist, l.slots, l.vals, l.blist, l.bpos, l.bcount, l.n_lists, use_table); \
    } while (0)
    if (g.mode == PS_MODE_3D) { if (keep) PS_SPLIT(PS_MODE_3D, true); else PS_SPLIT(PS_MODE_3D, false); }
    else { if (keep) PS_SPLIT(PS_MODE_2D, true); else PS_SPLIT(PS_MODE_2D, false); }
#undef PS_SPLIT
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int ps_launch_debug_keys(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, uint64_t *keys, cudaStream_t s)
{
    if (n_work <= 0) return 0;
    if (g.mode == PS_MODE_3D) debug_keys_kernel<PS_MODE_3D><<<n_work, 256, 0, s>>>(g, t.depth, l.offsets, l.worklist, l.vals, keys);
    else debug_keys_kernel<PS_MODE_2D><<<n_work, 256, 0, s>>>(g, t.depth, l.offsets, l.worklist, l.vals, keys);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
