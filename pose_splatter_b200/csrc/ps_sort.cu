// ps_sort.cu -- tile binning back half (SURVEY.md 2.2 K3', K4'): hand-written stable LSD radix
// sort of (int64 key, int32 value) pairs and tile-range extraction.  Replaces gsplat's
// cub::DeviceRadixSort::SortPairs + isect_offset_encode (absent from the reference tree).
//
// One 8-bit pass = three launches over a fixed grid of G CTAs, each owning a contiguous span:
//   hist    : per-CTA digit histogram (shared-memory atomics)            reads 8 B / key
//   scan    : exclusive scan of the 256 x G matrix in digit-major order  (one CTA, L2-resident)
//   scatter : CTA walks its span 256 keys at a time; rank = warp match_any + cross-warp counts,
//             so equal digits keep their input order (stable)           reads 12 B, writes 12 B / pair
// All passes are HBM-bound byte shuffling: no tensor cores, grid sized to the SM count.
#include "ps_contract.cuh"
#include "ps_internal.h"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_MAX_BLOCKS = 148 * 2;

struct SortPlan {
    int blocks;
    int64_t span; // elements per CTA, multiple of SORT_THREADS
};

inline SortPlan make_plan(int64_t M)
{
    SortPlan p;
    int64_t chunks = (M + SORT_THREADS - 1) / SORT_THREADS;
    p.blocks = (int)(chunks < SORT_MAX_BLOCKS ? (chunks < 1 ? 1 : chunks) : SORT_MAX_BLOCKS);
    int64_t per = (chunks + p.blocks - 1) / p.blocks;
    p.span = per * SORT_THREADS;
    return p;
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint64_t *__restrict__ keys, int64_t M, int64_t span, int shift, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * span;
    const int64_t hi = lo + span < M ? lo + span : M;
    for (int64_t i = lo + threadIdx.x; i < hi; i += SORT_THREADS) {
        const uint32_t d = (uint32_t)(keys[i] >> shift) & 255u;
        atomicAdd(&s_hist[d], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_hist[threadIdx.x];
}

// exclusive scan over n = 256*G counters, one CTA of 1024 threads
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *hist, int n)
{
    __shared__ uint32_t s_part[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    uint32_t acc = 0;
    for (int i = lo; i < hi; ++i) acc += hist[i];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t add = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - acc;
    for (int i = lo; i < hi; ++i) {
        uint32_t c = hist[i];
        hist[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t M, int64_t span,
                    int shift, const uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_off[256];                // running output offset of every digit for this CTA
    __shared__ uint32_t s_wcount[SORT_WARPS][256]; // per-warp digit counts of the current 256-key chunk
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_off[threadIdx.x] = hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) s_wcount[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * span;
    const int64_t hi = lo + span < M ? lo + span : M;
    for (int64_t base = lo; base < hi; base += SORT_THREADS) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < hi;
        uint64_t key = 0;
        uint32_t val = 0, digit = 0;
        if (live) {
            key = keys_in[i];
            val = vals_in[i];
            digit = (uint32_t)(key >> shift) & 255u;
        }
        // rank among equal digits inside the warp (lower lane = earlier element)
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        uint32_t peers = 0, rank = 0, count = 0;
        bool leader = false;
        if (live) {
            peers = __match_any_sync(live_mask, digit);
            rank = __popc(peers & ((1u << lane) - 1u));
            count = __popc(peers);
            leader = (rank == 0);
            if (leader) s_wcount[wid][digit] = count;
        }
        __syncthreads();
        uint32_t pos = 0;
        if (live) {
            uint32_t pre = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) pre += (w < wid) ? s_wcount[w][digit] : 0u;
            pos = s_off[digit] + pre + rank;
        }
        __syncthreads();
        if (leader) {
            atomicAdd(&s_off[digit], count);
            s_wcount[wid][digit] = 0;
        }
        if (live) {
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        // no third barrier: the next chunk's s_wcount writes touch only this warp's row, and its
        // reads of s_off / s_wcount come after the next chunk's first barrier
    }
}

// offsets[view*n_tiles + tile] = first sorted index of that (view, tile); offsets[V*n_tiles] = M
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint64_t *__restrict__ keys, int64_t M, int n_tiles, int tile_bits, int total,
                   int32_t *__restrict__ offsets)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint32_t mask = (1u << tile_bits) - 1u;
    const uint32_t hi = (uint32_t)(keys[i] >> 32);
    const int cur = (int)(hi >> tile_bits) * n_tiles + (int)(hi & mask);
    int prev = -1;
    if (i > 0) {
        const uint32_t hp = (uint32_t)(keys[i - 1] >> 32);
        prev = (int)(hp >> tile_bits) * n_tiles + (int)(hp & mask);
    }
    for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    if (i == M - 1)
        for (int t = cur + 1; t <= total; ++t) offsets[t] = (int32_t)M;
}

} // namespace

size_t ps_sort_hist_elems(int64_t M) { return (size_t)256 * make_plan(M).blocks; }

int ps_launch_sort(uint64_t *keys, uint32_t *vals, uint64_t *keys_alt, uint32_t *vals_alt, int64_t M, int bit_lo,
                   int bit_hi, uint32_t *hist, int *passes_out, cudaStream_t s)
{
    int launches = 0, passes = 0;
    if (M > 0) {
        const SortPlan p = make_plan(M);
        uint64_t *kin = keys, *kout = keys_alt;
        uint32_t *vin = vals, *vout = vals_alt;
        for (int shift = bit_lo; shift < bit_hi; shift += 8) {
            sort_hist_kernel<<<p.blocks, SORT_THREADS, 0, s>>>(kin, M, p.span, shift, hist);
            sort_scan_kernel<<<1, 1024, 0, s>>>(hist, 256 * p.blocks);
            sort_scatter_kernel<<<p.blocks, SORT_THREADS, 0, s>>>(kin, vin, kout, vout, M, p.span, shift, hist);
            launches += 3;
            ++passes;
            uint64_t *tk = kin; kin = kout; kout = tk;
            uint32_t *tv = vin; vin = vout; vout = tv;
        }
        if (cudaGetLastError() != cudaSuccess) return -1;
    }
    *passes_out = passes; // odd number of passes: the sorted data is in the alt buffers
    return launches;
}

int ps_launch_tile_ranges(const PsGeometry &g, const uint64_t *keys, int64_t M, int32_t *offsets, cudaStream_t s)
{
    const int total = g.V * g.n_tiles;
    if (M == 0) {
        if (cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)(total + 1), s) != cudaSuccess) return -1;
        return 0;
    }
    tile_ranges_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(keys, M, g.n_tiles, g.tile_bits, total, offsets);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
