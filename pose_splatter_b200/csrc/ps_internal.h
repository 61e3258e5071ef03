// ps_internal.h -- host-side launcher declarations shared by the translation units of libpsplat.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psplat.h"

#define PS_PROJ_BLOCK 256   // (view, Gaussian) pairs per projection / emission block
#define PS_RASTER_BATCH 256 // tile-list entries staged in shared memory per round
#define PS_ACC_STRIDE 12    // floats per (view, Gaussian) gradient accumulator row (9 used, 16-byte aligned)

struct PsGeometry {
    int mode, W, H, F, N, V;
    int tiles_x, tiles_y, n_tiles;
    int tile_bits, view_bits;
    float near_plane, far_plane, radius_clip, eps2d;
};

// per-(view,Gaussian) table produced by the projection stage
struct PsTable {
    float4 *rec0, *rec1, *rec2; // [V*N]
    uint2 *tile_rect;           // [V*N] packed tx0|ty0<<16, tx1|ty1<<16
    int32_t *tiles_touched;     // [V*N]
    int32_t *block_sums;        // [ceil(V*N / PS_PROJ_BLOCK) + 1] -> exclusive offsets after the scan
};

// every launcher returns the number of kernels it launched (for gpu_launches) or -1 on error
int ps_launch_project(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                      const float *Ks, const PsTable &t, cudaStream_t s);
int ps_launch_scan_block_sums(const PsGeometry &g, const PsTable &t, int64_t *total_out, cudaStream_t s);
int ps_launch_emit(const PsGeometry &g, const PsTable &t, uint64_t *keys, uint32_t *vals, cudaStream_t s);
int ps_launch_project_bwd(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                          const float *Ks, const PsTable &t, const float *acc, float *d_params, cudaStream_t s);

// stable LSD radix sort of (key, val) pairs on key bits [bit_lo, bit_hi); result ends in keys/vals
// (alt buffers are scratch).  hist is scratch of ps_sort_hist_elems() uint32.
size_t ps_sort_hist_elems(int64_t M);
int ps_launch_sort(uint64_t *keys, uint32_t *vals, uint64_t *keys_alt, uint32_t *vals_alt, int64_t M, int bit_lo,
                   int bit_hi, uint32_t *hist, int *passes_out, cudaStream_t s);
int ps_launch_tile_ranges(const PsGeometry &g, const uint64_t *keys, int64_t M, int32_t *offsets, cudaStream_t s);

int ps_launch_raster_fwd(const PsGeometry &g, const PsTable &t, const uint32_t *vals, const int32_t *offsets,
                         const float *background, float *rgb, float *alpha, int32_t *n_contrib, int32_t *last,
                         float *t_pen, unsigned long long *stats, cudaStream_t s);
int ps_launch_raster_bwd(const PsGeometry &g, const PsTable &t, const uint32_t *vals, const int32_t *offsets,
                         const float *background, const int32_t *last, const float *t_pen, const float *d_rgb,
                         const float *d_alpha, float *acc, cudaStream_t s);

int ps_launch_math_probe(const float *x, int n, float *y, cudaStream_t s);
int ps_launch_fp32_probe(float *sink, int iters, cudaStream_t s);
