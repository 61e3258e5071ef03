// ps_internal.h -- host-side launcher declarations shared by the translation units of libpsplat.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psplat.h"

#define PS_PROJ_BLOCK 256    // (view, Gaussian) pairs per projection / partition block
#define PS_ACC_STRIDE 9      // floats per (view, Gaussian) gradient accumulator row
#define PS_HIST_SMEM_TILES 8192  // per-view tile histograms live in shared memory up to this many tiles
#define PS_RANK_THREADS 1024     // one CTA per view in the depth-ranking kernel
#define PS_N_CLASSES 32          // tile-list size classes (floor(log2(len))) of the work list
#define PS_CLS_WORDS (3 * PS_N_CLASSES + 1)  // class bases | fill counters | class counts | non-empty total

struct PsGeometry {
    int mode, W, H, F, N, V;
    int tiles_x, tiles_y, n_tiles;
    int tile_bits, view_bits;
    float near_plane, far_plane, radius_clip, eps2d;
    int activated; // 3D rows hold activated values (PS_FLAG_ACTIVATED_INPUTS)
};

// per-(view,Gaussian) table produced by the projection stage
struct PsTable {
    float4 *rec;                // [V*N][4] splat records, 64 B each (one DRAM burst per gather): rec0 | rec1 | rec2 | spare
    uint2 *tile_rect;           // [V*N] packed tx0|ty0<<16, tx1|ty1<<16
    int32_t *tiles_touched;     // [V*N]
    uint32_t *depth;            // [V*N] 3D: bits of the camera-space depth (low word of the sort key); 2D: unused
    uint32_t *order;            // [V*N] 3D: order[view*N + r] = Gaussian with depth rank r in that view (2D: unused)
    uint32_t *rank;             // [V*N] 3D: inverse of order                                     (2D: unused)
};

#define PS_REC(t, idx, k) ((t).rec + 4 * (size_t)(idx) + (k))

// per-(view,tile) lists
struct PsLists {
    int32_t *offsets;   // [T+1], T = V*n_tiles: counts after projection, exclusive offsets after the scan
    int32_t *fill;      // [T] running fill of every list during the partition pass
    int32_t *worklist;  // [T] non-empty (view,tile) ids, longest size class first
    int32_t *cls;       // [PS_CLS_WORDS] size-class bases / fill counters / counts of the work list
    uint32_t *slots;    // [M] depth ranks (3D) / row indices (2D) in list order, unsorted inside a list
    uint32_t *vals;     // [M] view*N + Gaussian, sorted (tile, depth | row)
    uint32_t *blist;    // [8*M] per-block lists: tile with range [s, s+len) owns [8s, 8s+8len), block k at +k*len;
                        //       entries = view*N + Gaussian, in tile-list order
    uint32_t *cids;     // [8*M] or NULL: contributor lists, written by the forward rasterizer for its backward, in the regions of
    uint32_t *cmask;    //       blist: view*N + Gaussian and the 32-bit mask of the block's pixels it was composited into, for
    int32_t *ccount;    //       every block-list entry with at least one such pixel, in list order; ccount [8*n_work] = how many
    uint32_t *bpos;     // [8*M] or NULL: position (relative to s) in the tile list of every block-list entry (last-id tap)
    int32_t *bcount;    // [8*n_work] length of every block list, indexed by work-list item
    const int32_t *n_lists; // device: number of non-empty lists (= cls[3 * PS_N_CLASSES]); kernels launched over an upper
                            // bound of it exit beyond this count
};

// every launcher returns the number of kernels it launched (for gpu_launches) or -1 on error
// frame_off / frame_views (the CSR of ps_launch_frame_csr) != NULL in 3D: one thread per (frame, Gaussian), the
// camera-independent part of the projection computed once per frame
int ps_launch_project(const PsGeometry &g, const float *params, const int32_t *view_frame, const float *viewmats,
                      const float *Ks, const PsTable &t, int32_t *tile_counts, const int32_t *frame_off,
                      const int32_t *frame_views, cudaStream_t s);
// view_frame [V] -> CSR (frame_off [F+1], frame_views [V]); cursor [F] is scratch
// also keeps a copy of the forward's background colour (bg_saved [3]) for the backward
int ps_launch_frame_csr(const PsGeometry &g, const int32_t *view_frame, int32_t *frame_off, int32_t *cursor,
                        int32_t *frame_views, const float *background, float *bg_saved, cudaStream_t s);
// peers == nullptr: rows stored into d_params; else rows pushed into slot [my_rank][frame / world] of the staging
// buffer of rank frame % world (peers = device array of the ranks' staging base pointers)
int ps_launch_project_bwd(const PsGeometry &g, const float *params, const int32_t *frame_off, const int32_t *frame_views,
                          const float *viewmats, const float *Ks, const PsTable &t, const float *acc, float *d_params,
                          float *const *peers, int my_rank, int world, cudaStream_t s);
int ps_launch_peer_sum(const float *stage, int world, size_t n, float *out, cudaStream_t s);

// binning (ps_bin.cu)
size_t ps_rank_scratch_elems(const PsGeometry &g); // uint32 elements of global scratch the ranking needs (0 if it fits smem)
// flags [V] int32 scratch (or NULL = radix ranking only): views the one-pass bucket ranking left to the radix kernel
int ps_launch_depth_rank(const PsGeometry &g, const PsTable &t, uint32_t *scratch, int32_t *flags, cudaStream_t s);
// exclusive scan of the T counts in place (offsets[T] = M), size classes; mailbox[0] = M, mailbox[1] = non-empty lists
size_t ps_scan_scratch_elems(const PsGeometry &g); // int64 elements of scratch the scan needs
int ps_launch_scan_lists(const PsGeometry &g, const PsLists &l, long long *chunk_scratch, int64_t *mailbox, cudaStream_t s);
// masks: 0 = plain keys, 1 = block-rectangle masks, 2 = exact block masks packed above the key (PS_SLOT_MASK_SHIFT)
int ps_launch_partition(const PsGeometry &g, const PsTable &t, const PsLists &l, int masks, cudaStream_t s);
int ps_launch_build_worklist(const PsGeometry &g, const PsLists &l, cudaStream_t s);
// m8s [M] or NULL: the slot words' block masks in sorted order (needs ps_mask_bytes_fit_smem)
bool ps_mask_bytes_fit_smem(const PsGeometry &g);
int ps_launch_sort_lists(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, uint8_t *m8s, cudaStream_t s);
// sort + split into the eight block lists in one kernel (no record gathers); needs ps_split_fits_smem(g)
bool ps_split_fits_smem(const PsGeometry &g);
int ps_launch_sort_split(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, cudaStream_t s);
int ps_launch_debug_keys(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, uint64_t *keys, cudaStream_t s);

// rasterizers (ps_raster.cu)
// rgb / alpha / rgba8 may each be NULL (rgba8: uint8 RGBA, one uint32 per pixel, quantised like the reference's writer)
int ps_launch_fill_empty(const PsGeometry &g, const int32_t *offsets, const float *background, float *rgb, float *alpha,
                         int32_t *n_contrib, int32_t *last, uint32_t *rgba8, cudaStream_t s);
// m8s != NULL: stream (id, mask byte) pairs instead of gathering records
int ps_launch_block_lists(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const uint8_t *m8s, cudaStream_t s);
// last: tile-list position + 1 of the last contributor (tap); blast: the same as an index into the block list (backward)
int ps_launch_raster_fwd(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const float *background,
                         float *rgb, float *alpha, int32_t *n_contrib, int32_t *last, int32_t *blast, float *t_pen,
                         uint32_t *rgba8, unsigned long long *stats, cudaStream_t s);
int ps_launch_raster_bwd(const PsGeometry &g, const PsTable &t, const PsLists &l, int n_work, const float *background,
                         const int32_t *last, const float *t_pen, const float *d_rgb, const float *d_alpha, float *acc,
                         unsigned *next_task /* zeroed by the caller: the persistent warps' task counter */,
                         unsigned long long *stats /* NULL, or this kernel's pair counters [4..7] */, cudaStream_t s);

// per-view training loss + its gradient (ps_loss.cu); stats [V*8] doubles and adj [V*9*H*W] floats are scratch
int ps_launch_view_loss(int V, int H, int W, const float *rgb, const float *alpha, const float *timg, const float *mask,
                        float ssim_lambda, float img_lambda, double *stats, float *adj, float *losses, float *d_rgb,
                        float *d_alpha, cudaStream_t s);

int ps_launch_iou_loss(int V, int H, int W, const float *alpha, const float *mask, double *stats, float *losses, float *d_alpha,
                       cudaStream_t s);

// parameter-head tail (ps_head.cu): activations + pose transform of the rows render() takes, forward and backward
int ps_launch_head_fwd(int mode, int n, const float *net_out, const float *probs, const float *grid, const float *scale0,
                       float voxel_size, float pt, float clip_lo, float clip_hi, int pose, double angle, const float *p_host,
                       const float *poses, const int32_t *row_frame, int n_frames, float *rows, cudaStream_t s);
int ps_launch_head_bwd(int mode, int n, const float *net_out, const float *probs, float voxel_size, float pt, float clip_lo,
                       float clip_hi, int pose, double angle, const float *poses, const int32_t *row_frame, int n_frames,
                       const float *d_rows, float *d_net, float *d_probs, float *d_scale0, cudaStream_t s);

int ps_launch_math_probe(const float *x, int n, float *y, cudaStream_t s);
int ps_launch_adapter3d_probe(const float *rows, int n, const float *v_act, float *act, float *d_rows, cudaStream_t s);
int ps_launch_fp32_probe(float *sink, int iters, cudaStream_t s);
