/*
 * psplat.h -- C ABI of libpsplat.so, the B200 (sm_100a) Gaussian-splatting renderer that
 * replaces the arithmetic behind pose-splatter's renderer API.
 *
 * Boundary.  The reference has no native code; its "FFI" for this path is
 *   - src/gaussian_renderer.py:196-208  -> gsplat.rendering.rasterization (third-party CUDA)  [3D]
 *   - src/gaussian_renderer.py:326-328  -> GaussianRenderer2D._render_vectorized (torch ops)  [2D]
 * Each entry point below names the reference interface it replaces.  All pointers are
 * plain DEVICE pointers (fp32 / int32 / int64, C-contiguous), sizes are ints, the stream is
 * a cudaStream_t passed as void*.  No torch types.  Every function returns 0 on success or
 * a non-zero code with a message available from ps_last_error() (thread-local).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PSPLAT_H
#define PSPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_ABI_VERSION 4

#define PS_MODE_2D 2 /* GaussianRenderer2D, rows of 9 floats  (src/gaussian_renderer.py:214-334) */
#define PS_MODE_3D 3 /* GaussianRenderer3D, rows of 14 floats (src/gaussian_renderer.py:110-211) */

#define PS_FLAG_SAVE_FOR_BACKWARD 1 /* keep the state ps_backward needs                        */
#define PS_FLAG_KEEP_BINNING 2      /* also materialise the sorted int64 keys and keep last ids for the debug taps */
#define PS_FLAG_RASTER_STATS 4      /* count (pixel, Gaussian) pairs inside the rasterizers, forward and backward (bench only) */
#define PS_FLAG_ACTIVATED_INPUTS 8  /* 3D rows = means | scales | quats | colours | opacity as gsplat's rasterization()
                                       takes them (src/model.py:342-361): no exp / q/(|q|+1e-8) / clamp / sigmoid, and
                                       the gradient is w.r.t. those values                                          */

/* stages timed by ps_ctx_set_profiling (CUDA events on the launching stream) */
#define PS_STAGE_PROJECT 0     /* activations + projection + per-(view,tile) list lengths            */
#define PS_STAGE_RANK 1        /* 3D: per-view depth ranking                                         */
#define PS_STAGE_SCAN 2        /* list lengths -> tile ranges, M                                     */
#define PS_STAGE_PARTITION 3   /* Gaussians -> lists                                                 */
#define PS_STAGE_SORT 4        /* work list + per-list sort by depth rank / row index                */
#define PS_STAGE_RASTER_FWD 5  /* background fill of empty tiles + forward rasterizer                */
#define PS_STAGE_RASTER_BWD 6
#define PS_STAGE_PROJECT_BWD 7
#define PS_STAGE_BLOCKS 8      /* tile lists -> lists of the eight 8x4 pixel blocks of every tile    */
#define PS_N_STAGES 9

/* what ps_saved_copy can read back (bit-exact parity taps) */
#define PS_TAP_ISECT_KEYS 1    /* int64 [M]   sorted keys  view<<(32+tile_bits) | tile<<32 | low (low = depth bits in 3D, row in 2D);
                                  implied by the lists, materialised for this tap (PS_FLAG_KEEP_BINNING)  */
#define PS_TAP_FLATTEN_IDS 2   /* int32 [M]   sorted values view*N + gaussian                     */
#define PS_TAP_TILE_OFFSETS 3  /* int32 [V*n_tiles + 1] first sorted index of every (view, tile)  */
#define PS_TAP_LAST_IDS 4      /* int32 [V,H,W] 1 + index of last contributing entry              */
#define PS_TAP_TILES_TOUCHED 5 /* int32 [V*N]                                                    */
#define PS_TAP_REC0 6          /* float4 [V*N] 3D: x,y,log(255*opacity),opacity  2D: u,v,log(opacity/tau),opacity */
#define PS_TAP_REC1 7          /* float4 [V*N] 3D: A/2,B,C/2,0   2D: cos,sin,iax,iay             */
#define PS_TAP_REC2 8          /* float4 [V*N] 3D: r,g,b,0       2D: r,g,b,0                     */
#define PS_TAP_DEPTH 9         /* uint32 [V*N] 3D: bits of the camera-space depth (key low word)  */

typedef struct ps_ctx ps_ctx;     /* per-device context (stream-ordered scratch pool, pinned mailbox slots); calls on one
                                     context may come from several host threads: every forward owns its mailbox slot and
                                     its ps_saved, shared bookkeeping is locked                                          */
typedef struct ps_saved ps_saved; /* state of one forward kept for its backward / taps               */

typedef struct ps_render_desc {
    int32_t mode;      /* PS_MODE_2D | PS_MODE_3D                                            */
    int32_t width;     /* image width  (renderer.width,  src/gaussian_renderer.py:47)        */
    int32_t height;    /* image height (renderer.height, :48)                                 */
    int32_t n_frames;  /* F: parameter sets                                                   */
    int32_t n_gauss;   /* N: Gaussians per frame (rows of gaussian_params)                    */
    int32_t n_views;   /* V: (frame, camera) views rendered by this call                      */
    int32_t flags;     /* PS_FLAG_*                                                           */
    float near_plane;  /* 3D, gsplat default 0.01  (src/model.py:351)                         */
    float far_plane;   /* 3D, gsplat default 1e10  (src/model.py:352)                         */
    float radius_clip; /* 3D, 0.0 at :196-208, 2.0 in the legacy splat (src/model.py:339)     */
    float eps2d;       /* 3D, gsplat default 0.3                                              */
} ps_render_desc;

typedef struct ps_saved_info {
    int64_t n_isect;   /* M = sum of tiles touched                                            */
    int32_t tile_bits; /* floor(log2(n_tiles)) + 1                                            */
    int32_t view_bits;
    int32_t tiles_x, tiles_y;
    int32_t n_views, n_gauss, n_frames, mode, width, height;
    int32_t n_lists;   /* non-empty (view, tile) lists                                        */
    int32_t reserved;
} ps_saved_info;

int ps_abi_version(void);
const char *ps_last_error(void);

/* Context for one CUDA device. */
int ps_ctx_create(int device, ps_ctx **out);
int ps_ctx_destroy(ps_ctx *ctx);

/*
 * Forward render of V views.
 * Replaces: GaussianRenderer3D.render (src/gaussian_renderer.py:157-211, i.e. the adapter
 * activations :183-193 + gsplat rasterization :196-208) and GaussianRenderer2D.render
 * (:269-334 incl. _render_vectorized :336-427).  One reference call is V = F = 1.
 *   params      [F, N, 14|9]  raw gaussian_params rows (activations are applied inside)
 *   view_frame  [V] int32     which parameter set each view renders
 *   viewmats    [V, 4, 4]     world->camera, row-major (3D; may be NULL in 2D)
 *   Ks          [V, 3, 3]     intrinsics, row-major      (3D; may be NULL in 2D)
 *   background  [3]           renderer.background_color (:53-56)
 * Outputs (caller-allocated): rgb [V,H,W,3], alpha [V,H,W], n_contrib [V,H,W] int32 or NULL.
 * *saved receives the state for ps_backward when PS_FLAG_SAVE_FOR_BACKWARD is set (else NULL).
 */
int ps_forward(ps_ctx *ctx, const ps_render_desc *desc, const float *params, const int32_t *view_frame,
               const float *viewmats, const float *Ks, const float *background, float *rgb, float *alpha,
               int32_t *n_contrib, ps_saved **saved, void *stream);

/*
 * Inference forward that writes uint8 RGBA [V,H,W,4] directly: replaces render + torch.cat([rgb, alpha], -1) +
 * (255 * clip(x, 0, 1)).astype(uint8) of the evaluation writer (scripts/utils/evaluate_model.py:101-113), so a
 * pixel leaves the GPU as 4 bytes instead of 16.  Same arguments as ps_forward; no state is saved.
 */
int ps_forward_rgba8(ps_ctx *ctx, const ps_render_desc *desc, const float *params, const int32_t *view_frame,
                     const float *viewmats, const float *Ks, const float *background, uint8_t *rgba8, void *stream);

/*
 * Backward of ps_forward: d_params [F,N,P] = dL/d gaussian_params given d_rgb [V,H,W,3] and
 * d_alpha [V,H,W].  Replaces autograd through :183-211 / :314-427 (gsplat's
 * rasterize_to_pixels_bwd + fully_fused_projection_bwd in 3D).  params / viewmats / Ks must hold what the forward
 * was given; `background` is ignored (the forward kept its own copy; may be NULL).  d_params is overwritten (not
 * accumulated into).  `stream` may differ from the forward's: the call is then ordered (one event) behind the work
 * already queued on the saved buffers; the same holds for ps_saved_copy and ps_saved_release, so a block released on one
 * stream never reaches a new owner while another stream still reads it.  Ordering the caller's own tensors (params,
 * cotangents, outputs) across streams stays the caller's job.
 */
int ps_backward(ps_ctx *ctx, ps_saved *saved, const float *params, const int32_t *view_frame, const float *viewmats,
                const float *Ks, const float *background, const float *d_rgb, const float *d_alpha, float *d_params,
                void *stream);

/*
 * Backward with the cross-GPU gradient exchange fused in (no reference counterpart: the reference is
 * single-GPU).  For use when the cameras of one frame are rendered by different ranks.  Frame f is owned by rank
 * f % world.  Instead of writing its partial d_params and all-reducing it, the projection-backward kernel pushes
 * every finished block of rows, with coalesced stores, straight into slot [my_rank][f / world] of the owner's
 * staging buffer -- local memory or a peer GPU's over NVLink.
 *   stage_ranks  DEVICE array [world] of device pointers: rank r's staging buffer [world][ceil(F/world)][N][P] fp32,
 *                each mapped into this process (e.g. torch symmetric memory); entry [my_rank] is local
 * Protocol (host side): all ranks meet at a barrier (the owners are done with the previous step's staging), every
 * rank calls ps_backward_peer, all ranks meet at a barrier again, then each rank calls ps_peer_sum on its own
 * staging buffer: out [ceil(F/world)][N][P] = the complete gradient of the frames it owns (f = my_rank + world*k).
 */
int ps_backward_peer(ps_ctx *ctx, ps_saved *saved, const float *params, const float *viewmats, const float *Ks,
                     const float *background, const float *d_rgb, const float *d_alpha, float *const *stage_ranks,
                     int my_rank, int world, void *stream);
int ps_peer_sum(ps_ctx *ctx, const float *stage_local, int world, size_t n_per_slot, float *out, void *stream);

int ps_saved_info_get(const ps_saved *saved, ps_saved_info *out);
/* Copy one tap (PS_TAP_*) into dst (device or host pointer), at most `bytes`; waits on the stream. */
int ps_saved_copy(ps_ctx *ctx, const ps_saved *saved, int what, void *dst, size_t bytes, void *stream);
int ps_saved_release(ps_ctx *ctx, ps_saved *saved, void *stream);

/* Number of kernels launched by this context so far (bench.py's gpu_launches). */
int64_t ps_ctx_launch_count(const ps_ctx *ctx);

/*
 * Measurement hooks (bench.py).  With profiling on, every stage is bracketed by CUDA events on
 * the launching stream; ps_ctx_stage_times synchronises, adds the finished intervals to
 * ms[PS_N_STAGES] / calls[PS_N_STAGES] (totals since the last reset) and optionally resets.
 */
int ps_ctx_set_profiling(ps_ctx *ctx, int on);
int ps_ctx_stage_times(ps_ctx *ctx, double *ms, int64_t *calls, int reset);
/* Pair counters of the kernels that ran with PS_FLAG_RASTER_STATS (the SAME kernels that are timed, with the counters
 * compiled in; a backward counts when its forward had the flag).  pairs[8]:
 *   [0] forward: (pixel, entry) pairs whose sigma / q was evaluated   [4] backward: the same for the replay
 *   [1] forward: contributing pairs                                   [5] backward: contributing pairs
 *   [2] forward: block-list entries that reach a live pixel           [6] backward: entries that pass the chunk cull
 *   [3] forward: block-list entries staged (chunks x 32)              [7] backward: entries staged               */
int ps_ctx_raster_stats(ps_ctx *ctx, uint64_t *pairs, int reset, void *stream);
/* FP32 FFMA micro-benchmark on this device: returns achieved TFLOP/s (2 flops per FMA) in *tflops. */
int ps_fp32_peak_probe(ps_ctx *ctx, double *tflops, void *stream);

/*
 * The per-view training loss that follows render() in every training step, fused with its own backward
 * (SURVEY.md 8f-f1).  Replaces scripts/training/train_script.py:30-36 (get_iou_loss) and :129-133 of the reference:
 *   iou  = 1 - (sum(alpha*mask) + 1e-6) / (sum(alpha + mask - alpha*mask) + 1e-6)
 *   ssim = ssim_lambda * (1 - SSIM(target_img, rgb))      torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0)
 *   img  = img_lambda * sum|target_img - rgb| / sum(mask)
 * for V views at once, each view its own loss (the reference runs one view per step, batch size 1).
 *   rgb [V,H,W,3], alpha [V,H,W]          what ps_forward wrote
 *   target_img [V,3,H,W], target_mask [V,H,W]   as the reference's loader yields them (img[0, img_idx], mask[0, img_idx])
 *   losses [V,3]                          (iou, ssim, img) per view
 *   d_rgb [V,H,W,3], d_alpha [V,H,W]      gradient of sum_v (iou + ssim + img)_v: the cotangents ps_backward takes;
 *                                         both NULL = losses only (validation, :39-66).  H, W >= 11.
 */
int ps_view_loss(ps_ctx *ctx, int n_views, int height, int width, const float *rgb, const float *alpha,
                 const float *target_img, const float *target_mask, float ssim_lambda, float img_lambda,
                 float *losses, float *d_rgb, float *d_alpha, void *stream);

/*
 * The soft-IoU term alone -- get_iou_loss(predicted_mask, target_mask, eps=1e-6) of scripts/training/train_script.py:30-36,
 * also used on its own for validation (:39-66): losses [V] = 1 - (sum(a m) + 1e-6) / (sum(a + m - a m) + 1e-6) and, if
 * d_alpha is not NULL, its gradient [V,H,W].  Any image size; an all-zero target mask gives a finite value (the eps).
 */
int ps_iou_loss(ps_ctx *ctx, int n_views, int height, int width, const float *alpha, const float *target_mask, float *losses,
                float *d_alpha, void *stream);

/*
 * The parameter-head tail that produces the rows ps_forward takes, one thread per Gaussian.
 * Replaces src/model.py:207-257 (activations of get_gaussian_params_from_volume_unified after the MLP) and, with
 * pose != 0 in 3D mode, apply_pose_transform_3d :261-298 incl. quaternion_matrix_torch_batch :368-391 and
 * quaternion_from_matrix_torch_batch :394-421 (a float64 torch.linalg.eigh per Gaussian, here a register-resident Jacobi).
 *   net_out   [n,14]  3D: quats 4 | scales 3 | opacity 1 (unused) | colours 3 | delta_means 3   (the MLP output, :210-212)
 *             [n, 9]  2D: means_2d 2 | scales_2d 2 | rotation 1 | colours 3 | opacity 1 (unused)         (:236-238)
 *   probs_sel [n]     probs[mask]: sigmoid(volume[0] - mask_threshold) of the selected voxels            (:186-187)
 *   grid_sel  [n,3]   self.grid.view(-1,3)[mask] (3D only)                                               (:223)
 *   scale0    [1]     DEVICE pointer to the trainable self.scale                                         (:86,219)
 *   voxel_size, prob_threshold, clip_lo/hi = self.voxel_size, self.prob_threshold, self.color_clip
 *   pose, angle, p_3d_host[3] (HOST pointer)  yaw and translation of the frame (:275-280); or, for the rows of several
 *             frames in one launch, poses [n_frames,5] = (cos, sin, px, py, pz) per frame and row_frame [n] int32 (DEVICE;
 *             both NULL = the scalar pose); a row whose frame id is outside [0, n_frames) comes out as NaN
 *   rows      [n,14|9] gaussian_params as render() takes them
 * Backward: d_net_out [n,14|9], d_probs_sel [n], d_scale0 [1] (device, overwritten) from d_rows.
 */
int ps_param_head_forward(ps_ctx *ctx, int mode, int n, const float *net_out, const float *probs_sel, const float *grid_sel,
                          const float *scale0, float voxel_size, float prob_threshold, float clip_lo, float clip_hi,
                          int pose, double angle, const float *p_3d_host, const float *poses, const int32_t *row_frame,
                          int n_frames, float *rows, void *stream);
int ps_param_head_backward(ps_ctx *ctx, int mode, int n, const float *net_out, const float *probs_sel, float voxel_size,
                           float prob_threshold, float clip_lo, float clip_hi, int pose, double angle, const float *poses,
                           const int32_t *row_frame, int n_frames, const float *d_rows, float *d_net_out, float *d_probs_sel, float *d_scale0,
                           void *stream);

/*
 * Device probe of the arithmetic contract (PSM-1): y[5][n] = exp, log(|x|+1e-30), sigmoid, sin, cos
 * of x[n], computed by the same device functions the kernels use.  For the bit-exactness tests.
 */
int ps_math_probe(ps_ctx *ctx, const float *x, int n, float *y, void *stream);

/*
 * Device probe of the 3D adapter stage (src/gaussian_renderer.py:183-193), through the device functions the projection
 * kernels call: act [n,14] = means | exp(log_scales) | q / (|q| + 1e-8) | clamp(colours, 0, 1) | sigmoid(logit), i.e. the
 * tensors the reference hands to gsplat.rendering.rasterization (:196-208); with v_act [n,14] also d_rows [n,14] = J^T v_act,
 * what autograd propagates through those lines.  For the fixture test against the reference's own code
 * (tests/golden/adapter3d_reference.npz).  v_act and d_rows may both be NULL.
 */
int ps_adapter3d_probe(ps_ctx *ctx, const float *rows, int n, const float *v_act, float *act, float *d_rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PSPLAT_H */
