/*
 * gi2d.h -- C ABI of libgi2d.so: a B200 (sm_100a) native 2-D Gaussian image rasterizer.
 *
 * This is the drop-in boundary for the hot path of Sweethyh/GaussianImage_plus:
 *   project -> bin/sort -> rasterize-sum forward -> rasterize-sum backward -> project backward (+Adam)
 *
 * Every entry point replaces one function of the reference's pybind11 module
 * (`gsplat/gsplat/cuda/csrc/ext.cpp:4-69`, implemented in `csrc/bindings.cu`) or one
 * third-party PyTorch call on the path (`gsplat/gsplat/utils.py:248,301,302`).  The
 * reference binding each one replaces is cited next to its declaration.
 *
 * Conventions
 *   - plain pointers and sizes only: all pointers are DEVICE pointers unless the name ends
 *     in `_host`; no torch types; outputs and workspaces are caller-owned.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, never allocates; safe to capture in a CUDA graph.
 *   - return value: 0 on success, a negative GI2D_ERR_* otherwise; gi2d_last_error()
 *     returns a thread-local message.  Launch errors are checked (the reference never does).
 *   - array layouts are the reference's: xys f32[N,2], conics f32[N,3] = (a,b,c),
 *     colors f32[N,3], images f32[H,W,3] interleaved, tile_bins i32[rows,2],
 *     tiles are 16x16 pixels (`csrc/config.h:1-4`).
 */
#ifndef GI2D_H_
#define GI2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GI2D_ABI_VERSION 1
#define GI2D_TILE 16 /* BLOCK_X == BLOCK_Y, csrc/config.h:1-2 */
#define GI2D_MAX_PER_TILE 256 /* the live rasterizer stops after one 256-batch, csrc/forward.cu:673 */

#define GI2D_OK 0
#define GI2D_ERR_INVALID (-1)   /* bad argument (null pointer, negative size, unsupported value) */
#define GI2D_ERR_CUDA (-2)      /* a CUDA launch or runtime call failed */
#define GI2D_ERR_WORKSPACE (-3) /* workspace too small; see the *_workspace_size query */

typedef void *gi2d_stream_t; /* cudaStream_t */

int gi2d_abi_version(void);
const char *gi2d_last_error(void);

/* ------------------------------------------------------------------------------------------
 * R1-R3  projection forward (one thread per Gaussian)
 * ------------------------------------------------------------------------------------------ */

/* Replaces `project_gaussians_2d_covariance_forward` (ext.cpp:38, bindings.cu:1455-1513,
 * kernel foward2d.cu:192-288).  means2d f32[N,2] in pixels, cov2d f32[N,3]=(sxx,sxy,syy).
 * Outputs are fully written (culled Gaussians get radii=0, num_tiles_hit=0, depths=0 and
 * zeroed xys/conics exactly like the reference's torch::zeros + early return). */
int gi2d_project_cov_fwd(int num_points, const float *means2d, const float *cov2d,
                         int img_width, int img_height, int tiles_x, int tiles_y,
                         float clip_coe, float radius_clip,
                         float *xys, float *depths, int32_t *radii, float *conics,
                         int32_t *num_tiles_hit, gi2d_stream_t stream);

/* Replaces `project_gaussians_2d_forward` (ext.cpp:31, bindings.cu:1317-1381, kernel
 * foward2d.cu:12-69).  means2d in [-1,1], L f32[N,3]=(l11,l21,l22). */
int gi2d_project_chol_fwd(int num_points, const float *means2d, const float *L_elements,
                          int img_width, int img_height, int tiles_x, int tiles_y,
                          float clip_coe, float radius_clip,
                          float *xys, float *depths, int32_t *radii, float *conics,
                          int32_t *num_tiles_hit, gi2d_stream_t stream);

/* Replaces `project_gaussians_2d_scale_rot_forward` (ext.cpp:33, bindings.cu:1384-1448,
 * kernel foward2d.cu:130-187).  scales f32[N,2], rotation f32[N]. */
int gi2d_project_rs_fwd(int num_points, const float *means2d, const float *scales2d,
                        const float *rotation, int img_width, int img_height, int tiles_x,
                        int tiles_y, float clip_coe, float radius_clip,
                        float *xys, float *depths, int32_t *radii, float *conics,
                        int32_t *num_tiles_hit, gi2d_stream_t stream);

/* Replaces `compute_cov2d_bounds` (ext.cpp:58, bindings.cu:21-39,41-67).
 * radii_f32 f32[N] = radius.x (the reference returns the float major radius). */
int gi2d_compute_cov2d_bounds(int num_points, float clip_coe, const float *cov2d,
                              float *conics, float *radii_f32, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * R7  projection backward.  All outputs fully written (zeros where radii<=0).
 * ------------------------------------------------------------------------------------------ */

/* Replaces `project_gaussians_2d_covariance_backward` (ext.cpp:39, bindings.cu:1565-1612,
 * kernel backward2d.cu:157-214). */
int gi2d_project_cov_bwd(int num_points, const int32_t *radii, const float *conics,
                         const float *v_xy, const float *v_conic,
                         float *v_cov2d, float *v_mean2d, float *v_L, gi2d_stream_t stream);

/* Replaces `project_gaussians_2d_backward` (ext.cpp:32, bindings.cu:1516-1562,
 * kernel backward2d.cu:8-51) -- including the doubled off-diagonal term. */
int gi2d_project_chol_bwd(int num_points, const float *L_elements, int img_width,
                          int img_height, const int32_t *radii, const float *conics,
                          const float *v_xy, const float *v_conic,
                          float *v_cov2d, float *v_mean2d, float *v_L, gi2d_stream_t stream);

/* Replaces `project_gaussians_2d_scale_rot_backward` (ext.cpp:34, bindings.cu:1615-1668,
 * kernel backward2d.cu:53-101). */
int gi2d_project_rs_bwd(int num_points, const float *scales2d, const float *rotation,
                        const int32_t *radii, const float *conics, const float *v_xy,
                        const float *v_conic, float *v_cov2d, float *v_mean2d,
                        float *v_scale, float *v_rot, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * R4  binning -- integer exact
 * ------------------------------------------------------------------------------------------ */

/* Replaces `torch.cumsum(num_tiles_hit, dtype=int32)` (gsplat/utils.py:248): inclusive scan.
 * `total` (nullable) receives cum[N-1] on the device.  Workspace: gi2d_scan_workspace_size. */
size_t gi2d_scan_workspace_size(int num_points);
int gi2d_cumsum_i32(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit,
                    int32_t *total, void *workspace, size_t workspace_bytes,
                    gi2d_stream_t stream);

/* Replaces `map_gaussian_to_intersects` (ext.cpp:60, bindings.cu:283-365, kernel
 * forward.cu:141-206): isect_ids[k] = (tile_id<<32) | sign-extended bits(depth). */
int gi2d_map_gaussian_to_intersects(int num_points, const float *xys, const float *depths,
                                    const int32_t *radii, const int32_t *cum_tiles_hit,
                                    int tiles_x, int tiles_y, float radius_clip,
                                    int64_t *isect_ids, int32_t *gaussian_ids,
                                    gi2d_stream_t stream);

/* Replaces `torch.sort(isect_ids)` + `torch.gather(gaussian_ids, perm)` (gsplat/utils.py:
 * 301-302): stable LSD radix sort of signed 64-bit keys carrying 32-bit values.  Only bits
 * [begin_bit,end_bit) are inspected (0,64 = full signed order). */
size_t gi2d_sort_workspace_size(int num_items);
int gi2d_sort_pairs_i64(int num_items, const int64_t *keys_in, const int32_t *vals_in,
                        int64_t *keys_out, int32_t *vals_out, int begin_bit, int end_bit,
                        void *workspace, size_t workspace_bytes, gi2d_stream_t stream);

/* Replaces the whole of `compute_cumulative_intersects` + `bin_and_sort_gaussians` (gsplat/utils.py:231-311:
 * torch.cumsum, the `.item()` host synchronisation, map_gaussian_to_intersects, torch.sort, torch.gather,
 * get_tile_bin_edges) by ONE call that keeps num_intersects ON THE DEVICE: per-tile overlap counts,
 * prefix sum, counting-sort placement, in-tile rank sort by Gaussian id.  Contract: every depth has the same bit
 * pattern (the 2-D projections emit 0.0), so the reference's stable sort of Gaussian-major keys orders by tile, then
 * by Gaussian id.  Outputs have `capacity` rows (tile_bins: tiles_x*tiles_y rows, empty tiles (0,0));
 * info i32[3] (device) = {rows written = min(num_intersects, capacity), num_intersects, overflow flag}.
 * Bit-identical to the multi-call path on rows [0, num_intersects). */
size_t gi2d_bin_sort_workspace_size(int num_points, int tiles_x, int tiles_y, int capacity);
int gi2d_bin_sort(int num_points, const float *xys, const float *depths, const int32_t *radii,
                  int tiles_x, int tiles_y, float radius_clip, int capacity,
                  int64_t *isect_ids_sorted, int32_t *gaussian_ids_sorted, int32_t *tile_bins,
                  int32_t *info, void *workspace, size_t workspace_bytes, gi2d_stream_t stream);

/* Replaces `get_tile_bin_edges` (ext.cpp:66, bindings.cu:368-383, kernel forward.cu:211-233).
 * tile_bins i32[num_bins_rows,2] is zero-filled first (the reference's torch::zeros), rows
 * whose tile id is >= num_bins_rows are skipped instead of written out of bounds. */
int gi2d_get_tile_bin_edges(int num_intersects, const int64_t *isect_ids_sorted,
                            int32_t *tile_bins, int num_bins_rows, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * R5/R6  rasterize-sum forward / backward (3 channels, 16x16 tiles, first 256 per tile)
 * ------------------------------------------------------------------------------------------ */

/* Replaces `rasterize_sum_plus_forward` / `rasterize_sum_forward` (ext.cpp:23,16,
 * bindings.cu:529-610, kernel forward.cu:570-691).  out_img f32[H,W,3], final_Ts f32[H,W]
 * (== 1), final_idx i32[H,W]; all fully written.  tile_bins must have >= tiles_x*tiles_y rows
 * OR num_bins_rows gives the number of valid rows (tiles beyond it are treated as empty). */
int gi2d_rasterize_sum_fwd(int tiles_x, int tiles_y, int img_width, int img_height,
                           const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                           int num_bins_rows, const float *xys, const float *conics,
                           const float *colors, const float *opacities,
                           float *out_img, float *final_Ts, int32_t *final_idx,
                           gi2d_stream_t stream);

/* The same with num_intersects on the DEVICE (info[1] of gi2d_bin_sort): when it is < 1 the kernel writes
 * `background` (f32[3], nullable = ones) to every pixel -- the branch rasterize_sum_plus.py:110-118 takes on the
 * host after its `.item()`. */
int gi2d_rasterize_sum_fwd_dev(int tiles_x, int tiles_y, int img_width, int img_height,
                               const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                               int num_bins_rows, const float *xys, const float *conics,
                               const float *colors, const float *opacities,
                               float *out_img, float *final_Ts, int32_t *final_idx,
                               const int32_t *num_intersects_dev, const float *background,
                               gi2d_stream_t stream);

/* Replaces `rasterize_sum_plus_backward` / `rasterize_sum_backward` (ext.cpp:24,17,
 * bindings.cu:1241-1314, kernel backward.cu:1168-1350).  v_xy f32[N,2], v_conic f32[N,3],
 * v_colors f32[N,3], v_opacity f32[N] are zero-filled by the call (ONE fill when the four lie back to back in one
 * allocation, as the binding allocates them), then accumulated. */
int gi2d_rasterize_sum_bwd(int num_points, int tiles_x, int tiles_y, int img_width,
                           int img_height, const int32_t *gaussian_ids_sorted,
                           const int32_t *tile_bins, int num_bins_rows, const float *xys,
                           const float *conics, const float *colors, const float *opacities,
                           const int32_t *final_idx, const float *v_output,
                           float *v_xy, float *v_conic, float *v_colors, float *v_opacity,
                           gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused fit step: the whole `train_iter` of
 * models/gaussianimage_covariance.py:249-259 for the covariance model, with no host synchronisation, in
 * 2 launches (bucketed binning, see gi2d_fit_bucket_capacity below): [project backward + Adam of the previous
 * step] + project + placement of the 64-bit (tile|gaussian) keys and 32-B records into per-tile buckets ->
 * in-tile key sort + rasterize forward + loss gradient (mse / l1 inline; SSIM / MS-SSIM through more kernels) +
 * rasterize backward.  (Scan + placement path -- tile-row split, tiny capacities: per-tile overlap counts ->
 * prefix sum + placement -> the same rasterizer: 3 launches up to 2048 tiles, 6 beyond.)
 * ------------------------------------------------------------------------------------------ */

typedef struct gi2d_fit_params {
    int32_t num_points;
    int32_t img_width, img_height;
    int32_t tiles_x, tiles_y;   /* full image tile grid */
    int32_t tile_row_begin;     /* this rank renders tile rows [begin,end) (multi-GPU split) */
    int32_t tile_row_end;
    int32_t isect_capacity;     /* rows of sorted_keys (and of the workspace copy) */
    float clip_coe, radius_clip;
    float lr0, beta1, beta2, eps; /* Adam (torch.optim.Adam, eps=1e-15 in the reference) */
    int32_t lr_step_size;       /* StepLR(step_size, gamma): lr = lr0 * gamma^floor((step-1)/size) */
    float lr_gamma;
    int32_t color_sigmoid;      /* 1: colours = sigmoid(features) (the reference's color_norm) */
    float loss_scale;           /* dL/d(out) = loss_scale * (clamp(out) - gt); 2/(3*H*W) for mse */
    int32_t external_optimizer; /* 1: a training step leaves b->grads alone (nothing pending): the caller reads
                                   them with gi2d_fit_input_grads and runs its own optimiser; 2: tile-row split
                                   (gi2d_tilerow_step: the exchange kernel owns optimiser and step counters) */
    /* loss_fn of models/utils.py:60-80 as three weights:  loss = w2 * mse + w1 * l1 + ws * (1 - ssim)
     *   L2 (1,0,0)  L1 (0,1,0)  SSIM (0,0,1)  Fusion1 (l,0,1-l)  Fusion2 (0,l,1-l)  Fusion3 (l,1-l,0),  l = 0.7
     * loss_scale = 2 w2 / (3HW) (above), loss_l1_scale = w1 / (3HW), loss_ssim_weight = ws.  With ws == 0 the
     * loss gradient is evaluated inside the rasterize launch; with ws != 0 the launch is split in forward /
     * SSIM gradient (2 kernels, gi2d_loss.cu) / backward. */
    float loss_l1_scale;
    float loss_ssim_weight;
    int32_t dynamic_points;     /* 1: the per-Gaussian arrays hold num_points ROWS (a capacity) and the live count is
                                   stats[GI2D_STAT_NUM_POINTS]: gi2d_fit_prune / gi2d_fit_densify change the model's
                                   size on the device, with no reallocation, no host round trip and no new graph */
    float loss_msssim_weight;   /* wm: loss += wm * (1 - ms_ssim) -- Fusion4 (0, l, 0; wm = 1-l, win 11) and Fusion_hinerv
                                   (win 5) of models/utils.py:76-79; exclusive with loss_ssim_weight; the rasterize
                                   launch is split around gi2d_image_msssim_loss_grad's kernels */
    int32_t loss_msssim_win;    /* 11 or 5 */
} gi2d_fit_params;

/* stats layout (f64): the device-side step counter makes the step graph-replayable with no
 * per-iteration host input. */
#define GI2D_STAT_STEP 0        /* number of Adam steps taken so far */
#define GI2D_STAT_ISECTS 1      /* num_intersects of the last forward */
#define GI2D_STAT_OVERFLOW 2    /* != 0: num_intersects exceeded isect_capacity (step skipped) */
#define GI2D_STAT_LR 3          /* lr used by the last Adam step */
#define GI2D_STAT_BEST_SSE 9    /* smallest squared error of any training step so far (+inf before the first) */
#define GI2D_STAT_BEST_STEP 10  /* the step (1-based) that achieved it; b->best holds the parameters AFTER its update */
#define GI2D_STAT_NON_PSD 11    /* Gaussians whose covariance is not positive definite, counted by gi2d_fit_adam */
#define GI2D_STAT_SSIM_SUM 13   /* sum of the SSIM map over channels and valid windows of the last step (ws != 0) */
#define GI2D_STAT_ABS_SUM 14    /* sum |clamp(out)-gt| of the last step (w1 != 0) */
#define GI2D_STAT_SSE 16        /* 64 partial sums of squared error of the clamped render */
#define GI2D_STAT_SSE_SLOTS 64
#define GI2D_STAT_NUM_POINTS 80 /* the Gaussian count when p->dynamic_points: prune / densify change it ON THE DEVICE */
#define GI2D_STAT_BEST_N 81     /* Gaussian count of the best-state snapshot (rows of b->best / b->best_bound) */
#define GI2D_STAT_PRUNED 82     /* Gaussians removed by the last gi2d_fit_prune */
#define GI2D_STAT_ADDED 83      /* Gaussians appended by the last gi2d_fit_densify */
#define GI2D_STAT_MAX_TILE 84   /* != 0: the largest per-tile overlap count of a forward whose bucketed binning
                                   overflowed (more than isect_capacity / #tiles overlaps in one tile) */
#define GI2D_STAT_MSSSIM 89     /* ms_ssim of the last training step (loss_msssim_weight != 0) */
#define GI2D_STAT_COUNT 96

typedef struct gi2d_fit_buffers {
    /* parameters + optimiser state (updated in place) */
    float *xyz;        /* f32[N,2] pixel coords          (_xyz) */
    float *cov;        /* f32[N,3] raw covariance params (_cov2d) */
    float *cov_bound;  /* f32[N,3] added to cov          (cholesky_bound) */
    float *rgb;        /* f32[N,3]                       (_features_dc) */
    float *m_xyz, *v_xyz, *m_cov, *v_cov, *m_rgb, *v_rgb; /* Adam exp_avg / exp_avg_sq */
    /* target and outputs */
    const float *gt_hwc; /* f32[H,W,3] target image, interleaved (nullable for render-only) */
    float *out_img;      /* nullable.  fit: f32[H,W,3] unclamped render (== rasterize output);
                            render-only (with_backward=0): f32[3,H,W] clamped to [0,1], i.e. the
                            model's `render` tensor (gaussianimage_covariance.py:210-211) */
    /* per-step scratch, caller-owned */
    float *grads;          /* f32[N,8] = v_xy(2) v_conic(3) v_rgb(3); atomically accumulated */
    float *proj;           /* f32[N,8] = x, y, conic a,b,c, colour r,g,b */
    uint64_t *sorted_keys; /* u64[capacity] (tile<<32 | gaussian), ascending */
    int32_t *tile_bins;    /* i32[tiles_x*tiles_y, 2] */
    double *stats;         /* f64[GI2D_STAT_COUNT] */
    void *workspace;
    size_t workspace_bytes;
    const uint8_t *gt_u8_hwc; /* u8[H,W,3] target as stored (PNG bytes); used when gt_hwc is NULL: the
                                 kernel computes u8/255 exactly like torchvision's ToTensor (utils.py:21-27) */
    float *best;    /* nullable f32[N,8] = xyz(2) cov(3) rgb(3): best-state snapshot.  train.py:132-137 deep-copies
                       the state dict whenever the step's PSNR beats the best so far; here the kernel that applies
                       a step's Adam update also stores the updated parameters when that step's squared error was
                       a new minimum (GI2D_STAT_BEST_SSE / _STEP) -- no host round trip, no extra launch */
    float *err_map; /* nullable f32[H,W]: a training step also writes the per-pixel L1 error of its clamped
                       render, sum_c |clamp(out)-gt| (the `errors` map of train.py:87 that drives densification) */
    float *best_bound; /* nullable f32[N,3]: cov_bound rows of the best-state snapshot (the `slv_bound` copy of
                          train.py:136), written together with b->best */
} gi2d_fit_buffers;

size_t gi2d_fit_workspace_size(const gi2d_fit_params *p);

/* Number of kernels one gi2d_fit_forward_backward (+ gi2d_fit_adam when with_backward) launches. */
int gi2d_fit_launch_count(const gi2d_fit_params *p, int with_backward);

/* Zero the stats block (and set the step counter): call once before the first step. */
int gi2d_fit_reset(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int step, gi2d_stream_t stream);

/* One step: [Adam of the previous step, if its gradient is still pending] + project + bin + rasterize.
 * with_backward != 0 also computes the L2 loss gradient and the rasterize backward into b->grads,
 * advances the step counter and leaves that gradient PENDING: projection backward + Adam are folded
 * into the first kernel of the next call (same per-Gaussian thread, no extra launch).  A multi-GPU
 * caller all-reduces b->grads between two calls.  Call gi2d_fit_adam to apply a pending gradient now. */
int gi2d_fit_forward_backward(const gi2d_fit_params *p, const gi2d_fit_buffers *b,
                              int with_backward, gi2d_stream_t stream);
/* Flush: projection backward + Adam on xyz/cov/rgb from a pending b->grads (no update when nothing is
 * pending).  Needed before the host reads or edits parameters, and after the last step.  Also counts the
 * Gaussians whose covariance cov+cov_bound is not positive definite into GI2D_STAT_NON_PSD (the test of
 * gaussianimage_covariance.py:373-382), so the prune decision costs no extra launch. */
int gi2d_fit_adam(const gi2d_fit_params *p, const gi2d_fit_buffers *b, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Changing the model's size on the device (p->dynamic_points).
 * ------------------------------------------------------------------------------------------ */

/* Set the live Gaussian count (rows [0,n) of the per-Gaussian arrays are the model). */
int gi2d_fit_set_num_points(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int n, gi2d_stream_t stream);

/* `non_semi_definite_prune` (models/gaussianimage_covariance.py:354-382): flush a pending optimiser step, then
 * drop every Gaussian whose covariance cov + cov_bound is not positive definite (det > 0, sxx > 0, syy > 0 in
 * torch's operation order), keeping the order of the survivors: parameters, both Adam moments and the bound rows
 * are compacted together (stable, through a scratch copy), the live count and GI2D_STAT_PRUNED are updated.
 * Nothing moves when no row fails or when every row would fail (`if to_prune_nums and cur - to_prune > 0`).
 * Asynchronous, no host round trip.  workspace: gi2d_fit_prune_workspace_size(p->num_points) bytes. */
size_t gi2d_fit_prune_workspace_size(int capacity);
int gi2d_fit_prune(const gi2d_fit_params *p, const gi2d_fit_buffers *b, void *workspace, size_t workspace_bytes,
                   gi2d_stream_t stream);

/* `add_sample_positions` + `densification_postfix` (train.py:85-118, models/gaussianimage_covariance.py:307-334):
 * the k pixels of largest error in b->err_map (f32[H,W], written by a training step; descending, ties by the
 * smaller pixel index -- a full 64-bit radix sort of (error, index) keys) become new Gaussians at (x, y) with
 * zero colour, zero Adam moments and covariance new_cov2d[i] (f32[k_rows,3], the caller's random draw + (0.5,0,0.5):
 * the reference draws it on the CPU generator); candidates whose covariance is not positive definite are skipped;
 * every appended row gets the bound (lp, 0, lp), lp = min(HW / (9 pi n_new), 300) of the NEW count (slv != 0), and
 * the live count and GI2D_STAT_ADDED are updated.  k = min(k_rows, p->num_points - live count).
 * A pending optimiser step is flushed first.  workspace: gi2d_fit_densify_workspace_size(H, W) bytes. */
size_t gi2d_fit_densify_workspace_size(int img_height, int img_width);
int gi2d_fit_densify(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int k_rows, const float *new_cov2d,
                     int slv, void *workspace, size_t workspace_bytes, gi2d_stream_t stream);

/* Bucketed binning (the default of the fit step except for the tile-row exchange): tile t owns the rows
 * [t * C, (t+1) * C), C = isect_capacity / #tiles, of b->sorted_keys and of the record workspace; the projection
 * kernel places every intersection with one cursor atomic, so a step is TWO launches (no prefix sum, no placement
 * kernel).  A tile with more than C overlaps raises GI2D_STAT_OVERFLOW like a full intersection buffer does
 * (GI2D_STAT_MAX_TILE then holds the largest count).  gi2d_fit_bucket_capacity: C, or 0 when the scan + placement
 * path is in use.  gi2d_fit_export_binning: the reference's view of the last forward's binning -- ONE ascending key
 * array u64[isect_capacity] (tile << 32 | gaussian: isect_ids_sorted >> 32 and gaussian_ids_sorted of
 * utils.py:301-302) and tile_bins i32[#tiles,2] (forward.cu:211-233) -- whichever layout the step used; a
 * diagnostic entry point, it SYNCHRONISES. */
int gi2d_fit_bucket_capacity(const gi2d_fit_params *p);
int gi2d_fit_export_binning(const gi2d_fit_params *p, const gi2d_fit_buffers *b, uint64_t *sorted_keys_out,
                            int32_t *tile_bins_out, gi2d_stream_t stream);

/* Gradients of the last training step with respect to the step's INPUTS, for a caller that keeps its own
 * parameters and optimiser (p->external_optimizer: the quantisation-aware pass feeds de-quantised means /
 * covariances / colours and back-propagates into its quantisers): out f32[N,8] = d loss / d (x, y, sxx, sxy,
 * syy, r, g, b) -- b->grads pushed through the projection backward (-X G X, backward2d.cu:157-214). */
int gi2d_fit_input_grads(const gi2d_fit_params *p, const gi2d_fit_buffers *b, float *out, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Quantisation-aware iteration of the compression pass, default `lsq` quantisers:
 * GaussianImage_Covariance.forward_quantize / train_iter_quantize (models/gaussianimage_covariance.py:219-247,
 * 384-410) = UniformQuantizer (LSQ+, quantize.py:39-156) on positions (2 channels) and colours (3 channels),
 * HybirdQuant (quantize.py:336-389) on the covariance: the two variances through LogQuantizer(learned=False), whose
 * range is the min / max of log(|x|+1e-6) over the whole tensor ON EVERY CALL with gradients through min and max
 * (quantize.py:223-235), the off-diagonal through a 1-channel UniformQuantizer; plus the reference's four
 * torch.optim.Adam + StepLR (attributes: eps 1e-15, StepLR(20000); the three quantisers: lr 1e-3, StepLR(10000),
 * eps 1e-15 except the position quantiser's, which keeps torch's default 1e-8; :116-146).
 * One iteration = gi2d_quant_forward -> a gi2d_fit_forward_backward training step with external_optimizer = 1 on
 * the de-quantised attributes (b->out_* are the fit step's xyz / cov / rgb, cov_bound of the fit step all zero)
 * -> gi2d_fit_input_grads -> gi2d_quant_backward_step; eight launches, no host round trip, graph-capturable.
 * ------------------------------------------------------------------------------------------ */
typedef struct gi2d_quant_params {
    int32_t num_points;
    int32_t xy_qmax, cov_qmax, color_qmax; /* 2^bits - 1 (unsigned ranges: qmin = 0) */
    int32_t color_sigmoid;                 /* 1: colours = sigmoid(features) before quantisation (color_norm) */
    float lr0;                             /* attributes: lr = lr0 * gamma^floor((step-1)/lr_step) */
    int32_t lr_step;
    float lr_q0;                           /* quantiser parameters, same schedule with lr_q_step */
    int32_t lr_q_step;
    float lr_gamma;
    float beta1, beta2;
    float eps;                             /* Adam eps of the attributes, the covariance and the colour quantiser */
    float eps_xyz_q;                       /* Adam eps of the position quantiser */
} gi2d_quant_params;

typedef struct gi2d_quant_buffers {
    float *xyz, *cov, *rgb;                /* f32[N,2] [N,3] [N,3] raw attributes, updated in place */
    const float *bound;                    /* f32[N,3] cholesky_bound (cov + bound is what gets quantised) */
    float *m_xyz, *v_xyz, *m_cov, *v_cov, *m_rgb, *v_rgb; /* Adam moments of the attributes */
    float *qparams;                        /* f32[12] scale_xyz[2] beta_xyz[2] scale_cov beta_cov scale_rgb[3] beta_rgb[3] */
    float *qm, *qv;                        /* f32[12] their Adam moments */
    double *qstats;                        /* f64[32]: [0] iterations done, [1] min [2] max of the log range,
                                              [3] [4] how many elements are the min / the max, [5..18] the
                                              quantiser-parameter gradients of the last iteration (log scale, log
                                              beta, 6 LSQ scales, 6 LSQ betas; channels xyz 0,1 | cov | rgb 0,1,2) */
    float *out_xyz, *out_cov, *out_rgb;    /* de-quantised attributes (the fit step's inputs) */
    const float *in_grads;                 /* f32[N,8] from gi2d_fit_input_grads */
    float *dbg_grads;                      /* optional f32[N,8]: gradient of the raw attributes (tests) */
} gi2d_quant_buffers;

/* UniformQuantizer._init_data / HybirdQuant._init_data (quantize.py:72-80, 352-354): scale / beta of the learned
 * quantisers from the per-channel min / max of the current attributes; zeroes qm, qv and qstats. */
int gi2d_quant_init(const gi2d_quant_params *p, const gi2d_quant_buffers *b, gi2d_stream_t stream);
/* forward_quantize up to the rasterizer's inputs: log range (1 launch) + quantise / de-quantise (1 launch). */
int gi2d_quant_forward(const gi2d_quant_params *p, const gi2d_quant_buffers *b, gi2d_stream_t stream);
/* Straight-through backward of the quantisers + the four optimiser steps (2 launches). */
int gi2d_quant_backward_step(const gi2d_quant_params *p, const gi2d_quant_buffers *b, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU tile-row split of ONE image (the reference is single-GPU, train.py:39).
 * Rank q rasterizes tile rows [band_edge[q], band_edge[q+1]).  Every Gaussian has ONE owner rank (equal
 * contiguous slices [own_begin, own_end)): only the owner holds its parameters and Adam moments, applies
 * projection backward + Adam to it and projects it.  Per step and Gaussian the owner loads the partial
 * gradient rows of the ranks whose band the Gaussian's tile box overlaps (P2P loads, fixed rank order) and
 * stores the new projected record + tile box to the ranks whose band the old or the new box overlaps (P2P
 * stores): about 1.1 x 72 bytes per Gaussian cross NVLink, not (world-1) x 64.  Cross-GPU ordering is two flag
 * words per peer inside the kernels (st.release.sys / ld.acquire.sys on peer-mapped memory): no barrier
 * launches, no host involvement, graph-capturable.  A step in which any rank overflowed its intersection
 * buffers updates nothing anywhere (GI2D_STAT_OVERFLOW is then set on every rank and GI2D_STAT_STEP does not
 * advance).  p->external_optimizer must be 2; b->grads / b->proj must be peer_grads[rank] / peer_proj[rank];
 * p->tile_row_begin/end must be the rank's band.  All peer_* buffers are peer-mapped device memory
 * (symmetric allocations); sync = 0 switches the in-kernel flags off for a caller that orders the ranks
 * itself (several ranks emulated on ONE GPU, stream-ordered: phase 1 for every rank, then phase 2).
 * ------------------------------------------------------------------------------------------ */
#define GI2D_MAX_RANKS 8
typedef struct gi2d_tilerow {
    int32_t rank, world;
    int32_t band_edge[GI2D_MAX_RANKS + 1];
    int32_t own_begin, own_end;
    int32_t sync;
    float *peer_grads[GI2D_MAX_RANKS];     /* f32[N,8] partial gradients of each rank's band */
    float *peer_proj[GI2D_MAX_RANKS];      /* f32[N,8] projected records (every rank holds all of them) */
    void *peer_boxes[GI2D_MAX_RANKS];      /* u16[N,4] tile box (x0,y0,x1,y1) of every Gaussian, unclipped */
    uint32_t *peer_flags[GI2D_MAX_RANKS];  /* u32[16]: [q] "rank q finished the backward of epoch e" (e*2 + overflow),
                                              [8+q] "rank q's exchange of epoch e has landed" */
    uint32_t *ctrl;                        /* u32[8], this rank only: epoch, tickets, [3] != 0: a flag wait timed out */
} gi2d_tilerow;

/* Set the control block (epoch 1) and project + scatter the owned slice.  Before the call every rank has ZEROED
 * its peer-visible buffers and the caller has put a cross-rank barrier (host side); another barrier follows the
 * call (every rank's records have landed everywhere). */
int gi2d_tilerow_init(const gi2d_fit_params *p, const gi2d_fit_buffers *b, const gi2d_tilerow *tr,
                      gi2d_stream_t stream);
/* One step.  phase 1: count + place + rasterize the band (with_backward: + loss gradient + backward into
 * b->grads); phase 2: exchange + Adam + projection of the owned slice; 3: both. */
int gi2d_tilerow_step(const gi2d_fit_params *p, const gi2d_fit_buffers *b, const gi2d_tilerow *tr,
                      int with_backward, int phase, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training steps fed from HOST memory (bench.py's `e2e`): one C call per step does what a driver that gets a
 * new target every step has to do -- upload the target from pinned host memory into the device buffer the
 * step will read (on the pipe's own copy stream, double-buffered by the caller: alternate two
 * gi2d_fit_buffers that differ in their target pointer, and the upload of step k+1 runs under step k), run
 * the step on `stream`, and copy the stats block (f64[GI2D_STAT_COUNT]) to pinned host memory.  Everything
 * is asynchronous; gi2d_host_pipe_wait(slot) blocks until the stats of the call that returned `slot` are on
 * the host.  The pipe owns only a stream and a few events.
 * ------------------------------------------------------------------------------------------ */
typedef struct gi2d_host_pipe gi2d_host_pipe;
int gi2d_host_pipe_create(gi2d_host_pipe **out);
int gi2d_host_pipe_destroy(gi2d_host_pipe *pipe);
int gi2d_fit_step_host(gi2d_host_pipe *pipe, const gi2d_fit_params *p, const gi2d_fit_buffers *b,
                       const void *host_target, size_t target_bytes, double *host_stats,
                       gi2d_stream_t stream, int *slot_out);
int gi2d_host_pipe_wait(gi2d_host_pipe *pipe, int slot);

/* Stand-alone form of the loss gradient the fit step uses when loss_ssim_weight != 0: everything autograd does
 * between the rasterizer output and the loss in gaussianimage_covariance.py:210,252-253 --
 *   v_out = d/d(out) [ ssim_weight * (1 - ssim(clamp(out), gt)) + (l2_scale/2) * sum (clamp(out)-gt)^2
 *                      + l1_scale * sum |clamp(out)-gt| ]
 * with pytorch_msssim.ssim(data_range=1, size_average=True) semantics (11-tap sigma-1.5 window, valid
 * filtering).  render/gt/v_out are f32[H,W,3] (gt alternatively u8); *ssim_sum (nullable, device f64) receives
 * the sum of the SSIM map (mean = sum / (3 (H-10)(W-10))).  workspace: gi2d_ssim_workspace_size bytes. */
size_t gi2d_ssim_workspace_size(int img_height, int img_width);
int gi2d_image_loss_grad(int img_height, int img_width, const float *render_hwc, const float *gt_hwc,
                         const uint8_t *gt_u8_hwc, float ssim_weight, float l2_scale, float l1_scale,
                         float *v_out_hwc, double *ssim_sum, void *workspace, size_t workspace_bytes,
                         gi2d_stream_t stream);

/* MS-SSIM, the evaluation metric of train.py:190 (pytorch_msssim.ms_ssim(data_range=1): 5 levels, 2x2 average
 * pooling with padding size%2 between them).  level_sums (device, f64[5][3][2]) receives per level and channel the
 * sums of the SSIM map and of the contrast-structure map over the level's valid windows; the caller forms
 * prod_l relu(mean)^w_l (cs for levels 0..3, ssim for level 4; weights 0.0448 0.2856 0.3001 0.2363 0.1333). */
size_t gi2d_ms_ssim_workspace_size(int img_height, int img_width);
int gi2d_ms_ssim(int img_height, int img_width, const float *render_hwc, const float *gt_hwc,
                 const uint8_t *gt_u8_hwc, double *level_sums, void *workspace, size_t workspace_bytes,
                 gi2d_stream_t stream);

/* MS-SSIM as a TRAINING loss -- `Fusion4` and `Fusion_hinerv` of models/utils.py:76-79:
 *   loss = l1_weight * l1 + msssim_weight * (1 - ms_ssim(clamp(render), gt, data_range = 1, win_size = win)),
 * win = 11 (pytorch_msssim's default) or 5 (Fusion_hinerv).  v_out_hwc f32[H,W,3] = d loss / d render (through
 * torch.clamp's mask, the five levels, their avg_pool2d chain and the relu / weighted product of ms_ssim);
 * *ms_value (DEVICE pointer, optional) = ms_ssim.  l1_scale = l1_weight / (3 H W).  The smaller image side must
 * exceed (win - 1) * 16 pixels (pytorch_msssim's own assertion). */
size_t gi2d_msssim_grad_workspace_size(int img_height, int img_width);
int gi2d_image_msssim_loss_grad(int img_height, int img_width, int win, const float *render_hwc, const float *gt_hwc,
                                const uint8_t *gt_u8_hwc, float msssim_weight, float l1_scale, float *v_out_hwc,
                                double *ms_value, void *workspace, size_t workspace_bytes, gi2d_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Measurement utilities for bench.py (these two SYNCHRONISE; never call them while capturing).
 * ------------------------------------------------------------------------------------------ */

/* One full fit step with a CUDA event between kernels.  ms_host[0..4] (HOST pointer) =
 * [adam +] project + count, device-wide scan of the tile counts (0 up to 2048 tiles), place,
 * sort + raster fwd+bwd, 0.  (Event brackets add a few us per kernel: see gi2d_fit_profile_raster.) */
int gi2d_fit_profile(const gi2d_fit_params *p, const gi2d_fit_buffers *b, float *ms_host,
                     gi2d_stream_t stream);

/* Duration of the rasterize kernel alone: one full step, then `reps` back-to-back launches of
 * fit_raster_kernel<Fit> on the state that step left, between two events; *ms_host = average per launch.
 * (CUDA events around ONE ~20 us kernel add several us of launch / drain latency; back-to-back replays do
 * not.)  Leaves no gradient pending and the loss accumulators zeroed; b->grads holds garbage afterwards. */
int gi2d_fit_profile_raster(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int reps, float *ms_host,
                            gi2d_stream_t stream);

/* FP32 FMA throughput of the current device in TFLOP/s (HOST pointer): the roofline denominator
 * of the rasterize kernels. */
int gi2d_measure_fp32_peak(float *tflops_host, gi2d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GI2D_H_ */
