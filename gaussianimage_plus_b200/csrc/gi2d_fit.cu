// gi2d_fit.cu -- the fused, synchronisation-free fit step: everything one
// `train_iter` of models/gaussianimage_covariance.py:249-259 does on the device, in 3 launches
// for images of up to 2048 tiles (768x512) and 6 beyond, with no host round trip, no allocation
// and a device-side step counter, so the whole iteration replays from one CUDA graph.
//
//   K1 fit_project_kernel   [Adam of the previous step] + projection (R1) + colour activation +
//                           packed 32-B record per Gaussian + zeroing of the gradient row + the
//                           per-tile overlap COUNT (one red.global per touched tile)   [HBM/latency]
//   K2 fit_place_kernel     prefix sum over the per-tile overlap counts -> tile ranges and
//                           num_intersects (redone per CTA in shared memory for <= 2048 tiles; a
//                           3-launch device-wide scan beyond), then the counting-sort placement of
//                           every (tile, gaussian) pair: slot = start[tile] + atomic cursor; writes
//                           the 64-bit key (tile<<32|gaussian) and the 32-B record, so a tile's
//                           Gaussians are ONE contiguous block; step/lr bookkeeping     [HBM/latency]
//   K3 fit_raster_kernel    per 16x16 tile: finish the key sort (the placement is ordered by tile
//                           only: the <= 256 entries of the tile are rank-sorted by gaussian id in
//                           shared memory while they are staged, and written back sorted);
//                           rasterize-sum forward (R5), loss gradient and squared error,
//                           rasterize-sum backward (R6) with register accumulation + transposed
//                           warp reduction + one red.global per (tile, Gaussian, component) [FP32 issue]
//   (K5 fit_adam_kernel     stand-alone flush of a pending optimiser step; in steady state K1 does it)
//
// The 64-bit key sort of the reference (torch.sort of tile<<32|depth-bits, stable, Gaussian-major
// emission: utils.py:301, forward.cu:187-196) orders by tile, then by ascending Gaussian id.  Here
// the tile word is sorted by ONE counting pass whose digit is the whole tile id (no limit on the
// number of tiles, no multi-pass radix), the gaussian word by a comparison-rank sort inside the
// tile: same keys, same order, bit for bit (tests/test_gpu_fit.py::test_fit_binning_bit_exact).
// The LSD radix sort of arbitrary 64-bit keys lives in gi2d_binning.cu (gi2d_sort_pairs_i64).
#include "gi2d_project_core.cuh"
#include "gi2d_raster_core.cuh"
#include "gi2d_raster_quad.cuh"
#include "gi2d_scan.cuh"

namespace gi2d {

// implemented in gi2d_binning.cu: inclusive prefix sum of n ints (1 launch up to 2048 items, else 3)
int cumsum_i32_launch(int n, const int32_t *in, int32_t *out, int32_t *total, int32_t *block_sums,
                      cudaStream_t st);
size_t cumsum_i32_workspace(int n);

size_t msssim_grad_workspace_floats(int H, int W);
int msssim_grad_launch(int H, int W, int win, const float *render, const float *gt, const uint8_t *gt_u8, float *ws,
                       float weight, float l1_scale, float *v_out, double *value, cudaStream_t st);
int ssim_grad_launch(int H, int W, const float *render, const float *gt, const uint8_t *gt_u8, float *dm_ws,
                     float ssim_weight, float l2_scale, float l1_scale, float *v_out, double *ssim_sum,
                     cudaStream_t st);  // gi2d_loss.cu

namespace {

constexpr int kMaxSmemTiles = 2048;               // tile starts are scanned in shared memory up to here
constexpr int kMaxOrderTiles = 2048;                // the work list is built up to here (one batch; at 10880 tiles the one CTA took 57 us for a 3 % faster rasterizer)
constexpr int kOrderPerThread = 32;                 // tiles per thread of the CTA that orders the rasterizer's work list
constexpr int kProjThreads = 256;                   // launch bound; small scenes launch 64-thread CTAs (latency bound:
                                                    // spread over the SMs), large ones 256
constexpr int kPlaceWarps = 8;
constexpr int kPlaceThreads = kPlaceWarps * 32;
constexpr int kPlaceGpw = 8;                      // Gaussians per warp of K2
constexpr int kRasterThreads = 256;
constexpr int kRasterWarps = kRasterThreads / 32;

// private slots of the stats block (beyond the public GI2D_STAT_* ones)
constexpr int kStatB1Pow = 4;  // beta1^step
constexpr int kStatB2Pow = 5;  // beta2^step
constexpr int kStatStepSize = 6;  // lr / (1 - beta1^step)   of the step in flight
constexpr int kStatBc2Sqrt = 7;   // sqrt(1 - beta2^step)
constexpr int kStatPending = 8;   // != 0: grads of the last step have not been applied yet (Adam is folded
                                  // into the NEXT step's projection kernel, or flushed by gi2d_fit_adam)

[[maybe_unused]] constexpr int kStatDebug = 15;      // GI2D_DEBUG_CHECKS builds: number of violated device-side invariants (stays 0)
constexpr int kStatNonPsdAcc = 12;
// bucketed binning (see fit_project_kernel<true>): which of the two per-tile counter arrays holds the counts of
// the most recent forward, the intersection / largest-tile accumulators of the forward in flight (integers in
// the slots' 8 bytes) and the ticket that elects the last CTA of K1
constexpr int kStatBank = 85, kStatMaxFillAcc = 87, kStatTicket = 88;  // accumulator behind GI2D_STAT_NON_PSD (moved + zeroed by the clear kernel)

// compute-sanitizer is closed on the development pool: a -DGI2D_DEBUG_CHECKS build counts violated index /
// range invariants of the binning and staging code into stats[kStatDebug] instead (tools/debug_checks.py).
#ifdef GI2D_DEBUG_CHECKS
#define GI2D_CHECK(stats, cond) do { if (!(cond)) atomicAdd((stats) + kStatDebug, 1.0); } while (0)
#else
#define GI2D_CHECK(stats, cond) do { } while (0)
#endif

// Live Gaussian count: the launch-time num_points, or (dynamic_points) the device-side count that
// gi2d_fit_prune / gi2d_fit_densify maintain -- num_points is then the CAPACITY of the per-Gaussian arrays.
__device__ __forceinline__ int live_points(const gi2d_fit_params &p, const double *__restrict__ stats) {
    return p.dynamic_points ? (int)__ldcg(stats + GI2D_STAT_NUM_POINTS) : p.num_points;
}

// Squared error of the last training step: the 64 partials summed by ONE warp in a fixed order, so that
// every CTA of the optimiser kernel and the bookkeeping thread take the same best-so-far decision.
__device__ __forceinline__ double sse_total_warp(const double *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    double v = __ldcg(stats + GI2D_STAT_SSE + lane) + __ldcg(stats + GI2D_STAT_SSE + 32 + lane);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Was the step whose gradient is pending a new best (train.py:132: `if best_psnr < psnr`)?  Warp 0 of the
// CTA evaluates, everybody reads s_flag after the caller's __syncthreads().  Returns the squared error of
// that step (in warp 0; 0 elsewhere).
__device__ __forceinline__ double best_flag_warp0(const double *__restrict__ stats, bool candidate, int *s_flag) {
    double tot = 0.0;
    if (threadIdx.x < 32) {
        tot = sse_total_warp(stats);
        if (threadIdx.x == 0) *s_flag = (candidate && tot < __ldcg(stats + GI2D_STAT_BEST_SSE)) ? 1 : 0;
    }
    return tot;
}

// Bookkeeping half of the same decision, by warp 0 of ONE CTA after every optimiser thread has read it.
__device__ __forceinline__ void best_commit_warp0(double *__restrict__ stats) {
    const double tot = sse_total_warp(stats);
    if ((threadIdx.x & 31) == 0 && stats[kStatPending] != 0.0 && stats[GI2D_STAT_OVERFLOW] == 0.0 &&
        tot < stats[GI2D_STAT_BEST_SSE]) {
        stats[GI2D_STAT_BEST_SSE] = tot;
        stats[GI2D_STAT_BEST_STEP] = stats[GI2D_STAT_STEP];
        stats[GI2D_STAT_BEST_N] = stats[GI2D_STAT_NUM_POINTS];   // (rows of the snapshot the optimiser threads wrote)
    }
}

// GI2D_TILE_ORDER=0 switches the heaviest-first work list off (A/B timing)
bool tile_order_enabled() {
    static const bool on = [] {
        const char *e = getenv("GI2D_TILE_ORDER");
        return !e || atoi(e) != 0;
    }();
    return on;
}

int raster_variant(int num_tiles);

// GI2D_BUCKET=0 switches the bucketed binning off (A/B timing; the scan + placement path is also what the
// tile-row split and gi2d_bin_sort use)
bool bucket_enabled() {
    static const bool on = [] {
        const char *e = getenv("GI2D_BUCKET");
        const char *r = getenv("GI2D_RASTER");
        return (!e || atoi(e) != 0) && !(r && atoi(r) == 0);
    }();
    return on;
}

struct Plan {
    int num_tiles;
    bool smem_scan;    // tile starts fit the in-kernel scan
    bool ordered;      // K2 also emits the heaviest-first work list the rasterizer's CTAs follow
    int bucket_cap;    // > 0: bucketed binning -- tile t owns rows [t * bucket_cap, (t+1) * bucket_cap) of the key /
                       // record arrays, K1 places straight into them and there is no scan and no K2
    int gpb;           // Gaussians per CTA of K2
    int nblocks;       // CTAs of K2
};

Plan make_plan(const gi2d_fit_params &p) {
    Plan pl;
    pl.num_tiles = p.tiles_x * p.tiles_y;
    pl.smem_scan = pl.num_tiles <= kMaxSmemTiles;
    // heaviest-first tile order: only where the grid is a wave or two (beyond, the tail is a few per cent and a
    // band of a tile-row split keeps its launch order)
    // bucketed binning: everywhere except the round-1 rasterizer and the tile-row split (measured at 8192^2 / 1M
    // over 8 GPUs: a bucketed band step that walks all 1M boxes through the warp-cooperative placement takes
    // 274-293 us, count + scan + placement 249-261 us: seven of eight warps find nothing to place); needs at
    // least 8 rows per tile (tiny capacities -- overflow tests -- keep the scan)
    pl.bucket_cap = 0;
    if (p.external_optimizer != 2 && bucket_enabled() && pl.num_tiles > 0 && p.isect_capacity / pl.num_tiles >= 8)
        pl.bucket_cap = p.isect_capacity / pl.num_tiles;
    // (bucketed: an extra CTA of K1 builds the list from the PREVIOUS forward's counts; up to 16384 tiles)
    pl.ordered = (pl.bucket_cap ? pl.num_tiles <= kMaxOrderTiles : pl.smem_scan) && p.tile_row_begin == 0 &&
                 p.tile_row_end == p.tiles_y && tile_order_enabled();
    // at most ~16 CTAs per SM of K2: beyond that, more Gaussians per CTA
    int gpb = kPlaceWarps * kPlaceGpw;
    while ((long long)gpb * 2368 < p.num_points) gpb *= 2;
    pl.gpb = gpb;
    pl.nblocks = p.num_points > 0 ? cdiv(p.num_points, gpb) : 1;
    return pl;
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct Workspace {
    int32_t *tile_count;  // [T] overlap count per tile (K1 adds, K3 hands it back zeroed)
    int32_t *tile_fill;   // [T] placement cursor per tile (K2 adds, K3 hands it back zeroed)
    int32_t *tile_incl;   // [T] inclusive prefix of tile_count (only beyond kMaxSmemTiles tiles)
    int32_t *scan_ws;     //     block sums of that scan
    ushort4 *boxes;       // [N] clipped tile box per Gaussian
    int32_t *n_isect;     // [1] device copy of num_intersects (clamped to capacity)
    int4 *tile_work;      // [T] heaviest-first work list of the rasterizer (only up to kMaxSmemTiles tiles)
    float4 *records;      // [capacity][2] projected record of every intersection, tile order
    uint64_t *keys_tmp;   // [capacity] scratch for tiles with more than 256 entries (full in-tile sort)
    float *loss_render;   // [H,W,3] unclamped render   } only with loss_ssim_weight != 0:
    float *loss_dm;       // [9][H,W] SSIM partials      } the rasterize launch is split around
    float *loss_vout;     // [H,W,3] dL/d(out)           } the SSIM gradient kernels
    size_t total;
};

Workspace carve(const gi2d_fit_params &p, const Plan &pl, void *base) {
    Workspace w;
    char *c = (char *)base;
    size_t off = 0;
    const size_t T = (size_t)(pl.num_tiles > 0 ? pl.num_tiles : 1);
    // (tile_count and tile_fill are adjacent: one memset zeroes both when the workspace is created)
    w.tile_count = (int32_t *)(c + off);  off += align_up(T * 4);
    w.tile_fill = (int32_t *)(c + off);   off += align_up(T * 4);
    w.tile_incl = nullptr;
    w.scan_ws = nullptr;
    if (!pl.smem_scan) {
        w.tile_incl = (int32_t *)(c + off);  off += align_up(T * 4);
        w.scan_ws = (int32_t *)(c + off);    off += align_up(cumsum_i32_workspace((int)T));
    }
    w.boxes = (ushort4 *)(c + off);       off += align_up((size_t)(p.num_points > 0 ? p.num_points : 1) * 8);
    w.n_isect = (int32_t *)(c + off);     off += 256;
    w.tile_work = nullptr;
    if (pl.smem_scan || pl.bucket_cap) { w.tile_work = (int4 *)(c + off); off += align_up(T * 16); }
    w.records = (float4 *)(c + off);      off += align_up((size_t)p.isect_capacity * 32);
    w.keys_tmp = (uint64_t *)(c + off);   off += align_up((size_t)p.isect_capacity * 8);
    w.loss_render = w.loss_dm = w.loss_vout = nullptr;
    if (p.loss_ssim_weight != 0.f || p.loss_msssim_weight != 0.f) {
        const size_t px = (size_t)p.img_width * p.img_height;
        w.loss_render = (float *)(c + off);  off += align_up(px * 3 * 4);
        // (loss_dm: the 9 derivative planes of SSIM, or the whole workspace of the MS-SSIM gradient)
        const size_t dm_floats = p.loss_msssim_weight != 0.f ? msssim_grad_workspace_floats(p.img_height, p.img_width) : px * 9;
        w.loss_dm = (float *)(c + off);      off += align_up(dm_floats * 4);
        w.loss_vout = (float *)(c + off);    off += align_up(px * 3 * 4);
    }
    w.total = off;
    return w;
}


__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }

// ------------------------------------------------------------------------- Adam (shared)
struct AdamPtrs {
    float *xyz, *cov, *rgb, *m_xyz, *v_xyz, *m_cov, *v_cov, *m_rgb, *v_rgb;
};

// Projection backward (R7, backward2d.cu:157-214: v_cov = -X G X with the off-diagonal summed,
// v_mean = v_xy; for culled Gaussians conic == 0 and the incoming gradients are 0, the same zeros
// the reference's early return leaves) followed by torch.optim.Adam's update
// (torch/optim/adam.py _single_tensor_adam: step_size = lr/(1-beta1^t), denom = sqrt(v)/sqrt(1-beta2^t)+eps;
// the scalars were evaluated in double by the bookkeeping thread, tensors are float) for ONE Gaussian.
// Every load is issued before the first dependent store.  sqrt / divide go through the SFU
// (sqrt.approx, div.approx: 1-2 ulp): the IEEE versions branch into slow paths on the denormal second
// moments Adam produces with eps = 1e-15; 2 ulp is far inside the 1e-4 budget of the parameters.
// Returns the updated parameters in (x, c, q) so the caller can go on projecting them.
struct AdamRegs {     // everything the optimiser thread of one Gaussian reads
    float2 x, mx, vx;
    float c[3], mc[3], vc[3], q[3], mq[3], vq[3];
    float4 p0, p1, g0, g1;   // projected record and gradient row of the step being applied
};

__device__ __forceinline__ void adam_load(const AdamPtrs &a, int g, const float4 *proj, const float4 *grads,
                                          AdamRegs &r) {
    r.p0 = __ldcg(proj + 2 * g);   r.p1 = __ldcg(proj + 2 * g + 1);
    r.g0 = __ldcg(grads + 2 * g);  r.g1 = __ldcg(grads + 2 * g + 1);
    r.x = reinterpret_cast<float2 *>(a.xyz)[g];
    r.mx = reinterpret_cast<float2 *>(a.m_xyz)[g];
    r.vx = reinterpret_cast<float2 *>(a.v_xyz)[g];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.c[k] = a.cov[3 * g + k];  r.mc[k] = a.m_cov[3 * g + k];  r.vc[k] = a.v_cov[3 * g + k];
        r.q[k] = a.rgb[3 * g + k];  r.mq[k] = a.m_rgb[3 * g + k];  r.vq[k] = a.v_rgb[3 * g + k];
    }
}

__device__ __forceinline__ void adam_apply_scalars(const gi2d_fit_params &p, const AdamPtrs &a, int g, AdamRegs &r,
                                                   float step_size, float bc2_sqrt);

__device__ __forceinline__ void adam_apply(const gi2d_fit_params &p, const AdamPtrs &a, int g, AdamRegs &r,
                                           const double *__restrict__ stats) {
    adam_apply_scalars(p, a, g, r, (float)__ldcg(stats + kStatStepSize), (float)__ldcg(stats + kStatBc2Sqrt));
}

__device__ __forceinline__ void adam_apply_scalars(const gi2d_fit_params &p, const AdamPtrs &a, int g, AdamRegs &r,
                                                   float step_size, float bc2_sqrt) {
    const float w1 = (float)(1.0 - (double)p.beta1), w2 = (float)(1.0 - (double)p.beta2);
    float gc[3];
    conic_vjp(r.p0.z, r.p0.w, r.p1.x, r.g0.z, r.g0.w, r.g1.x, gc[0], gc[1], gc[2]);
    float gq[3] = {r.g1.y, r.g1.z, r.g1.w};
    if (p.color_sigmoid) {
        gq[0] *= r.p1.y * (1.f - r.p1.y);
        gq[1] *= r.p1.z * (1.f - r.p1.z);
        gq[2] *= r.p1.w * (1.f - r.p1.w);
    }
    const float inv_bc2 = 1.f / bc2_sqrt;
    auto adam = [&](float &param, float &m, float &v, float grad) {
        m = m + (grad - m) * w1;                                // exp_avg.lerp_(grad, 1-beta1)
        v = v * p.beta2 + w2 * grad * grad;                     // mul_(beta2).addcmul_(grad, grad, 1-beta2)
        float sq;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
        const float denom = fmaf(sq, inv_bc2, p.eps);
        param = param - step_size * __fdividef(m, denom);       // addcdiv_(exp_avg, denom, -step_size)
    };
    adam(r.x.x, r.mx.x, r.vx.x, r.g0.x);
    adam(r.x.y, r.mx.y, r.vx.y, r.g0.y);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        adam(r.c[k], r.mc[k], r.vc[k], gc[k]);
        adam(r.q[k], r.mq[k], r.vq[k], gq[k]);
    }
    reinterpret_cast<float2 *>(a.xyz)[g] = r.x;
    reinterpret_cast<float2 *>(a.m_xyz)[g] = r.mx;
    reinterpret_cast<float2 *>(a.v_xyz)[g] = r.vx;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a.cov[3 * g + k] = r.c[k];  a.m_cov[3 * g + k] = r.mc[k];  a.v_cov[3 * g + k] = r.vc[k];
        a.rgb[3 * g + k] = r.q[k];  a.m_rgb[3 * g + k] = r.mq[k];  a.v_rgb[3 * g + k] = r.vq[k];
    }
}

// load + (unless vetoed) apply; returns the resulting parameters in (x, c, q)
__device__ __forceinline__ void adam_update_gaussian(const gi2d_fit_params &p, const AdamPtrs &a, int g,
                                                     const float4 *proj, const float4 *grads,
                                                     const double *__restrict__ stats, bool skip,
                                                     float2 &x, float (&c)[3], float (&q)[3]) {
    AdamRegs r;
    adam_load(a, g, proj, grads, r);
    if (!skip) adam_apply(p, a, g, r, stats);
    x = r.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) { c[k] = r.c[k]; q[k] = r.q[k]; }
}

// Bookkeeping of the step in flight, by warp 0 of ONE CTA of K2 (K1 has already consumed the previous
// step's values; K3 has not started): commit the best-so-far decision, zero the loss accumulators and
// the overflow flag, and -- for a training step -- advance the step counter, the bias-correction powers
// and the StepLR schedule.  torch evaluates beta^t and gamma^floor((t-1)/size) in double precision as
// well.  The Adam that uses them runs inside the NEXT K1 (or gi2d_fit_adam).
__device__ __forceinline__ void step_bookkeeping_warp0(const gi2d_fit_params &p, int with_backward,
                                                       double *__restrict__ stats, bool overflow) {
    best_commit_warp0(stats);  // (the optimiser threads of K1 took the same decision for their snapshot)
    __syncwarp();
    if (with_backward) {       // (a render-only call accumulates no loss: the last step's numbers stay readable)
        stats[GI2D_STAT_SSE + threadIdx.x] = 0.0;
        stats[GI2D_STAT_SSE + 32 + threadIdx.x] = 0.0;
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        // A step whose intersection list does not fit the buffers is a NO-OP for the optimiser: its (truncated)
        // gradient is never applied and neither the step counter, the bias-correction powers nor the StepLR
        // schedule move, so the host can re-run exactly the iterations that did not happen (device step vs the
        // number it asked for) after it has grown the buffers.  The flag is recomputed by every step.
        stats[GI2D_STAT_OVERFLOW] = overflow ? 1.0 : 0.0;
        if (with_backward) {
            stats[GI2D_STAT_SSIM_SUM] = 0.0;
            stats[GI2D_STAT_ABS_SUM] = 0.0;
        }
        stats[kStatPending] = (with_backward && !p.external_optimizer && !overflow) ? 1.0 : 0.0;
        if (with_backward && !overflow && p.external_optimizer != 2) {
            const double step = stats[GI2D_STAT_STEP] + 1.0;
            stats[GI2D_STAT_STEP] = step;
            stats[kStatB1Pow] *= (double)p.beta1;
            stats[kStatB2Pow] *= (double)p.beta2;
            const long long k = (long long)step - 1;
            if (k > 0 && p.lr_step_size > 0 && k % p.lr_step_size == 0) stats[GI2D_STAT_LR] *= (double)p.lr_gamma;
            stats[kStatStepSize] = stats[GI2D_STAT_LR] / (1.0 - stats[kStatB1Pow]);
            stats[kStatBc2Sqrt] = sqrt(1.0 - stats[kStatB2Pow]);
        }
    }
}

// The same bookkeeping by ONE thread from values it loaded before anybody could have changed them (the
// bucketed K1: its last CTA is known only after a ticket atomic, and every load after that atomic would be one
// more L2 round trip on the step's critical path).
struct BookInputs {
    double pending, overflow_prev, best_sse, step, num_points, b1pow, b2pow, lr, sse_total;
};

__device__ __forceinline__ BookInputs book_inputs_load(const double *__restrict__ stats) {
    BookInputs k;
    k.pending = __ldcg(stats + kStatPending);
    k.overflow_prev = __ldcg(stats + GI2D_STAT_OVERFLOW);
    k.best_sse = __ldcg(stats + GI2D_STAT_BEST_SSE);
    k.step = __ldcg(stats + GI2D_STAT_STEP);
    k.num_points = __ldcg(stats + GI2D_STAT_NUM_POINTS);
    k.b1pow = __ldcg(stats + kStatB1Pow);
    k.b2pow = __ldcg(stats + kStatB2Pow);
    k.lr = __ldcg(stats + GI2D_STAT_LR);
    k.sse_total = 0.0;
    return k;
}

__device__ __forceinline__ void step_bookkeeping_values(const gi2d_fit_params &p, int with_backward,
                                                        double *__restrict__ stats, bool overflow,
                                                        const BookInputs &k) {
    if (k.pending != 0.0 && k.overflow_prev == 0.0 && k.sse_total < k.best_sse) {   // best_commit_warp0
        stats[GI2D_STAT_BEST_SSE] = k.sse_total;
        stats[GI2D_STAT_BEST_STEP] = k.step;
        stats[GI2D_STAT_BEST_N] = k.num_points;
    }
    stats[GI2D_STAT_OVERFLOW] = overflow ? 1.0 : 0.0;
    if (with_backward) {
#pragma unroll 8
        for (int i = 0; i < GI2D_STAT_SSE_SLOTS; ++i) stats[GI2D_STAT_SSE + i] = 0.0;
        stats[GI2D_STAT_SSIM_SUM] = 0.0;
        stats[GI2D_STAT_ABS_SUM] = 0.0;
    }
    stats[kStatPending] = (with_backward && !p.external_optimizer && !overflow) ? 1.0 : 0.0;
    if (with_backward && !overflow && p.external_optimizer != 2) {
        const double step = k.step + 1.0;
        stats[GI2D_STAT_STEP] = step;
        const double b1 = k.b1pow * (double)p.beta1, b2 = k.b2pow * (double)p.beta2;
        stats[kStatB1Pow] = b1;
        stats[kStatB2Pow] = b2;
        double lr = k.lr;
        const long long kk = (long long)step - 1;
        if (kk > 0 && p.lr_step_size > 0 && kk % p.lr_step_size == 0) lr *= (double)p.lr_gamma;
        stats[GI2D_STAT_LR] = lr;
        stats[kStatStepSize] = lr / (1.0 - b1);
        stats[kStatBc2Sqrt] = sqrt(1.0 - b2);
    }
}

// Ticket of the bucketed K1 (thread 0 of every CTA): ONE returning atomic carries the ticket (bits 0-19), the
// CTA's intersection count (bits 20-62) and "a tile overflowed" (bit 63, added by the first CTA to raise
// kStatMaxFillAcc from zero); the old value it returns holds the totals of all the others, and whoever sees the
// count complete does the step's bookkeeping from values it loaded before anybody could change them: the tail
// of the kernel is one L2 round trip.
__device__ __forceinline__ void k1_ticket(const gi2d_fit_params &p, int with_backward, double *__restrict__ stats,
                                          int bank, unsigned long long isects, bool first_overflow,
                                          const BookInputs &bk) {
    const unsigned long long word = 1ull | (isects << 20) | (first_overflow ? (1ull << 63) : 0ull);
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(stats + kStatTicket), word);
    const unsigned long long sum = old + word;
    if ((sum & 0xFFFFFull) != (unsigned long long)gridDim.x) return;
    const bool overflow = (sum >> 63) != 0ull;
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(stats);
    step_bookkeeping_values(p, with_backward, stats, overflow, bk);
    stats[GI2D_STAT_ISECTS] = (double)((sum >> 20) & ((1ull << 43) - 1ull));
    stats[GI2D_STAT_MAX_TILE] = overflow ? (double)__ldcg(acc + kStatMaxFillAcc) : 0.0;
    stats[kStatBank] = (double)bank;
    if (overflow) acc[kStatMaxFillAcc] = 0ull;
    acc[kStatTicket] = 0ull;
}

// The extra CTA of the bucketed K1 that orders the rasterizer's work list: tiles by descending count of the
// PREVIOUS forward (that counter array is read-only while K1 runs; K3 zeroes it afterwards) -- a counting sort
// over min(count, 255), ties in any order.  Counts move slowly from step to step: the long tiles start first,
// the short ones fill the tail of the rasterizer's last wave.  Every global load is issued up front (both
// counter arrays: which one is the previous forward's comes back with the same round trip).
__device__ __forceinline__ void k1_order_cta(const gi2d_fit_params &p, int with_backward, double *__restrict__ stats,
                                             const int32_t *__restrict__ tile_count,
                                             const int32_t *__restrict__ tile_fill, int32_t *__restrict__ tile_order,
                                             int *s_bin) {
    const int lane = threadIdx.x & 31;
    const int T = p.tiles_x * p.tiles_y;
    const int per_batch = kOrderPerThread * (int)blockDim.x;   // one batch up to 2048 tiles (64 threads); more loop
    int ca[kOrderPerThread], cb[kOrderPerThread];
#pragma unroll
    for (int k = 0; k < kOrderPerThread; ++k) {
        const int t = k * blockDim.x + threadIdx.x;
        ca[k] = t < T ? __ldcg(tile_count + t) : -1;
        cb[k] = t < T ? __ldcg(tile_fill + t) : -1;
    }
    const int bank = 1 - (int)__ldcg(stats + kStatBank);   // the array this forward fills; the other is `prev`
    const int32_t *prev = bank ? tile_count : tile_fill;
    BookInputs bk;
    if (threadIdx.x == 0) bk = book_inputs_load(stats);
    double sse_tot = 0.0;
    if (threadIdx.x < 32) sse_tot = sse_total_warp(stats);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bin[i] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kOrderPerThread; ++k) {
        ca[k] = min(bank ? ca[k] : cb[k], 255);
        if (ca[k] >= 0) atomicAdd(&s_bin[255 - ca[k]], 1);
    }
    for (int base = per_batch; base < T; base += per_batch) {   // (more than one batch: histogram of the rest)
#pragma unroll
        for (int k = 0; k < kOrderPerThread; ++k) {
            const int t = base + k * blockDim.x + threadIdx.x;
            cb[k] = t < T ? min(__ldcg(prev + t), 255) : -1;
        }
#pragma unroll
        for (int k = 0; k < kOrderPerThread; ++k)
            if (cb[k] >= 0) atomicAdd(&s_bin[255 - cb[k]], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = s_bin[8 * lane + k]; sum += v[k]; }
        int run = warp_scan_inclusive(sum) - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_bin[8 * lane + k] = run; run += v[k]; }
    }
    __syncthreads();
    for (int base = 0; base < T; base += per_batch) {
        if (base > 0) {
#pragma unroll
            for (int k = 0; k < kOrderPerThread; ++k) {
                const int t = base + k * blockDim.x + threadIdx.x;
                ca[k] = t < T ? min(__ldcg(prev + t), 255) : -1;
            }
        }
#pragma unroll
        for (int k = 0; k < kOrderPerThread; ++k)
            if (ca[k] >= 0) {
                const int t = base + k * blockDim.x + threadIdx.x;
                const int pos = atomicAdd(&s_bin[255 - ca[k]], 1);
                const int ty = t / p.tiles_x;
                tile_order[pos] = (t - ty * p.tiles_x) | (ty << 16);
            }
    }
    if (threadIdx.x == 0) {
        bk.sse_total = sse_tot;
        k1_ticket(p, with_backward, stats, bank, 0ull, false, bk);
    }
}

// Placement of a warp's intersections into the buckets + the ticket (shared by the bucketed K1 and the tile-row
// placement kernel).  Every thread of the CTA calls it, warp-converged; (x0,y0,x1,y1) is the thread's tile box
// (empty: x1 == x0), (rec0, rec1) its Gaussian's record, g its Gaussian id (lane-consecutive within the warp).
// s_isect: kProjThreads / 32 ints of shared memory; *s_ovf must have been zeroed before the CTA's last barrier.
__device__ __forceinline__ void bucket_place_and_ticket(const gi2d_fit_params &p, int with_backward,
                                                        double *__restrict__ stats, int bank, const BookInputs &bk,
                                                        int g, int x0, int y0, int x1, int y1, float4 rec0,
                                                        float4 rec1, int32_t *__restrict__ tile_count,
                                                        int32_t *__restrict__ tile_fill,
                                                        uint64_t *__restrict__ keys_out,
                                                        float4 *__restrict__ records, int bucket_cap, int *s_isect,
                                                        int *s_ovf) {
    const int lane = threadIdx.x & 31;
    int32_t *cursor = bank ? tile_fill : tile_count;
    const int bw = x1 - x0;
    const int n = bw * (y1 - y0);
    const int incl = warp_scan_inclusive(n);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int max_fill = 0;
    for (int batch = 0; batch < total; batch += 256) {
        int slot[8], tile[8], owner[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            slot[k] = 0; tile[k] = 0; owner[k] = 0;
            if (batch + 32 * k < total) {   // warp-uniform
                const int it = batch + 32 * k + lane;
                int lo = 0;   // owner = smallest j with incl_j > it
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                    if (probe <= it) lo += step;
                }
                const int ow = min(lo, 31);
                const int o_incl = __shfl_sync(0xffffffffu, incl, ow);
                const int o_n = __shfl_sync(0xffffffffu, n, ow);
                const int o_w = __shfl_sync(0xffffffffu, bw, ow);
                const int o_x0 = __shfl_sync(0xffffffffu, x0, ow);
                const int o_y0 = __shfl_sync(0xffffffffu, y0, ow);
                owner[k] = ow;
                if (it < total) {
                    const int kk = it - (o_incl - o_n);
                    const int ry = kk / o_w;
                    tile[k] = (o_y0 + ry) * p.tiles_x + o_x0 + (kk - ry * o_w);
                    GI2D_CHECK(stats, tile[k] >= 0 && tile[k] < p.tiles_x * p.tiles_y);
                    slot[k] = atomicAdd(cursor + tile[k], 1);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (batch + 32 * k < total) {   // warp-uniform
                const int ow = owner[k];
                float4 q0, q1;
                q0.x = __shfl_sync(0xffffffffu, rec0.x, ow); q0.y = __shfl_sync(0xffffffffu, rec0.y, ow);
                q0.z = __shfl_sync(0xffffffffu, rec0.z, ow); q0.w = __shfl_sync(0xffffffffu, rec0.w, ow);
                q1.x = __shfl_sync(0xffffffffu, rec1.x, ow); q1.y = __shfl_sync(0xffffffffu, rec1.y, ow);
                q1.z = __shfl_sync(0xffffffffu, rec1.z, ow); q1.w = __shfl_sync(0xffffffffu, rec1.w, ow);
                if (batch + 32 * k + lane < total) {
                    if (slot[k] < bucket_cap) {
                        const size_t pos = (size_t)tile[k] * bucket_cap + slot[k];
                        keys_out[pos] = ((uint64_t)(uint32_t)tile[k] << 32) | (uint32_t)(g - lane + ow);
                        records[2 * pos] = q0;
                        records[2 * pos + 1] = q1;
                    }
                    max_fill = max(max_fill, slot[k] + 1);
                }
            }
        }
    }
    // ---- ticket: the last CTA to get here does the bookkeeping of the step in flight (k1_ticket)
    max_fill = __reduce_max_sync(0xffffffffu, max_fill);
    if (lane == 0) s_isect[threadIdx.x >> 5] = total;
    if (lane == 0 && max_fill > bucket_cap) {   // (rare) largest count seen, for the host's regrow
        const unsigned long long old =
            atomicMax(reinterpret_cast<unsigned long long *>(stats + kStatMaxFillAcc), (unsigned long long)max_fill);
        if (old == 0ull) *s_ovf = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long mine_isects = 0ull;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mine_isects += (unsigned long long)s_isect[w];
        k1_ticket(p, with_backward, stats, bank, mine_isects, *s_ovf != 0, bk);
    }
}

// ------------------------------------------------------------------------------------ K1
// Optimiser of the PREVIOUS step + projection of THIS step, per Gaussian, in one launch: the thread
// that owns Gaussian g first applies projection-backward + Adam to it when a gradient is pending
// (stats[kStatPending]; the overflow flag of that step vetoes it), then projects the fresh parameters,
// writes the 32-B record and zeroes the gradient row for the coming backward.  Then
//   kBucket == false: adds 1 to the overlap count of every tile its box touches ("per-tile overlap counts", one
//       fire-and-forget red.global each); the scan + placement kernel (K2) follows and does the bookkeeping;
//   kBucket == true : PLACES the intersections itself.  Tile t owns the rows [t * C, (t+1) * C) of the key / record
//       arrays (C = capacity / #tiles), so a slot is just the tile's cursor atomic -- no prefix sum, no second
//       kernel: the warp walks the intersections of its 32 Gaussians cooperatively (lane = intersection, owner by
//       shuffle search, the owner's record by shuffles), all cursor atomics of up to 256 intersections in flight
//       before the first dependent store.  The order inside a tile is arbitrary and fixed by the rank sort of K3.
//       The cursors are double-buffered (stats[kStatBank]): this forward fills one array, K3 reads it and zeroes
//       the OTHER one for the next forward, so K3 can be replayed and the counts outlive the step.  A tile with
//       more than C overlaps raises the overflow flag (the step becomes an optimiser no-op, the host regrows).
//       The LAST CTA to finish (ticket) does the bookkeeping K2 does otherwise.
#ifndef GI2D_K1_MINBLOCKS
#define GI2D_K1_MINBLOCKS 2
#endif
template <bool kBucket>
__global__ void __launch_bounds__(kProjThreads, GI2D_K1_MINBLOCKS)
fit_project_kernel(gi2d_fit_params p, AdamPtrs a, const float *__restrict__ cov_bound,
                   float4 *__restrict__ proj, float4 *__restrict__ grads, ushort4 *__restrict__ boxes,
                   int32_t *__restrict__ tile_count, double *__restrict__ stats, int with_backward,
                   float4 *__restrict__ best, int expect_pending, float *__restrict__ best_bound,
                   int32_t *__restrict__ tile_fill, uint64_t *__restrict__ keys_out, float4 *__restrict__ records,
                   int bucket_cap, int32_t *__restrict__ tile_order) {
    __shared__ int s_best;
    __shared__ int s_ovf;
    __shared__ int s_bin[kBucket ? 256 : 1];
    __shared__ int s_isect[kProjThreads / 32];
    pdl_launch_dependents();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (kBucket && threadIdx.x == 0) s_ovf = 0;
    const bool in_cap = g < p.num_points;   // a row of the arrays (the live count comes back with the flags below)
    pdl_wait();  // the previous step's rasterizer wrote grads (and read proj, and zeroed the tile counters)
    if (kBucket && tile_order && blockIdx.x == gridDim.x - 1) {   // (CTA-uniform: the extra CTA has no Gaussians)
        k1_order_cta(p, with_backward, stats, tile_count, tile_fill, tile_order, s_bin);
        return;
    }
    // A training step nearly always finds a gradient pending: issue every load of the optimiser BEFORE the
    // flags that say so come back (one L2 round trip less on this latency-bound kernel); a render-only call
    // (expect_pending == 0) loads lazily.
    const bool early = expect_pending && a.m_xyz != nullptr && in_cap;
    AdamRegs r;
    if (early) adam_load(a, g, proj, grads, r);
    float bnd[3] = {0.f, 0.f, 0.f};
    if (in_cap) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bnd[k] = __ldcg(cov_bound + 3 * g + k);   // (prune / densify rewrite the bounds)
    }
    const bool mine = g < live_points(p, stats);
    const bool pending = a.m_xyz != nullptr && __ldcg(stats + kStatPending) != 0.0;
    const bool veto = __ldcg(stats + GI2D_STAT_OVERFLOW) != 0.0;  // that step overflowed: the host re-runs it
    const int bank = kBucket ? (1 - (int)__ldcg(stats + kStatBank)) : 0;   // the counter array this forward fills
    BookInputs bk;
    if (kBucket && threadIdx.x == 0) bk = book_inputs_load(stats);   // (thread 0 may turn out to be the bookkeeper)
    const double sse_tot = best_flag_warp0(stats, best != nullptr && pending && !veto, &s_best);
    if (kBucket && threadIdx.x == 0) bk.sse_total = sse_tot;
    __syncthreads();
    const bool snapshot = s_best != 0;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    float4 rec0 = make_float4(0.f, 0.f, 0.f, 0.f), rec1 = rec0;
    if (mine) {
        float2 m;
        float c[3], q[3];
        if (pending) {
            if (!early) adam_load(a, g, proj, grads, r);
            if (!veto) adam_apply(p, a, g, r, stats);
            m = r.x;
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = r.c[k]; q[k] = r.q[k]; }
            if (snapshot) {  // the state dict right after optimizer.step() of the best iteration (train.py:132-137)
                best[2 * g] = make_float4(m.x, m.y, c[0], c[1]);
                best[2 * g + 1] = make_float4(c[2], q[0], q[1], q[2]);
                if (best_bound) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) best_bound[3 * g + k] = bnd[k];
                }
            }
        } else if (early) {
            m = r.x;
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = r.c[k]; q[k] = r.q[k]; }
        } else {
            m = reinterpret_cast<const float2 *>(a.xyz)[g];
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = a.cov[3 * g + k]; q[k] = a.rgb[3 * g + k]; }
        }
        // get_cov2d_elements = _cov2d + cholesky_bound (gaussianimage_covariance.py:169)
        const float sx = __fadd_rn(c[0], bnd[0]);
        const float sxy = __fadd_rn(c[1], bnd[1]);
        const float sy = __fadd_rn(c[2], bnd[2]);
        float cr = q[0], cg = q[1], cb = q[2];
        if (p.color_sigmoid) { cr = sigmoidf(cr); cg = sigmoidf(cg); cb = sigmoidf(cb); }
        const Projected pr = project_cov(m.x, m.y, sx, sxy, sy, p.clip_coe, p.radius_clip, p.tiles_x, p.tiles_y);
        rec0 = make_float4(pr.x, pr.y, pr.a, pr.b);
        rec1 = make_float4(pr.c, cr, cg, cb);
        proj[2 * g] = rec0;
        proj[2 * g + 1] = rec1;
        if (with_backward) {
            grads[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
            grads[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // the map kernel's own cull (forward.cu:161) and the band owned by this rank
        if (pr.ntiles > 0 && !((float)pr.radius < p.radius_clip)) {
            x0 = pr.box.x0; x1 = pr.box.x1;
            y0 = max(pr.box.y0, p.tile_row_begin);
            y1 = min(pr.box.y1, p.tile_row_end);
            if (y1 <= y0) { x0 = x1 = y0 = y1 = 0; }
        }
    }
    // (a row beyond the live count keeps an empty box: it stays out of every walk)
    if (in_cap) boxes[g] = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
    if (!kBucket) {
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                GI2D_CHECK(stats, ty >= p.tile_row_begin && ty < p.tile_row_end && tx >= 0 && tx < p.tiles_x);
                atomicAdd(tile_count + ty * p.tiles_x + tx, 1);
            }
        return;
    }
    // ---- placement + ticket (warp-converged from here on)
    bucket_place_and_ticket(p, with_backward, stats, bank, bk, g, x0, y0, x1, y1, rec0, rec1, tile_count, tile_fill,
                            keys_out, records, bucket_cap, s_isect, &s_ovf);
}

// Exclusive scan of count[0..T) into shared memory, T <= kMaxSmemTiles = kThreads * kPer.
template <int kThreads>
__device__ __forceinline__ int scan_counts_to_smem(const int32_t *__restrict__ count, int T, int *s_base,
                                                   int *s_warp) {
    constexpr int kPer = kMaxSmemTiles / kThreads;
    const int i0 = threadIdx.x * kPer;
    int v[kPer];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        v[k] = (i0 + k < T) ? __ldcg(count + i0 + k) : 0;
        sum += v[k];
    }
    int total;
    const int incl = block_scan_inclusive<kThreads>(sum, s_warp, &total);
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (i0 + k < T) s_base[i0 + k] = run;
        run += v[k];
    }
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------------------------ K2
// Walk the intersections of a warp's Gaussians (Gaussian-major, tiles row-major: forward.cu:187-196),
// 32 at a time, load balanced: lane i of a chunk owns Gaussian i, an inclusive scan of the box areas
// gives every intersection its slot, a 5-step shuffle search gives every slot its owner.
// `visit(valid, tile, gaussian)` is called warp-converged.
template <class Visit>
__device__ __forceinline__ void walk_intersections(int g_begin, int g_end, int tiles_x,
                                                   const ushort4 *__restrict__ boxes, ushort4 first_box,
                                                   Visit visit) {
    const int lane = threadIdx.x & 31;
    for (int base = g_begin; base < g_end; base += 32) {
        const int g = base + lane;
        ushort4 bx = make_ushort4(0, 0, 0, 0);
        if (base == g_begin) bx = first_box;  // (prefetched by the caller: boxes[g_begin + lane] or zeros)
        else if (g < g_end) bx = __ldcg(boxes + g);
        const int w = (int)bx.z - (int)bx.x;
        const int n = w * ((int)bx.w - (int)bx.y);
        const int incl = warp_scan_inclusive(n);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int s = 0; s < total; s += 32) {
            const int it = s + lane;
            // owner = smallest j with incl_j > it
            int lo = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
                if (probe <= it) lo += step;
            }
            const int owner = min(lo, 31);
            const int o_incl = __shfl_sync(0xffffffffu, incl, owner);
            const int o_n = __shfl_sync(0xffffffffu, n, owner);
            const int o_w = __shfl_sync(0xffffffffu, w, owner);
            const int o_x0 = __shfl_sync(0xffffffffu, (int)bx.x, owner);
            const int o_y0 = __shfl_sync(0xffffffffu, (int)bx.y, owner);
            const bool valid = it < total;
            int tile = 0;
            if (valid) {
                const int k = it - (o_incl - o_n);
                const int ry = k / o_w;
                tile = (o_y0 + ry) * tiles_x + o_x0 + (k - ry * o_w);
            }
            visit(valid, tile, base + owner);
        }
    }
}

// Prefix sum over the per-tile overlap counts + placement.  The start of a tile's range comes from the
// in-CTA scan (kSmemScan: every CTA redoes the <= 8 KiB scan instead of paying a launch for it) or from
// the device-wide inclusive scan tile_incl[] that ran between K1 and this kernel.  An intersection's
// slot inside its tile's range is handed out by an atomic cursor: the order within a tile is arbitrary
// here and fixed by the rank sort of K3.  CTA 0 also publishes the tile ranges (forward.cu:211-233;
// empty tiles keep the reference's (0,0)) and does the bookkeeping.
template <bool kSmemScan, bool kFit = true>
__global__ void __launch_bounds__(kPlaceThreads)
fit_place_kernel(gi2d_fit_params p, int with_backward, int gpb, int num_tiles,
                 const ushort4 *__restrict__ boxes, const int32_t *__restrict__ tile_count,
                 const int32_t *__restrict__ tile_incl, int32_t *__restrict__ tile_fill,
                 uint64_t *__restrict__ keys_out, const float4 *__restrict__ proj, float4 *__restrict__ records,
                 int32_t *__restrict__ tile_bins, int32_t *__restrict__ n_isect, double *__restrict__ stats,
                 int4 *__restrict__ tile_work) {
    __shared__ int s_base[kSmemScan ? kMaxSmemTiles : 1];
    __shared__ int s_warp[kPlaceWarps];
    __shared__ int s_bin[kSmemScan ? kPlaceThreads : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    const int gpw = gpb / kPlaceWarps;
    const int g_begin = min(p.num_points, blockIdx.x * gpb + warp * gpw);
    const int g_end = min(p.num_points, g_begin + gpw);
    pdl_wait();
    // the first chunk of boxes is in flight while the tile starts are scanned
    ushort4 first_box = make_ushort4(0, 0, 0, 0);
    if (g_begin + lane < g_end) first_box = __ldcg(boxes + g_begin + lane);
    int total;
    if (kSmemScan) {
        total = scan_counts_to_smem<kPlaceThreads>(tile_count, num_tiles, s_base, s_warp);
    } else {
        total = num_tiles > 0 ? __ldcg(tile_incl + num_tiles - 1) : 0;
    }
    // tile ranges: the CTAs share the tiles (every CTA has all the starts)
    if (kSmemScan) {
        for (int t = blockIdx.x * kPlaceThreads + threadIdx.x; t < num_tiles; t += gridDim.x * kPlaceThreads) {
            const int c = __ldcg(tile_count + t);
            int2 rg = c ? make_int2(s_base[t], s_base[t] + c) : make_int2(0, 0);
            if (!kFit) { rg.x = min(rg.x, p.isect_capacity); rg.y = min(rg.y, p.isect_capacity); }   // rows that exist
            reinterpret_cast<int2 *>(tile_bins)[t] = rg;
        }
    } else {
        for (int t = blockIdx.x * kPlaceThreads + threadIdx.x; t < num_tiles; t += gridDim.x * kPlaceThreads) {
            const int e = __ldcg(tile_incl + t), c = __ldcg(tile_count + t);
            int2 rg = c ? make_int2(e - c, e) : make_int2(0, 0);
            if (!kFit) { rg.x = min(rg.x, p.isect_capacity); rg.y = min(rg.y, p.isect_capacity); }
            reinterpret_cast<int2 *>(tile_bins)[t] = rg;
        }
    }
    // the rasterizer's work list: tiles by descending count (counting sort over min(count, 255), ties in any
    // order), written by the LAST CTA (the one with the fewest Gaussians to place)
    if (kSmemScan && tile_work && blockIdx.x == gridDim.x - 1) {
        s_bin[threadIdx.x] = 0;
        __syncthreads();
        for (int t = threadIdx.x; t < num_tiles; t += kPlaceThreads)
            atomicAdd(&s_bin[kPlaceThreads - 1 - min(__ldcg(tile_count + t), kPlaceThreads - 1)], 1);
        __syncthreads();
        const int mine = s_bin[threadIdx.x];
        int tot;
        const int incl = block_scan_inclusive<kPlaceThreads>(mine, s_warp, &tot);
        __syncthreads();
        s_bin[threadIdx.x] = incl - mine;
        __syncthreads();
        for (int t = threadIdx.x; t < num_tiles; t += kPlaceThreads) {
            const int c = __ldcg(tile_count + t);
            const int pos = atomicAdd(&s_bin[kPlaceThreads - 1 - min(c, kPlaceThreads - 1)], 1);
            const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
            tile_work[pos] = c ? make_int4(tx | (ty << 16), s_base[t], s_base[t] + c, c) : make_int4(tx | (ty << 16), 0, 0, 0);
        }
    }
    if (blockIdx.x == 0 && warp == 0) {
        if (kFit) step_bookkeeping_warp0(p, with_backward, stats, total > p.isect_capacity);
        if (lane == 0) {
            if (kFit) stats[GI2D_STAT_ISECTS] = (double)total;
            n_isect[0] = total > p.isect_capacity ? p.isect_capacity : total;
            if (!kFit) {   // gi2d_bin_sort: {clamped count, true count, overflow}
                n_isect[1] = total;
                n_isect[2] = total > p.isect_capacity ? 1 : 0;
            }
        }
    }
    walk_intersections(g_begin, g_end, p.tiles_x, boxes, first_box, [&](bool valid, int tile, int g) {
        if (valid) {
            // (record loads and the cursor atomic are independent: all in flight together)
            float4 r0, r1;
            if (kFit) { r0 = __ldcg(proj + 2 * g); r1 = __ldcg(proj + 2 * g + 1); }
            const int start = kSmemScan ? s_base[tile] : (__ldcg(tile_incl + tile) - __ldcg(tile_count + tile));
            const int slot = atomicAdd(tile_fill + tile, 1);
            const int pos = start + slot;
            if (kFit) GI2D_CHECK(stats, tile >= 0 && tile < num_tiles && g >= 0 && g < p.num_points && start >= 0 &&
                                  slot >= 0 && slot < __ldcg(tile_count + tile));
            if (pos < p.isect_capacity) {
                keys_out[pos] = ((uint64_t)(uint32_t)tile << 32) | (uint32_t)g;
                if (kFit) {
                    records[2 * (size_t)pos] = r0;
                    records[2 * (size_t)pos + 1] = r1;
                }
            }
        }
    });
}

// ------------------------------------------------------------------------------------------- gi2d_bin_sort
// The reference's binning (utils.py:231-311: cumsum -> .item() -> map_gaussian_to_intersects -> torch.sort ->
// gather -> get_tile_bin_edges) as ONE call with num_intersects kept on the device: per-tile counts (one
// red.global per touched tile), prefix sum, counting-sort placement, in-tile rank sort -- the fit step's scheme,
// emitting the reference's arrays.  Valid when every depth has the same bit pattern (the 2-D projections emit
// 0.0): the keys then order by tile, then by Gaussian id (stable sort of Gaussian-major keys).
__global__ void __launch_bounds__(256)
bs_count_kernel(int n, const float *__restrict__ xys, const int32_t *__restrict__ radii, int tiles_x, int tiles_y,
                float radius_clip, ushort4 *__restrict__ boxes, int32_t *__restrict__ tile_count) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= n) return;
    const int r = radii[g];
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (!((float)r < radius_clip)) {   // forward.cu:161 (int radius promoted to float)
        const float2 c = reinterpret_cast<const float2 *>(xys)[g];
        const TileBox b = tile_bbox(c.x, c.y, (float)r, tiles_x, tiles_y);   // helpers.cuh:43-47 on radii[idx]
        if (b.x1 > b.x0 && b.y1 > b.y0) { x0 = b.x0; x1 = b.x1; y0 = b.y0; y1 = b.y1; }
    }
    boxes[g] = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) atomicAdd(tile_count + ty * tiles_x + tx, 1);
}

// one CTA per tile: rank the tile's entries by gaussian id, emit isect_ids_sorted (tile << 32 | depth bits) and
// gaussian_ids_sorted
__global__ void __launch_bounds__(64)
bs_tile_sort_kernel(int capacity, const int32_t *__restrict__ tile_bins, const uint64_t *__restrict__ keys,
                    int64_t depth_bits, int64_t *__restrict__ isect_ids_sorted,
                    int32_t *__restrict__ gaussian_ids_sorted) {
    __shared__ __align__(16) int s_id[kMaxPerTile + 4];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int2 range = __ldcg(reinterpret_cast<const int2 *>(tile_bins) + tile);
    const int cnt = max(0, min(range.y, capacity) - range.x);
    if (cnt == 0) return;
    const int64_t hi = ((int64_t)tile << 32) | depth_bits;
    if (cnt <= kMaxPerTile) {
        for (int e = tid; e < cnt; e += 64) s_id[e] = (int)(uint32_t)__ldcg(keys + range.x + e);
        if (tid < 4) s_id[cnt + tid] = 0x7fffffff;
        __syncthreads();
        for (int e = tid; e < cnt; e += 64) {
            const int id = s_id[e];
            int rank = 0;
            for (int j = 0; j < cnt; j += 4) {
                const int4 o = *reinterpret_cast<const int4 *>(s_id + j);
                rank += (o.x < id) + (o.y < id) + (o.z < id) + (o.w < id);
            }
            isect_ids_sorted[range.x + rank] = hi;
            gaussian_ids_sorted[range.x + rank] = id;
        }
    } else {
        for (int e = tid; e < cnt; e += 64) {
            const int id = (int)(uint32_t)__ldcg(keys + range.x + e);
            int rank = 0;
            for (int j = 0; j < cnt; ++j) rank += ((int)(uint32_t)__ldcg(keys + range.x + j) < id) ? 1 : 0;
            isect_ids_sorted[range.x + rank] = hi;
            gaussian_ids_sorted[range.x + rank] = id;
        }
    }
}

// ---- optional TMA staging of a tile's records (-DGI2D_TMA_STAGE; measured, not the default: DESIGN.md 2).
// One elected thread arms an mbarrier with the byte count and issues ONE cp.async.bulk (SASS: UBLKCP) for the
// tile's contiguous cnt x 32-B block; everybody waits on the barrier phase.
#ifdef GI2D_TMA_STAGE
__device__ __forceinline__ uint32_t smem_addr(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
#endif

// ------------------------------------------------------------------------------------ K4
// Render      : forward, clamped CHW `render` tensor out
// Fit         : forward + pointwise loss gradient (mse / l1) + backward, one launch
// FitForward  : forward + squared error + unclamped HWC image out      } the split used when the loss has an
// FitBackward : backward from a dL/d(out) image (v_out, f32[H,W,3])    } SSIM term (needs neighbouring tiles)
enum class RasterMode { Render, Fit, FitForward, FitBackward };

// 6 CTAs/SM (40 registers, no spills) measured best on all three workloads: 768x512/5k +4 %, 2040x1356/20k
// +12 %, 8192^2/1M +18 % over 4 CTAs/SM (48 registers) -- latency hiding beats the few extra registers
#ifndef GI2D_FIT_MINBLOCKS
#define GI2D_FIT_MINBLOCKS 6
#endif
template <RasterMode kMode>
__global__ void __launch_bounds__(kRasterThreads, (kMode == RasterMode::Fit || kMode == RasterMode::FitBackward) ? GI2D_FIT_MINBLOCKS : 6)
fit_raster_kernel(gi2d_fit_params p, uint64_t *__restrict__ sorted_keys, uint64_t *__restrict__ keys_tmp,
                  const int32_t *__restrict__ tile_bins, int32_t *__restrict__ tile_count,
                  int32_t *__restrict__ tile_fill, const float4 *__restrict__ records,
                  const float *__restrict__ gt, const uint8_t *__restrict__ gt_u8,
                  float *__restrict__ out_img, float *__restrict__ grads, double *__restrict__ stats,
                  float *__restrict__ err_map, const float *__restrict__ v_out) {
    constexpr bool kHasFwd = kMode != RasterMode::FitBackward;
    constexpr bool kHasLoss = kMode == RasterMode::Fit || kMode == RasterMode::FitForward;
    constexpr bool kHasBwd = kMode == RasterMode::Fit || kMode == RasterMode::FitBackward;
    __shared__ TileRecords sg;
    __shared__ TileGrad tg;
    __shared__ int s_ids[kMaxPerTile];
    __shared__ int s_sort[kMaxPerTile];
    __shared__ float s_red[2][kRasterWarps];
#ifdef GI2D_TMA_STAGE
    __shared__ __align__(128) float4 s_raw[2 * kMaxPerTile];
    __shared__ __align__(8) uint64_t s_mbar;
    if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
    __syncthreads();
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_y = p.tile_row_begin + blockIdx.y;
    const int tile_id = tile_y * p.tiles_x + blockIdx.x;
    pdl_launch_dependents();
    const int blk = warp;  // forward: warp <-> 8x4 sub-block
    const int lx = blk_x(blk, lane), ly = blk_y(blk, lane);
    const int j = blockIdx.x * kTile + lx;
    const int i = tile_y * kTile + ly;
    const bool inside = i < p.img_height && j < p.img_width;
    const size_t pix = (size_t)i * p.img_width + j;
    // issue every load that does not depend on the tile range first: target pixel, scene flag
    float tr = 0.f, tgc = 0.f, tb = 0.f;
    if (kHasLoss && inside) {
        if (gt) {
            tr = __ldg(gt + 3 * pix);
            tgc = __ldg(gt + 3 * pix + 1);
            tb = __ldg(gt + 3 * pix + 2);
        } else {
            // 8-bit target: the value torchvision's ToTensor produces, u8 / 255 (IEEE division)
            tr = u8_to_unit(__ldg(gt_u8 + 3 * pix));
            tgc = u8_to_unit(__ldg(gt_u8 + 3 * pix + 1));
            tb = u8_to_unit(__ldg(gt_u8 + 3 * pix + 2));
        }
    }
    pdl_wait();  // the target image above is written by no kernel of the step; everything below is
    const double n_isect = __ldcg(stats + GI2D_STAT_ISECTS);
    const int2 range = __ldcg(reinterpret_cast<const int2 *>(tile_bins) + tile_id);
    const int total_cnt = max(0, min(range.y, p.isect_capacity) - range.x);
    const int cnt = min(kMaxPerTile, total_cnt);
    // the counters of this tile go back to K1 / K2 of the next step zeroed
    if (kHasFwd && tid == 0) {
        GI2D_CHECK(stats, range.x >= 0 && range.y >= range.x && tile_fill[tile_id] == tile_count[tile_id] &&
                              (range.y - range.x == tile_count[tile_id] || n_isect > (double)p.isect_capacity));
        tile_count[tile_id] = 0;
        tile_fill[tile_id] = 0;
    }
    // Finish the key sort: K2 placed this tile's entries in arbitrary order.  Rank every entry by its
    // gaussian id (ids are unique within a tile, so ranks are a permutation) and stage it at its rank:
    // shared memory then holds the reference's order (ascending id == the stable sort of Gaussian-major
    // emitted keys), truncated to the first 256 (forward.cu:673).  The sorted keys are written back so that
    // sorted_keys is the fully sorted 64-bit sequence -- except by the forward half of a split step, whose
    // backward half must find keys and records still paired.
    constexpr bool kWriteBack = kMode != RasterMode::FitForward;
    const float tx0 = (float)(blockIdx.x * kTile), ty0 = (float)(tile_y * kTile);
    if (total_cnt <= kMaxPerTile) {
        uint64_t key = 0;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
#ifdef GI2D_TMA_STAGE
        if (tid == 0 && cnt > 0) {
            mbar_expect_tx(&s_mbar, (uint32_t)cnt * 32u);
            bulk_copy_g2s(s_raw, records + 2 * (size_t)range.x, (uint32_t)cnt * 32u, &s_mbar);
        }
        if (tid < cnt) {
            key = __ldcg(sorted_keys + range.x + tid);
            s_sort[tid] = (int)(uint32_t)key;
        }
        if (cnt > 0) mbar_wait(&s_mbar, 0);
        __syncthreads();
        if (tid < cnt) {
            r0 = s_raw[2 * tid];
            r1 = s_raw[2 * tid + 1];
        }
#else
        if (tid < cnt) {
            // one contiguous block of cnt x (8 + 32) B
            key = __ldcg(sorted_keys + range.x + tid);
            r0 = __ldcg(records + 2 * (size_t)(range.x + tid));
            r1 = __ldcg(records + 2 * (size_t)(range.x + tid) + 1);
            s_sort[tid] = (int)(uint32_t)key;
        }
        __syncthreads();
#endif
        if (tid < cnt) {
            const int id = (int)(uint32_t)key;
            int rank = 0;
            for (int jj = 0; jj < cnt; ++jj) rank += (s_sort[jj] < id) ? 1 : 0;
            GI2D_CHECK(stats, rank >= 0 && rank < cnt && (int)(key >> 32) == tile_id && id >= 0 && id < p.num_points);
            stage_record(sg, rank, r0, r1, tx0, ty0);
            if (kHasBwd) s_ids[rank] = id;
            if (kWriteBack && rank != tid) sorted_keys[range.x + rank] = key;
        }
    } else {
        // more than 256 entries (a degenerate scene): full rank sort straight from global memory
        for (int e = tid; e < total_cnt; e += kRasterThreads) {
            const uint64_t key = __ldcg(sorted_keys + range.x + e);
            const int id = (int)(uint32_t)key;
            int rank = 0;
            for (int jj = 0; jj < total_cnt; ++jj) rank += ((int)(uint32_t)__ldcg(sorted_keys + range.x + jj) < id) ? 1 : 0;
            GI2D_CHECK(stats, rank >= 0 && rank < total_cnt && (int)(key >> 32) == tile_id);
            keys_tmp[range.x + rank] = key;
            if (rank < kMaxPerTile) {
                stage_record(sg, rank, __ldcg(records + 2 * (size_t)(range.x + e)),
                             __ldcg(records + 2 * (size_t)(range.x + e) + 1), tx0, ty0);
                if (kHasBwd) s_ids[rank] = id;
            }
        }
        if (kWriteBack) {
            __syncthreads();
            for (int e = tid; e < total_cnt; e += kRasterThreads) sorted_keys[range.x + e] = __ldcg(keys_tmp + range.x + e);
        }
    }
    if (kMode == RasterMode::FitBackward) {
        // dL/d(out) of this tile, computed by the loss kernels between the two halves
        const int gi = grad_index(lx, ly);
#pragma unroll
        for (int c = 0; c < 3; ++c) tg.v[c][gi] = (inside && n_isect != 0.0) ? __ldcg(v_out + 3 * pix + c) : 0.f;
    }
    __syncthreads();
    // ---- forward: thread = pixel
    float r = 0.f, g = 0.f, b = 0.f;
    int last = -1;
    if (kHasFwd) forward_sweep(sg, cnt, blk, inside, (float)j, (float)i, r, g, b, last);
    // no intersection at all: the reference returns ones * background (== 1) and no gradient
    // (rasterize_sum_plus.py:110-118)
    if (n_isect == 0.0) r = g = b = 1.f;
    if (kMode == RasterMode::Render) {
        // the model's `render`: clamp to [0,1], CHW planar (gaussianimage_covariance.py:210-211)
        if (inside && out_img) {
            const size_t plane = (size_t)p.img_width * p.img_height;
            out_img[pix] = fminf(fmaxf(r, 0.f), 1.f);
            out_img[plane + pix] = fminf(fmaxf(g, 0.f), 1.f);
            out_img[2 * plane + pix] = fminf(fmaxf(b, 0.f), 1.f);
        }
        return;
    }
    if (kHasLoss) {
    // ---- pointwise loss (mse and/or l1, models/utils.py:64-67,74-75):
    //      d/d out = loss_scale * d + loss_l1_scale * sign(d), d = clamp(out) - gt, where 0 <= out <= 1
    //      (torch.clamp backward mask); squared error of the clamped render for PSNR
    float se = 0.f, ae = 0.f, vr = 0.f, vg = 0.f, vb = 0.f;
    if (inside) {
        const float dr = fminf(fmaxf(r, 0.f), 1.f) - tr;
        const float dg = fminf(fmaxf(g, 0.f), 1.f) - tgc;
        const float db = fminf(fmaxf(b, 0.f), 1.f) - tb;
        se = dr * dr + dg * dg + db * db;
        ae = fabsf(dr) + fabsf(dg) + fabsf(db);
        const float l1 = p.loss_l1_scale;
        vr = (r >= 0.f && r <= 1.f) ? fmaf(l1, (float)((dr > 0.f) - (dr < 0.f)), p.loss_scale * dr) : 0.f;
        vg = (g >= 0.f && g <= 1.f) ? fmaf(l1, (float)((dg > 0.f) - (dg < 0.f)), p.loss_scale * dg) : 0.f;
        vb = (b >= 0.f && b <= 1.f) ? fmaf(l1, (float)((db > 0.f) - (db < 0.f)), p.loss_scale * db) : 0.f;
        if (out_img) {
            out_img[3 * pix] = r;
            out_img[3 * pix + 1] = g;
            out_img[3 * pix + 2] = b;
        }
        // torch.abs(render - gt).sum(dim=1) of train.py:87, channel order r,g,b
        if (err_map) err_map[pix] = __fadd_rn(__fadd_rn(fabsf(dr), fabsf(dg)), fabsf(db));
    }
    const int gi = grad_index(lx, ly);
    tg.v[0][gi] = vr;
    tg.v[1][gi] = vg;
    tg.v[2][gi] = vb;
    se = warp_sum(se);
    if (p.loss_l1_scale != 0.f) ae = warp_sum(ae);
    if (lane == 0) { s_red[0][warp] = se; s_red[1][warp] = ae; }
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f, tot1 = 0.f;
#pragma unroll
        for (int w = 0; w < kRasterWarps; ++w) { tot += s_red[0][w]; tot1 += s_red[1][w]; }
        atomicAdd(stats + GI2D_STAT_SSE + (tile_id & (GI2D_STAT_SSE_SLOTS - 1)), (double)tot);
        if (p.loss_l1_scale != 0.f) atomicAdd(stats + GI2D_STAT_ABS_SUM, (double)tot1);
    }
    }  // kHasLoss
    if (!kHasBwd || cnt == 0) return;
    // ---- backward: warp = Gaussian, lane = 8 pixels
    const LanePixels lp = lane_pixels(blockIdx.x, tile_y, p.img_width, p.img_height);
    auto grad_of = [&](int gid, int k) -> float * { return grads + 8 * (size_t)gid + k; };
    // (CTA-uniform) a tile that lies wholly inside the image needs no per-pixel inside test
    if ((blockIdx.x + 1) * kTile <= p.img_width && (tile_y + 1) * kTile <= p.img_height)
        backward_tile<false, kRasterWarps, true>(sg, s_ids, cnt, lp, tg, grad_of, nullptr);
    else
        backward_tile<false, kRasterWarps, false>(sg, s_ids, cnt, lp, tg, grad_of, nullptr);
}

// ------------------------------------------------------------------------------------ K3, round 2
// The same work as fit_raster_kernel, reorganised around what the hardware issues fastest (gi2d_raster_quad.cuh):
// one CTA per tile of kWarps warps (1, 2 or 4), every warp owning 4 / kWarps quadrants of 8x8 pixels with its
// lanes holding pixel PAIRS as f32x2, so that the sweeps are packed FFMA2 / FMUL2 / FADD2; the backward works on
// four Gaussians per warp at a time (8 lanes each) over pixel rows.  With 1 or 2 warps per tile a warp keeps its
// own region from the forward to the backward sweep (dL/d(out) changes layout through a per-warp shared-memory
// copy, one __syncwarp); with 4 warps per tile the backward region is the whole tile and the groups of four
// Gaussians are dealt to the warps (one more block barrier, twice the warps to hide latency with).
// Results: tile ranges, sorted keys and the image are bit-identical to fit_raster_kernel (same operations per
// pixel, same order); gradients agree to fp32 summation order.
#ifndef GI2D_RQ4_MINBLOCKS
#define GI2D_RQ4_MINBLOCKS 7
#endif
#ifndef GI2D_RQ2_MINBLOCKS
#define GI2D_RQ2_MINBLOCKS 14
#endif
template <RasterMode kMode, int kWarps>
__global__ void __launch_bounds__(32 * kWarps, kWarps == 1 ? 16 : (kWarps == 2 ? GI2D_RQ2_MINBLOCKS : GI2D_RQ4_MINBLOCKS))
fit_rasterq_kernel(gi2d_fit_params p, uint64_t *__restrict__ sorted_keys, uint64_t *__restrict__ keys_tmp,
                   const int32_t *__restrict__ tile_bins, int32_t *__restrict__ tile_count,
                   int32_t *__restrict__ tile_fill, const float4 *__restrict__ records,
                   const float *__restrict__ gt, const uint8_t *__restrict__ gt_u8,
                   float *__restrict__ out_img, float *__restrict__ grads, double *__restrict__ stats,
                   float *__restrict__ err_map, const float *__restrict__ v_out,
                   const int4 *__restrict__ tile_work, int bucket_cap) {
    constexpr bool kHasFwd = kMode != RasterMode::FitBackward;
    constexpr bool kHasLoss = kMode == RasterMode::Fit || kMode == RasterMode::FitForward;
    constexpr bool kHasBwd = kMode == RasterMode::Fit || kMode == RasterMode::FitBackward;
    constexpr int kNQ = kQuads / kWarps;            // quadrants per warp
    constexpr int kCols = QuadGeom<kNQ>::kCols;
    constexpr int kThreads = 32 * kWarps;
    constexpr bool kTileWide = kWarps == 4;         // backward region: the whole tile, groups dealt to the warps
    constexpr int kRegions = kTileWide ? 1 : kWarps;
    constexpr int kRegionRows = kTile / kRegions;
    // 4 warps per tile: ONE list for the CTA, built cooperatively while the entries are staged (each entry's sweep
    // shape is computed once and packed; a shared-memory counter per trip-count bucket hands out list slots) instead
    // of every warp building its own copy with ballots: +1 % at 768x512, -9 % kernel time when one tile holds 170
    // entries.  (-DGI2D_NO_WIDE_LIST: the per-warp ballot lists, as the 1- and 2-warp variants use.)
#ifndef GI2D_NO_WIDE_LIST
    constexpr bool kWideList = kTileWide && (kMode == RasterMode::Fit || kMode == RasterMode::FitBackward);
#else
    constexpr bool kWideList = false;
#endif
    __shared__ QuadRecords sg;
    __shared__ int s_ids[kMaxPerTile];
    // the ids being ranked share their shared memory with dL/d(out) and the backward's lists: the rank sort is over
    // (block barrier) before either exists.  Up to kSortMax entries of a tile are ranked in shared memory.
    constexpr int kSortMax = 960;
    struct BwdShared {
        WarpGrad<kRegionRows> wg[kRegions];
        unsigned char list[kWarps][kMaxPerTile];
    };
    union SortOrBwd {
        int sort[kSortMax + 4];
        BwdShared bwd;
    };
    __shared__ __align__(16) SortOrBwd s_u;
    int *s_sort = s_u.sort;
    auto &s_wg = s_u.bwd.wg;
    auto &s_list = s_u.bwd.list;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    if (kWideList && tid < 4) sg.bucket_count[tid] = 0;   // (visible after the first block barrier of the staging)
    // the warp's first quadrant (column, row) and its bit in the reach masks
    const int qcol0 = kWarps == 4 ? (warp & 1) : 0;
    const int qrow0 = kWarps == 4 ? (warp >> 1) : (kWarps == 2 ? warp : 0);
    const int qshift = 2 * qrow0 + qcol0;
    const int lx0 = 8 * qcol0 + (lane & 7), ly0 = 8 * qrow0 + (lane >> 3);   // tile-relative
    pdl_wait();
    // which tile: the launch order (tile_work == nullptr) or the placement kernel's heaviest-first work list
    // (one entry per CTA: tile x | y << 16, range begin, range end) -- the long tiles start first, the short ones
    // fill the tail of the grid's last wave
    int tile_x, tile_y;
    int2 range;
    int32_t *zero_a = tile_count, *zero_b = tile_fill;   // the counters this CTA hands back zeroed
    // first trip of the staging loop: key and record of entry `tid` stay in registers (one round trip)
    uint64_t key0 = 0;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), b0 = a0;
    bool have0 = false;
    if (bucket_cap > 0) {
        // bucketed binning: the tile's rows start at tile * bucket_cap, its count is the cursor K1 left in the
        // counter array of this forward (stats[kStatBank]); the OTHER array is zeroed for the next forward.
        // Nothing here depends on anything else: the flag, both counters and -- speculatively, the rows exist
        // whatever the count is -- the first 32 entries are all in flight together.
        tile_x = blockIdx.x;
        tile_y = p.tile_row_begin + blockIdx.y;
        if (tile_work) {   // (bucketed: one int per CTA, tile x | y << 16, K1's heaviest-first list)
            const int wk = __ldcg(reinterpret_cast<const int32_t *>(tile_work) + blockIdx.y * gridDim.x + blockIdx.x);
            tile_x = wk & 0xffff;
            tile_y = (int)((unsigned)wk >> 16);
        }
        const int t = tile_y * p.tiles_x + tile_x;
        const double bankd = __ldcg(stats + kStatBank);
        const int c0 = __ldcg(tile_count + t), c1 = __ldcg(tile_fill + t);
        if (tid < min(bucket_cap, 32)) {
            const size_t row = (size_t)t * bucket_cap + tid;
            key0 = __ldcg(sorted_keys + row);
            a0 = __ldcg(records + 2 * row);
            b0 = __ldcg(records + 2 * row + 1);
            have0 = true;
        }
        const bool bank = bankd != 0.0;
        const int c = bank ? c1 : c0;
        range = make_int2(t * bucket_cap, t * bucket_cap + min(c, bucket_cap));
        zero_a = zero_b = bank ? tile_count : tile_fill;
    } else if (tile_work) {
        const int4 wk = __ldcg(tile_work + blockIdx.y * gridDim.x + blockIdx.x);
        tile_x = wk.x & 0xffff;
        tile_y = (int)((unsigned)wk.x >> 16);
        range = make_int2(wk.y, wk.z);
    } else {
        tile_x = blockIdx.x;
        tile_y = p.tile_row_begin + blockIdx.y;
        range = __ldcg(reinterpret_cast<const int2 *>(tile_bins) + tile_y * p.tiles_x + tile_x);
    }
    const int tile_id = tile_y * p.tiles_x + tile_x;
    const int px0 = tile_x * kTile + lx0;
    const int py0 = tile_y * kTile + ly0;
    const QuadLane<kNQ> ln = quad_lane<kNQ>(px0, py0);
    // pixel (quadrant qi, pair element e) of this lane: (px0 + 8*(qi % kCols), py0 + 8*(qi / kCols) + 4*e)
    const bool full_tile = (tile_x + 1) * kTile <= p.img_width && (tile_y + 1) * kTile <= p.img_height;
    unsigned outside = 0;
    if (!full_tile) {
#pragma unroll
        for (int qi = 0; qi < kNQ; ++qi)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (px0 + 8 * (qi % kCols) >= p.img_width || py0 + 8 * (qi / kCols) + 4 * e >= p.img_height)
                    outside |= 1u << (2 * qi + e);
    }
    // the target pixels are written by no kernel of the step: pull their lines towards L2 while the list is
    // staged (they are loaded after the forward sweep, so they cost no registers here)
#ifdef GI2D_PREFETCH_ROWS
    if (kHasLoss && tid < kTile) {   // one lane per tile row
        const int row = tile_y * kTile + tid;
        if (row < p.img_height) {
            const int bpp = gt ? 12 : 3;
            const char *base = gt ? (const char *)gt : (const char *)gt_u8;
            const char *a = base + ((size_t)row * p.img_width + tile_x * kTile) * bpp;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            if (gt) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));   // (a float row is 192 bytes)
        }
    }
#else
    if (kHasLoss) {
        const int bpp = gt ? 12 : 3;
        const char *base = gt ? (const char *)gt : (const char *)gt_u8;
#pragma unroll
        for (int qi = 0; qi < kNQ; ++qi)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (!((outside >> (2 * qi + e)) & 1u) && (lane & 7) == 0) {
                    const size_t pix = (size_t)(py0 + 8 * (qi / kCols) + 4 * e) * p.img_width + px0 + 8 * (qi % kCols);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + pix * bpp));
                }
    }
#endif
    const double n_isect = __ldcg(stats + GI2D_STAT_ISECTS);
    const int total_cnt = max(0, min(range.y, p.isect_capacity) - range.x);
    const int cnt = min(kMaxPerTile, total_cnt);
    if (kHasFwd && tid == 0) {   // the counters of this tile go back to K1 / K2 of the next step zeroed
        GI2D_CHECK(stats, range.x >= 0 && range.y >= range.x &&
                              (bucket_cap > 0 || (tile_fill[tile_id] == tile_count[tile_id] &&
                               (range.y - range.x == tile_count[tile_id] || n_isect > (double)p.isect_capacity))));
        zero_a[tile_id] = 0;
        if (zero_b != zero_a) zero_b[tile_id] = 0;
    }
    // ---- finish the key sort (see fit_raster_kernel): rank by gaussian id, stage at the rank
    constexpr bool kWriteBack = kMode != RasterMode::FitForward;
    const float tx0 = (float)(tile_x * kTile), ty0 = (float)(tile_y * kTile);
    if (total_cnt <= kMaxPerTile) {
        if (tid < cnt) {
            if (!have0) {
                key0 = __ldcg(sorted_keys + range.x + tid);
                a0 = __ldcg(records + 2 * (size_t)(range.x + tid));
                b0 = __ldcg(records + 2 * (size_t)(range.x + tid) + 1);
            }
            s_sort[tid] = (int)(uint32_t)key0;
        }
        for (int e = tid + kThreads; e < cnt; e += kThreads) s_sort[e] = (int)(uint32_t)__ldcg(sorted_keys + range.x + e);
        if (tid < 4) s_sort[cnt + tid] = 0x7fffffff;   // pads the vectorised rank loop
        __syncthreads();
        for (int e = tid; e < cnt; e += kThreads) {
            float4 r0 = a0, r1 = b0;
            if (e != tid) {
                r0 = __ldcg(records + 2 * (size_t)(range.x + e));
                r1 = __ldcg(records + 2 * (size_t)(range.x + e) + 1);
            }
            const int id = s_sort[e];
            int rank = 0;
            for (int jj = 0; jj < cnt; jj += 4) {
                const int4 o = *reinterpret_cast<const int4 *>(s_sort + jj);
                rank += (o.x < id) + (o.y < id) + (o.z < id) + (o.w < id);
            }
            GI2D_CHECK(stats, rank >= 0 && rank < cnt && id >= 0 && id < p.num_points);
            if (kWideList) stage_quad_wide(sg, rank, r0, r1, tx0, ty0); else stage_quad(sg, rank, r0, r1, tx0, ty0);
            if (kHasBwd) s_ids[rank] = id;
            // (keys are rebuilt from the tile id: nobody re-reads sorted_keys of this tile after the barrier)
            if (kWriteBack && rank != e) sorted_keys[range.x + rank] = ((uint64_t)(uint32_t)tile_id << 32) | (uint32_t)id;
        }
    } else if (total_cnt <= kSortMax) {
        // more than the 256 entries that get staged (densification clusters new Gaussians: 150-400 per tile were
        // seen at 768x512): rank ALL ids in shared memory, stage the 256 smallest (forward.cu:673 truncates the
        // sorted list), write the fully sorted keys back through the scratch array
        for (int e = tid; e < total_cnt; e += kThreads) s_sort[e] = (int)(uint32_t)__ldcg(sorted_keys + range.x + e);
        if (tid < 4) s_sort[total_cnt + tid] = 0x7fffffff;
        __syncthreads();
        for (int e = tid; e < total_cnt; e += kThreads) {
            const int id = s_sort[e];
            int rank = 0;
            for (int jj = 0; jj < total_cnt; jj += 4) {
                const int4 o = *reinterpret_cast<const int4 *>(s_sort + jj);
                rank += (o.x < id) + (o.y < id) + (o.z < id) + (o.w < id);
            }
            GI2D_CHECK(stats, rank >= 0 && rank < total_cnt && id >= 0 && id < p.num_points);
            keys_tmp[range.x + rank] = ((uint64_t)(uint32_t)tile_id << 32) | (uint32_t)id;
            if (rank < kMaxPerTile) {
                const float4 q0 = __ldcg(records + 2 * (size_t)(range.x + e)), q1 = __ldcg(records + 2 * (size_t)(range.x + e) + 1);
                if (kWideList) stage_quad_wide(sg, rank, q0, q1, tx0, ty0); else stage_quad(sg, rank, q0, q1, tx0, ty0);
                if (kHasBwd) s_ids[rank] = id;
            }
        }
        if (kWriteBack) {
            __syncthreads();
            for (int e = tid; e < total_cnt; e += kThreads) sorted_keys[range.x + e] = __ldcg(keys_tmp + range.x + e);
        }
    } else if (bucket_cap > 0) {
        // beyond that (right after a densification ~1000 new Gaussians can sit in ONE tile): only the 256 smallest
        // ids get staged (forward.cu:673 truncates the sorted list), so SELECT them instead of sorting everything --
        // an MSB-first radix select over the ids (8-bit digits, a 256-bin shared-memory histogram per digit: two
        // passes over the tile's keys for up to 65536 Gaussians), then the usual rank sort of the 256 survivors.
        // O(n) instead of the O(n^2) global-memory ranking below, which made such a step take a millisecond.  The
        // tile's keys stay unsorted in their bucket; gi2d_fit_export_binning sorts such tiles when somebody asks.
        int *sel_id = s_sort, *sel_e = s_sort + kMaxPerTile, *hist = s_sort + 2 * kMaxPerTile, *s_sel = s_sort + 3 * kMaxPerTile;
        const int nbits = 32 - __clz(max(p.num_points - 1, 1));
        unsigned prefix = 0, prefix_mask = 0;
        int k = kMaxPerTile;
        for (int shift = ((nbits - 1) / 8) * 8; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kThreads) hist[i] = 0;
            __syncthreads();
            for (int e = tid; e < total_cnt; e += kThreads) {
                const unsigned id = (unsigned)__ldcg(sorted_keys + range.x + e);
                if ((id & prefix_mask) == prefix) atomicAdd(&hist[(id >> shift) & 255u], 1);
            }
            __syncthreads();
            if (warp == 0) {   // the digit at which the running count reaches k
                int v[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = hist[8 * lane + j]; sum += v[j]; }
                const int incl = warp_scan_inclusive(sum);
                int run = incl - sum;
                if (run < k && k <= incl) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (run < k && k <= run + v[j]) { s_sel[0] = 8 * lane + j; s_sel[1] = k - run; }
                        run += v[j];
                    }
                }
            }
            __syncthreads();
            prefix |= (unsigned)s_sel[0] << shift;
            prefix_mask |= 255u << shift;
            k = s_sel[1];
        }
        if (tid == 0) s_sel[2] = 0;
        __syncthreads();
        for (int e = tid; e < total_cnt; e += kThreads) {   // ids are unique inside a tile: exactly 256 are <= prefix
            const unsigned id = (unsigned)__ldcg(sorted_keys + range.x + e);
            if (id <= prefix) {
                const int slot = atomicAdd(&s_sel[2], 1);
                if (slot < kMaxPerTile) { sel_id[slot] = (int)id; sel_e[slot] = e; }
            }
        }
        __syncthreads();
        GI2D_CHECK(stats, s_sel[2] == kMaxPerTile);
        for (int e2 = tid; e2 < kMaxPerTile; e2 += kThreads) {
            const int id = sel_id[e2];
            int rank = 0;
            for (int jj = 0; jj < kMaxPerTile; jj += 4) {
                const int4 o = *reinterpret_cast<const int4 *>(sel_id + jj);
                rank += (o.x < id) + (o.y < id) + (o.z < id) + (o.w < id);
            }
            const size_t src = (size_t)(range.x + sel_e[e2]);
            const float4 q0 = __ldcg(records + 2 * src), q1 = __ldcg(records + 2 * src + 1);
            if (kWideList) stage_quad_wide(sg, rank, q0, q1, tx0, ty0); else stage_quad(sg, rank, q0, q1, tx0, ty0);
            if (kHasBwd) s_ids[rank] = id;
        }
    } else {
        // beyond that on the scan + placement path (whose compact key array is the step's public output): full
        // rank sort straight from global memory
        if (kWideList) __syncthreads();   // (the bucket counters were zeroed by four threads)
        for (int e = tid; e < total_cnt; e += kThreads) {
            const uint64_t key = __ldcg(sorted_keys + range.x + e);
            const int id = (int)(uint32_t)key;
            int rank = 0;
            for (int jj = 0; jj < total_cnt; ++jj) rank += ((int)(uint32_t)__ldcg(sorted_keys + range.x + jj) < id) ? 1 : 0;
            GI2D_CHECK(stats, rank >= 0 && rank < total_cnt && (int)(key >> 32) == tile_id);
            keys_tmp[range.x + rank] = key;
            if (rank < kMaxPerTile) {
                const float4 q0 = __ldcg(records + 2 * (size_t)(range.x + e)), q1 = __ldcg(records + 2 * (size_t)(range.x + e) + 1);
                if (kWideList) stage_quad_wide(sg, rank, q0, q1, tx0, ty0); else stage_quad(sg, rank, q0, q1, tx0, ty0);
                if (kHasBwd) s_ids[rank] = id;
            }
        }
        if (kWriteBack) {
            __syncthreads();
            for (int e = tid; e < total_cnt; e += kThreads) sorted_keys[range.x + e] = __ldcg(keys_tmp + range.x + e);
        }
    }
    __syncthreads();
    int n_wide = 0;
    if (kWideList) n_wide = finish_wide_list(sg, cnt, s_list[0]);   // (read after the barrier that precedes the backward)
    // ---- forward: lane = pixel pairs of the warp's quadrants
    f32x2 accR[kNQ], accG[kNQ], accB[kNQ];
#pragma unroll
    for (int qi = 0; qi < kNQ; ++qi) accR[qi] = accG[qi] = accB[qi] = 0ull;
    if (kHasFwd) quad_forward<kNQ>(sg, cnt, ln, qshift, accR, accG, accB);
    // no intersection in the whole image: the reference returns ones * background (== 1) and no gradient
    // (rasterize_sum_plus.py:110-118); a band of a tile-row split cannot know, and renders its (empty) sum
    const bool ones = n_isect == 0.0 && p.tile_row_begin == 0 && p.tile_row_end == p.tiles_y;
    // the region whose dL/d(out) this warp fills: its own rows (1, 2 warps per tile) or the whole tile (4)
    WarpGrad<kRegionRows> &wg = s_wg[kTileWide ? 0 : warp];
    const int region_row0 = kTileWide ? 0 : 8 * qrow0;
    float se = 0.f, ae = 0.f;
#pragma unroll
    for (int qi = 0; qi < kNQ; ++qi) {
        float cr[2], cg[2], cb[2];
        unpk2(accR[qi], cr[0], cr[1]);
        unpk2(accG[qi], cg[0], cg[1]);
        unpk2(accB[qi], cb[0], cb[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const bool inside = !((outside >> (2 * qi + e)) & 1u);
            const int lx = lx0 + 8 * (qi % kCols), ly = ly0 + 8 * (qi / kCols) + 4 * e;   // tile-relative
            const size_t pix = (size_t)(tile_y * kTile + ly) * p.img_width + tile_x * kTile + lx;
            float r = cr[e], g = cg[e], b = cb[e];
            if (ones) r = g = b = 1.f;
            float wr = 0.f, wgc = 0.f, wb = 0.f;
            if (kMode == RasterMode::Render) {
                // the model's `render`: clamp to [0,1], CHW planar (gaussianimage_covariance.py:210-211)
                if (inside && out_img) {
                    const size_t plane = (size_t)p.img_width * p.img_height;
                    out_img[pix] = fminf(fmaxf(r, 0.f), 1.f);
                    out_img[plane + pix] = fminf(fmaxf(g, 0.f), 1.f);
                    out_img[2 * plane + pix] = fminf(fmaxf(b, 0.f), 1.f);
                }
            } else if (kHasLoss) {
                // pointwise loss (mse and/or l1, models/utils.py:64-67,74-75): d/d out = loss_scale * d +
                // loss_l1_scale * sign(d), d = clamp(out) - gt, where 0 <= out <= 1 (torch.clamp backward mask)
                if (inside) {
                    float tr, tg, tb;
                    if (gt) {
                        tr = __ldg(gt + 3 * pix); tg = __ldg(gt + 3 * pix + 1); tb = __ldg(gt + 3 * pix + 2);
                    } else {   // 8-bit target: the value torchvision's ToTensor produces, u8 / 255
                        tr = u8_to_unit(__ldg(gt_u8 + 3 * pix));
                        tg = u8_to_unit(__ldg(gt_u8 + 3 * pix + 1));
                        tb = u8_to_unit(__ldg(gt_u8 + 3 * pix + 2));
                    }
                    const float kr = fminf(fmaxf(r, 0.f), 1.f), kg = fminf(fmaxf(g, 0.f), 1.f), kb = fminf(fmaxf(b, 0.f), 1.f);
                    const float dr = kr - tr, dg = kg - tg, db = kb - tb;
                    // torch.clamp's backward mask 0 <= out <= 1 is "the clamp changed nothing" (false for NaN too)
                    const bool mr = kr == r, mg = kg == g, mb = kb == b;
                    se += dr * dr + dg * dg + db * db;
                    const float l1 = p.loss_l1_scale;
                    if (l1 != 0.f) {
                        ae += fabsf(dr) + fabsf(dg) + fabsf(db);
                        wr = mr ? fmaf(l1, (float)((dr > 0.f) - (dr < 0.f)), p.loss_scale * dr) : 0.f;
                        wgc = mg ? fmaf(l1, (float)((dg > 0.f) - (dg < 0.f)), p.loss_scale * dg) : 0.f;
                        wb = mb ? fmaf(l1, (float)((db > 0.f) - (db < 0.f)), p.loss_scale * db) : 0.f;
                    } else {
                        wr = mr ? p.loss_scale * dr : 0.f;
                        wgc = mg ? p.loss_scale * dg : 0.f;
                        wb = mb ? p.loss_scale * db : 0.f;
                    }
                    if (out_img) {
                        out_img[3 * pix] = r;
                        out_img[3 * pix + 1] = g;
                        out_img[3 * pix + 2] = b;
                    }
                    // torch.abs(render - gt).sum(dim=1) of train.py:87, channel order r,g,b
                    if (err_map) err_map[pix] = __fadd_rn(__fadd_rn(fabsf(dr), fabsf(dg)), fabsf(db));
                }
            } else {   // FitBackward: dL/d(out) computed by the loss kernels between the two halves
                if (inside && n_isect != 0.0) {
                    wr = __ldcg(v_out + 3 * pix);
                    wgc = __ldcg(v_out + 3 * pix + 1);
                    wb = __ldcg(v_out + 3 * pix + 2);
                }
            }
            if (kHasBwd) {   // hand dL/d(out) to the backward's layout (0 outside the image)
                wg.v[0][ly - region_row0][lx] = wr;
                wg.v[1][ly - region_row0][lx] = wgc;
                wg.v[2][ly - region_row0][lx] = wb;
            }
        }
    }
    if (kMode == RasterMode::Render) return;
    if (kHasLoss) {
        se = warp_sum(se);
        if (p.loss_l1_scale != 0.f) ae = warp_sum(ae);
        if (lane == 0) {
            atomicAdd(stats + GI2D_STAT_SSE + ((tile_id * kWarps + warp) & (GI2D_STAT_SSE_SLOTS - 1)), (double)se);
            if (p.loss_l1_scale != 0.f) atomicAdd(stats + GI2D_STAT_ABS_SUM, (double)ae);
        }
    }
    if (!kHasBwd || cnt == 0) return;   // (CTA-uniform)
    // ---- backward: four Gaussians per warp at a time (gi2d_raster_quad.cuh, quad_backward4)
    // the staged Gaussians that can reach the region, ascending (every warp builds the list it walks)
    unsigned char *list = s_list[kWideList ? 0 : warp];
    const unsigned region_bits = kTileWide ? 0xFu : (((1u << kNQ) - 1u) << qshift);
    const int n = kWideList ? n_wide : build_group_list<kRegionRows>(sg, cnt, region_bits, region_row0, list);
    if (lane < kTile && (!kTileWide || warp == 0)) {   // the all-zero row a group reads past its own rows
        wg.v[0][kRegionRows][lane] = 0.f;
        wg.v[1][kRegionRows][lane] = 0.f;
        wg.v[2][kRegionRows][lane] = 0.f;
    }
    if (kTileWide) __syncthreads(); else __syncwarp();   // dL/d(out) and the list are in shared memory
    if (n == 0) return;
    const int region_py0 = tile_y * kTile + region_row0;
    const int first_group = kTileWide ? warp : 0, group_stride = kTileWide ? kWarps : 1;
    if (full_tile)
        quad_backward4<kRegionRows, false, kWideList>(sg, s_ids, list, n, first_group, group_stride, tile_x * kTile,
                                           region_py0, region_row0, p.img_width, kRegionRows, wg, grads);
    else
        quad_backward4<kRegionRows, true, kWideList>(sg, s_ids, list, n, first_group, group_stride, tile_x * kTile,
                                          region_py0, region_row0, p.img_width,
                                          min(kRegionRows, p.img_height - region_py0), wg, grads);
}

// which rasterizer a step launches: GI2D_RASTER=0 the round-1 kernel (8 warps per tile, scalar math); 1 / 2 / 4 the
// quadrant kernel with that many warps per tile.  Default (measured on the B200, profiles/README.md): 4 warps per
// tile while the whole grid fits the GPU in about one wave (768x512: 1536 tiles -- the extra warps hide latency),
// 2 warps per tile beyond (2040x1356, 8192^2: fewer instructions per pair win once there are many waves).
int raster_variant(int num_tiles) {
    static const int forced = [] {
        const char *e = getenv("GI2D_RASTER");
        return e ? atoi(e) : -1;
    }();
    if (forced >= 0) return forced;
    return num_tiles <= 4096 ? 4 : 2;
}

template <RasterMode kMode>
cudaError_t launch_raster(bool pdl, dim3 grid, cudaStream_t st, const gi2d_fit_params &p, uint64_t *sorted_keys,
                          uint64_t *keys_tmp, const int32_t *tile_bins, int32_t *tile_count, int32_t *tile_fill,
                          const float4 *records, const float *gt, const uint8_t *gt_u8, float *out_img, float *grads,
                          double *stats, float *err_map, const float *v_out, const int4 *tile_work, int bucket_cap) {
    const int v = raster_variant((int)(grid.x * grid.y));
#define GI2D_RASTER_ARGS p, sorted_keys, keys_tmp, tile_bins, tile_count, tile_fill, records, gt, gt_u8, out_img, grads, stats, err_map, v_out
    if (v == 1) {
        if (pdl) return launch_pdl(fit_rasterq_kernel<kMode, 1>, grid, dim3(32), 0, st, GI2D_RASTER_ARGS, tile_work, bucket_cap);
        fit_rasterq_kernel<kMode, 1><<<grid, 32, 0, st>>>(GI2D_RASTER_ARGS, tile_work, bucket_cap);
    } else if (v == 2) {
        if (pdl) return launch_pdl(fit_rasterq_kernel<kMode, 2>, grid, dim3(64), 0, st, GI2D_RASTER_ARGS, tile_work, bucket_cap);
        fit_rasterq_kernel<kMode, 2><<<grid, 64, 0, st>>>(GI2D_RASTER_ARGS, tile_work, bucket_cap);
    } else if (v == 4) {
        if (pdl) return launch_pdl(fit_rasterq_kernel<kMode, 4>, grid, dim3(128), 0, st, GI2D_RASTER_ARGS, tile_work, bucket_cap);
        fit_rasterq_kernel<kMode, 4><<<grid, 128, 0, st>>>(GI2D_RASTER_ARGS, tile_work, bucket_cap);
    } else {
        if (pdl) return launch_pdl(fit_raster_kernel<kMode>, grid, dim3(kRasterThreads), 0, st, GI2D_RASTER_ARGS);
        fit_raster_kernel<kMode><<<grid, kRasterThreads, 0, st>>>(GI2D_RASTER_ARGS);
    }
#undef GI2D_RASTER_ARGS
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------ K5
// Stand-alone optimiser launch: applies a pending gradient NOW (before the host reads or edits the
// parameters, renders, or all the steps are done).  In the steady state it is never launched: the
// next step's K1 does the same work.
__global__ void __launch_bounds__(256)
fit_adam_kernel(gi2d_fit_params p, AdamPtrs a, const float *__restrict__ cov_bound, const float4 *__restrict__ proj,
                const float4 *__restrict__ grads, double *__restrict__ stats, float4 *__restrict__ best,
                float *__restrict__ best_bound) {
    __shared__ int s_best;
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int n_live = live_points(p, stats);
    const bool pending = __ldcg(stats + kStatPending) != 0.0;
    const bool veto = __ldcg(stats + GI2D_STAT_OVERFLOW) != 0.0;
    best_flag_warp0(stats, best != nullptr && pending && !veto, &s_best);
    __syncthreads();
    bool bad = false;
    if (g < n_live) {
        float c[3];
        if (pending) {
            float2 x;
            float q[3];
            adam_update_gaussian(p, a, g, proj, grads, stats, veto, x, c, q);
            if (s_best) {
                best[2 * g] = make_float4(x.x, x.y, c[0], c[1]);
                best[2 * g + 1] = make_float4(c[2], q[0], q[1], q[2]);
                if (best_bound) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) best_bound[3 * g + k] = __ldcg(cov_bound + 3 * g + k);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) c[k] = a.cov[3 * g + k];
        }
        // check_non_semi_definite (gaussianimage_covariance.py:373-382) on cov + bound, torch's op order
        const float sx = __fadd_rn(c[0], __ldcg(cov_bound + 3 * g)), sxy = __fadd_rn(c[1], __ldcg(cov_bound + 3 * g + 1));
        const float sy = __fadd_rn(c[2], __ldcg(cov_bound + 3 * g + 2));
        const float det = __fsub_rn(__fmul_rn(sx, sy), __fmul_rn(sxy, sxy));
        bad = !((det > 0.f) && (sx > 0.f) && (sy > 0.f));
    }
    const int nbad = __syncthreads_count(bad);
    if (threadIdx.x == 0 && nbad) atomicAdd(stats + kStatNonPsdAcc, (double)nbad);
}

// second half of the flush: only after EVERY CTA of fit_adam_kernel has read the flags may they change
__global__ void fit_clear_pending_kernel(double *__restrict__ stats) {
    best_commit_warp0(stats);
    __syncwarp();
    if (threadIdx.x == 0) {
        stats[kStatPending] = 0.0;
        stats[GI2D_STAT_NON_PSD] = stats[kStatNonPsdAcc];
        stats[kStatNonPsdAcc] = 0.0;
    }
}

// --------------------------------------------------------------------- multi-GPU tile-row split
// ONE image over `world` GPUs of an NVSwitch box (the reference is single-GPU, train.py:39).
// Rank q rasterizes the tile rows [band_edge[q], band_edge[q+1]); the Gaussians are dealt to OWNERS in equal
// contiguous slices, and only the owner keeps a Gaussian's parameters and Adam moments and projects it
// (sharded optimiser AND sharded projection -- nothing per-Gaussian is replicated).  What crosses NVLink, per
// Gaussian and step, goes only between its owner and the ranks whose band its tile box overlaps (1.1 ranks on
// average, not world-1): the partial gradient row (32 B, P2P load by the owner) and the projected record +
// tile box (40 B, P2P store by the owner).  One step on a rank is
//   tr_count_kernel      wait "every peer's records of the previous exchange have landed" (flag spin on
//                        this rank's own memory) | every Gaussian's tile box clipped to the band -> per-tile
//                        overlap counts, zeroing of the gradient rows the band will touch
//   [device-wide scan] + fit_place_kernel + rasterizer on the band     (the single-GPU kernels, unchanged)
//   tr_exchange_kernel   wait "every peer finished its backward" | owner: sum the partial rows of the ranks
//                        the box overlaps (fixed rank order), projection backward + Adam, projection, and
//                        scatter of the new record + box to the ranks the old or the new box overlaps
// Cross-GPU ordering uses two monotonically increasing flag words per peer in peer-mapped memory (written with
// st.release.sys after a system fence by the first CTA of the waiting kernel's own grid, read with
// ld.acquire.sys): no separate barrier launches, no host involvement, the whole step is graph-capturable (the
// epoch lives in device memory).  A step in which ANY rank overflowed its intersection buffers applies no Adam
// update anywhere (the overflow bit travels with the flag), so the ranks never diverge.
constexpr int kTrFlagA = 0, kTrFlagB = GI2D_MAX_RANKS;          // offsets in a rank's flag block (u32[16])
constexpr int kCtrlEpoch = 0, kCtrlTicketEx = 1, kCtrlExitEx = 2, kCtrlError = 3, kCtrlTicketCnt = 4;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Thread 0 of every CTA: the first CTA of the grid to arrive tells every peer `value` (slot `rank` of their
// flag group), then everybody waits until every peer's slot in OUR flag block has reached `want`.  Returns the
// OR of the low bits of the peers' words when `low_bit` (the overflow veto).  Gives up after 20 s (a dead peer
// must not hang the GPU): ctrl[kCtrlError] is set and the host raises.
__device__ __forceinline__ uint32_t tr_signal_and_wait(const gi2d_tilerow &tr, uint32_t *ctrl, int ticket_slot,
                                                       int group, uint32_t value, uint32_t want, bool low_bit) {
    uint32_t acc = 0;
    const unsigned ticket = atomicAdd(ctrl + ticket_slot, 1u);
    if (ticket == 0) {
        __threadfence_system();
        for (int q = 0; q < tr.world; ++q)
            if (q != tr.rank) st_release_sys(tr.peer_flags[q] + group + tr.rank, value);
    }
    if (ticket == gridDim.x - 1) ctrl[ticket_slot] = 0;   // (the next launch of this kernel starts from 0)
    const unsigned long long t0 = global_ns();
    for (int q = 0; q < tr.world; ++q) {
        if (q == tr.rank) continue;
        const uint32_t *f = tr.peer_flags[tr.rank] + group + q;
        uint32_t v = ld_acquire_sys(f);
        while ((low_bit ? (v >> 1) : v) < want) {
            __nanosleep(64);
            if (global_ns() - t0 > 20000000000ull) { atomicExch(ctrl + kCtrlError, 1u); break; }
            v = ld_acquire_sys(f);
        }
        acc |= v & 1u;
    }
    return acc;
}

__device__ __forceinline__ bool box_hits_band(ushort4 bx, int y_begin, int y_end) {
    return bx.z > bx.x && (int)bx.w > y_begin && (int)bx.y < y_end;
}

// First kernel of a tile-row step (see above).  boxes_all: this rank's copy of every Gaussian's tile box
// (written by the owners); boxes_band: the box clipped to the band, for fit_place_kernel.
__global__ void __launch_bounds__(256)
tr_count_kernel(gi2d_fit_params p, gi2d_tilerow tr, const ushort4 *__restrict__ boxes_all,
                ushort4 *__restrict__ boxes_band, int32_t *__restrict__ tile_count, float4 *__restrict__ grads,
                int with_backward) {
    pdl_wait();
    if (tr.sync && threadIdx.x == 0) {
        const uint32_t epoch = *(volatile uint32_t *)(tr.ctrl + kCtrlEpoch);
        // every peer's exchange of the previous epoch has landed in our proj / boxes (and has read our grads)
        tr_signal_and_wait(tr, tr.ctrl, kCtrlTicketCnt, kTrFlagB, epoch - 1, epoch - 1, false);
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.num_points) return;
    const ushort4 bx = __ldcg(boxes_all + g);   // (peer-written: never through the non-coherent path)
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (box_hits_band(bx, p.tile_row_begin, p.tile_row_end)) {
        x0 = bx.x; x1 = bx.z;
        y0 = max((int)bx.y, p.tile_row_begin);
        y1 = min((int)bx.w, p.tile_row_end);
        if (with_backward) {   // the rows this band's backward accumulates into (and the owner will read)
            grads[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
            grads[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    boxes_band[g] = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx) atomicAdd(tile_count + ty * p.tiles_x + tx, 1);
}

// sync == 0 (ranks emulated on one GPU, ordered by the stream): publish the overflow bit of this rank's band
// step to the peers at the end of phase 1, so that the veto is global there too
__global__ void tr_publish_kernel(gi2d_tilerow tr, const double *__restrict__ stats) {
    const uint32_t epoch = *(volatile uint32_t *)(tr.ctrl + kCtrlEpoch);
    const uint32_t v = epoch * 2u + (__ldcg(stats + GI2D_STAT_OVERFLOW) != 0.0 ? 1u : 0u);
    for (int q = 0; q < tr.world; ++q)
        if (q != tr.rank) tr.peer_flags[q][kTrFlagA + tr.rank] = v;
}

// Last kernel of a tile-row step: exchange + optimiser + projection of the owned slice.  mode 0: a training
// step; mode 1: initial projection only (no gradient, no Adam, no flags: the host barriers around it).
__global__ void __launch_bounds__(256)
tr_exchange_kernel(gi2d_fit_params p, gi2d_tilerow tr, AdamPtrs a, const float *__restrict__ cov_bound,
                   double *__restrict__ stats, int mode) {
    __shared__ uint32_t s_veto;
    __shared__ float s_step_size, s_bc2;
    uint32_t *ctrl = tr.ctrl;
    uint32_t epoch = 0;
    if (threadIdx.x == 0) {
        epoch = *(volatile uint32_t *)(ctrl + kCtrlEpoch);
        uint32_t veto = __ldcg(stats + GI2D_STAT_OVERFLOW) != 0.0 ? 1u : 0u;
        if (mode == 0 && tr.sync)   // every peer has finished the backward of this epoch
            veto |= tr_signal_and_wait(tr, ctrl, kCtrlTicketEx, kTrFlagA, epoch * 2u + veto, epoch, true);
        else if (mode == 0)         // emulated ranks: tr_publish_kernel of every rank ran before (stream order)
            for (int q = 0; q < tr.world; ++q)
                if (q != tr.rank) veto |= *(volatile uint32_t *)(tr.peer_flags[tr.rank] + kTrFlagA + q) & 1u;
        s_veto = veto;
        // torch.optim.Adam's scalars for step t = STEP + 1, in double like torch; every CTA of every rank
        // evaluates the same expression (the counters themselves advance once, at the end of this kernel)
        const double b1 = __ldcg(stats + kStatB1Pow) * (double)p.beta1, b2 = __ldcg(stats + kStatB2Pow) * (double)p.beta2;
        const long long k = (long long)__ldcg(stats + GI2D_STAT_STEP);   // = t - 1
        double lr = __ldcg(stats + GI2D_STAT_LR);
        if (k > 0 && p.lr_step_size > 0 && k % p.lr_step_size == 0) lr *= (double)p.lr_gamma;
        s_step_size = (float)(lr / (1.0 - b1));
        s_bc2 = (float)sqrt(1.0 - b2);
    }
    __syncthreads();
    const bool veto = s_veto != 0;
    const int g = tr.own_begin + blockIdx.x * 256 + threadIdx.x;
    if (g < tr.own_end) {
        float4 *proj_own = reinterpret_cast<float4 *>(tr.peer_proj[tr.rank]);
        ushort4 *boxes_own = reinterpret_cast<ushort4 *>(tr.peer_boxes[tr.rank]);
        const ushort4 box_old = mode == 0 ? boxes_own[g] : make_ushort4(0, 0, 0, 0);
        float2 x;
        float c[3], q[3];
        if (mode == 0 && !veto) {
            // reduce: the partial rows of the ranks whose band the box overlapped, in rank order (so that the
            // sum does not depend on which rank owns the Gaussian)
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            for (int r = 0; r < tr.world; ++r) {
                if (!box_hits_band(box_old, tr.band_edge[r], tr.band_edge[r + 1])) continue;
                const float4 *gp = reinterpret_cast<const float4 *>(tr.peer_grads[r]);
                const float4 a0 = __ldcg(gp + 2 * g), a1 = __ldcg(gp + 2 * g + 1);
                s0.x += a0.x; s0.y += a0.y; s0.z += a0.z; s0.w += a0.w;
                s1.x += a1.x; s1.y += a1.y; s1.z += a1.z; s1.w += a1.w;
            }
            AdamRegs r;
            adam_load(a, g, proj_own, proj_own, r);   // (second pointer: placeholder, the gradient is set below)
            r.g0 = s0;
            r.g1 = s1;
            adam_apply_scalars(p, a, g, r, s_step_size, s_bc2);
            x = r.x;
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = r.c[k]; q[k] = r.q[k]; }
        } else {
            x = reinterpret_cast<const float2 *>(a.xyz)[g];
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = a.cov[3 * g + k]; q[k] = a.rgb[3 * g + k]; }
        }
        // projection of the (new) parameters, exactly as fit_project_kernel
        const float sx = __fadd_rn(c[0], __ldg(cov_bound + 3 * g));
        const float sxy = __fadd_rn(c[1], __ldg(cov_bound + 3 * g + 1));
        const float sy = __fadd_rn(c[2], __ldg(cov_bound + 3 * g + 2));
        float cr = q[0], cg = q[1], cb = q[2];
        if (p.color_sigmoid) { cr = sigmoidf(cr); cg = sigmoidf(cg); cb = sigmoidf(cb); }
        const Projected pr = project_cov(x.x, x.y, sx, sxy, sy, p.clip_coe, p.radius_clip, p.tiles_x, p.tiles_y);
        ushort4 box_new = make_ushort4(0, 0, 0, 0);
        if (pr.ntiles > 0 && !((float)pr.radius < p.radius_clip))
            box_new = make_ushort4((unsigned short)pr.box.x0, (unsigned short)pr.box.y0, (unsigned short)pr.box.x1,
                                   (unsigned short)pr.box.y1);
        const float4 rec0 = make_float4(pr.x, pr.y, pr.a, pr.b), rec1 = make_float4(pr.c, cr, cg, cb);
        // scatter: to every rank that saw the Gaussian in its band or will see it (so that a rank it leaves
        // learns that it left), and to ourselves
        for (int r = 0; r < tr.world; ++r) {
            if (r != tr.rank && mode == 0 && !box_hits_band(box_old, tr.band_edge[r], tr.band_edge[r + 1]) &&
                !box_hits_band(box_new, tr.band_edge[r], tr.band_edge[r + 1]))
                continue;
            float4 *pp = reinterpret_cast<float4 *>(tr.peer_proj[r]);
            pp[2 * g] = rec0;
            pp[2 * g + 1] = rec1;
            reinterpret_cast<ushort4 *>(tr.peer_boxes[r])[g] = box_new;
        }
    }
    if (mode != 0) return;
    // the last CTA to finish advances the epoch (every CTA has read it) and, unless the step was vetoed, the
    // optimiser's counters
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(ctrl + kCtrlExitEx, 1u) == gridDim.x - 1) {
            ctrl[kCtrlExitEx] = 0;
            if (!veto) {
                const double step = stats[GI2D_STAT_STEP] + 1.0;
                const long long k = (long long)step - 1;
                stats[GI2D_STAT_STEP] = step;
                stats[kStatB1Pow] *= (double)p.beta1;
                stats[kStatB2Pow] *= (double)p.beta2;
                if (k > 0 && p.lr_step_size > 0 && k % p.lr_step_size == 0) stats[GI2D_STAT_LR] *= (double)p.lr_gamma;
            }
            stats[GI2D_STAT_OVERFLOW] = veto ? 1.0 : 0.0;   // any rank's overflow vetoes (and is visible on) all
            __threadfence();
            *(volatile uint32_t *)(ctrl + kCtrlEpoch) = *(volatile uint32_t *)(ctrl + kCtrlEpoch) + 1u;
        }
    }
}

// d loss / d (the step's inputs) for callers with their own optimiser: projection backward only
__global__ void __launch_bounds__(256)
fit_input_grads_kernel(int n, const float4 *__restrict__ proj, const float4 *__restrict__ grads,
                       float4 *__restrict__ out, const double *__restrict__ stats, int dynamic) {
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= n) return;
    if (__ldcg(stats + GI2D_STAT_OVERFLOW) != 0.0 || (dynamic && g >= (int)__ldcg(stats + GI2D_STAT_NUM_POINTS))) {   // truncated list: the caller's optimiser step becomes a no-op
        out[2 * g] = out[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4 p0 = __ldcg(proj + 2 * g), p1 = __ldcg(proj + 2 * g + 1);
    const float4 g0 = __ldcg(grads + 2 * g), g1 = __ldcg(grads + 2 * g + 1);
    float gc[3];
    conic_vjp(p0.z, p0.w, p1.x, g0.z, g0.w, g1.x, gc[0], gc[1], gc[2]);
    out[2 * g] = make_float4(g0.x, g0.y, gc[0], gc[1]);
    out[2 * g + 1] = make_float4(gc[2], g1.y, g1.z, g1.w);
}

__global__ void fit_reset_kernel(gi2d_fit_params p, double *stats, int step) {
    const int i = threadIdx.x;
    if (i >= GI2D_STAT_COUNT) return;
    if (i == kStatBank) return;        // which counter array the last forward filled: not a statistic either
    if (i == GI2D_STAT_NUM_POINTS) {   // the live count is not a statistic: kept when the device owns it
        if (!p.dynamic_points) stats[i] = (double)p.num_points;
        return;
    }
    double v = 0.0;
    if (i == GI2D_STAT_STEP) v = (double)step;
    if (i == GI2D_STAT_BEST_SSE) v = __longlong_as_double(0x7ff0000000000000LL);  // +inf: nothing seen yet
    if (i == kStatB1Pow) v = pow((double)p.beta1, (double)step);
    if (i == kStatB2Pow) v = pow((double)p.beta2, (double)step);
    // lr the NEXT step will start from: lr0 * gamma^floor((step-1)/size) for step >= 1
    if (i == GI2D_STAT_LR)
        v = (double)p.lr0 * pow((double)p.lr_gamma,
                                (step >= 1 && p.lr_step_size > 0) ? floor((double)(step - 1) / p.lr_step_size) : 0.0);
    stats[i] = v;
}

// ------------------------------------------------------------------- prune / densify on the device
//   The model's size changes in place: the per-Gaussian arrays hold `capacity` rows, the live
// count is stats[GI2D_STAT_NUM_POINTS], and these kernels move rows and the count -- no reallocation, no host
// round trip, the captured step graph stays valid.
struct ModelPtrs {
    float *xyz, *cov, *rgb, *bound, *m_xyz, *v_xyz, *m_cov, *v_cov, *m_rgb, *v_rgb;
};
constexpr int kRowFloats = 27;   // xyz 2 + cov 3 + rgb 3 + bound 3 + m (2+3+3) + v (2+3+3)

__device__ __forceinline__ void load_row(const ModelPtrs &m, int g, float (&r)[kRowFloats]) {
    r[0] = m.xyz[2 * g]; r[1] = m.xyz[2 * g + 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r[2 + k] = m.cov[3 * g + k]; r[5 + k] = m.rgb[3 * g + k]; r[8 + k] = m.bound[3 * g + k];
        r[13 + k] = m.m_cov[3 * g + k]; r[16 + k] = m.m_rgb[3 * g + k];
        r[21 + k] = m.v_cov[3 * g + k]; r[24 + k] = m.v_rgb[3 * g + k];
    }
    r[11] = m.m_xyz[2 * g]; r[12] = m.m_xyz[2 * g + 1];
    r[19] = m.v_xyz[2 * g]; r[20] = m.v_xyz[2 * g + 1];
}

__device__ __forceinline__ void store_row(const ModelPtrs &m, int g, const float (&r)[kRowFloats]) {
    m.xyz[2 * g] = r[0]; m.xyz[2 * g + 1] = r[1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        m.cov[3 * g + k] = r[2 + k]; m.rgb[3 * g + k] = r[5 + k]; m.bound[3 * g + k] = r[8 + k];
        m.m_cov[3 * g + k] = r[13 + k]; m.m_rgb[3 * g + k] = r[16 + k];
        m.v_cov[3 * g + k] = r[21 + k]; m.v_rgb[3 * g + k] = r[24 + k];
    }
    m.m_xyz[2 * g] = r[11]; m.m_xyz[2 * g + 1] = r[12];
    m.v_xyz[2 * g] = r[19]; m.v_xyz[2 * g + 1] = r[20];
}

// check_non_semi_definite (gaussianimage_covariance.py:373-382) on (sxx, sxy, syy), torch's operation order
__device__ __forceinline__ bool is_pos_def(float sx, float sxy, float sy) {
    const float det = __fsub_rn(__fmul_rn(sx, sy), __fmul_rn(sxy, sxy));
    return (det > 0.f) && (sx > 0.f) && (sy > 0.f);
}

// keep[g] = 1 for the live rows whose covariance + bound is positive definite
__global__ void __launch_bounds__(256)
prune_flag_kernel(gi2d_fit_params p, ModelPtrs m, const double *__restrict__ stats, int32_t *__restrict__ keep) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= p.num_points) return;
    int k = 0;
    if (g < live_points(p, stats))
        k = is_pos_def(__fadd_rn(m.cov[3 * g], m.bound[3 * g]), __fadd_rn(m.cov[3 * g + 1], m.bound[3 * g + 1]),
                       __fadd_rn(m.cov[3 * g + 2], m.bound[3 * g + 2])) ? 1 : 0;
    keep[g] = k;
}

// stable compaction, first half: survivors to their new row of the scratch copy
__global__ void __launch_bounds__(256)
prune_gather_kernel(gi2d_fit_params p, ModelPtrs m, const double *__restrict__ stats, const int32_t *__restrict__ keep,
                    const int32_t *__restrict__ incl, float *__restrict__ scratch) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int n = live_points(p, stats);
    const int kept = p.num_points > 0 ? incl[p.num_points - 1] : 0;
    if (kept == n || kept == 0) return;    // nothing to prune / `cur - to_prune > 0` fails: leave everything
    if (g >= n || !keep[g]) return;
    float r[kRowFloats];
    load_row(m, g, r);
    float *dst = scratch + (size_t)(incl[g] - 1) * kRowFloats;
#pragma unroll
    for (int k = 0; k < kRowFloats; ++k) dst[k] = r[k];
}

// second half: scratch back to the arrays, new live count
__global__ void __launch_bounds__(256)
prune_scatter_kernel(gi2d_fit_params p, ModelPtrs m, double *__restrict__ stats, const int32_t *__restrict__ incl,
                     const float *__restrict__ scratch) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    const int n = live_points(p, stats);
    const int kept = p.num_points > 0 ? incl[p.num_points - 1] : 0;
    const bool moved = !(kept == n || kept == 0);
    if (moved && g < kept) {
        float r[kRowFloats];
        const float *src = scratch + (size_t)g * kRowFloats;
#pragma unroll
        for (int k = 0; k < kRowFloats; ++k) r[k] = src[k];
        store_row(m, g, r);
    }
    // (every thread has read the old count above; the last block to finish publishes the new one)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd((unsigned *)(stats + GI2D_STAT_COUNT - 1), 1u) == gridDim.x - 1;   // ticket in the last (unused) slot
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *(unsigned *)(stats + GI2D_STAT_COUNT - 1) = 0u;
        stats[GI2D_STAT_PRUNED] = moved ? (double)(n - kept) : 0.0;
        if (moved && p.dynamic_points) stats[GI2D_STAT_NUM_POINTS] = (double)kept;
    }
}

// error map -> sortable keys: descending error, ties by ascending pixel index
__global__ void __launch_bounds__(256)
densify_keys_kernel(int n_pix, const float *__restrict__ err, int64_t *__restrict__ keys, int32_t *__restrict__ vals) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_pix) return;
    const float e = __ldcg(err + i);
    uint32_t bits = __float_as_uint(e);
    if (!(e >= 0.f)) bits = 0;                      // (NaN / negative cannot occur in a sum of absolute values)
    if (bits > 0x7F800000u) bits = 0x7F800000u;
    keys[i] = ((int64_t)(0x7FFFFFFFu - bits) << 32) | (uint32_t)i;
    vals[i] = i;
}

// the first k sorted candidates become new Gaussians (train.py:99-112 + densification_postfix); ONE CTA
__global__ void __launch_bounds__(1024)
densify_append_kernel(gi2d_fit_params p, ModelPtrs m, double *__restrict__ stats, const int32_t *__restrict__ sorted_pix,
                      const float *__restrict__ new_cov, int k_rows, int slv) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int n = live_points(p, stats);
    const int k = max(0, min(k_rows, p.num_points - n));
    // pass 1: how many candidates have a positive-definite covariance (the bound is not added: the reference tests
    // new_cov2d itself, gaussianimage_covariance.py:309)
    int mine = 0;
    for (int i = tid; i < k; i += 1024)
        mine += is_pos_def(new_cov[3 * i], new_cov[3 * i + 1], new_cov[3 * i + 2]) ? 1 : 0;
    int tot;
    block_scan_inclusive<1024>(mine, s_warp, &tot);
    if (tid == 0) s_total = tot;
    __syncthreads();
    const int n_new = n + s_total;
    // low_pass = min(H*W / (9*pi*cur_num_points), 300) evaluated in double like the Python expression, then float32
    const float lp = slv ? (float)fmin((double)p.img_height * (double)p.img_width / (9.0 * 3.141592653589793 * (double)n_new), 300.0)
                         : 0.5f;
    // pass 2: append in candidate order (chunks of 1024, running base)
    int base = n;
    for (int c0 = 0; c0 < k; c0 += 1024) {
        const int i = c0 + tid;
        float cx = 0.f, cxy = 0.f, cy = 0.f;
        bool ok = false;
        if (i < k) {
            cx = new_cov[3 * i]; cxy = new_cov[3 * i + 1]; cy = new_cov[3 * i + 2];
            ok = is_pos_def(cx, cxy, cy);
        }
        int chunk;
        const int incl = block_scan_inclusive<1024>(ok ? 1 : 0, s_warp, &chunk);
        if (ok) {
            const int g = base + incl - 1;
            const int pix = sorted_pix[i];
            float r[kRowFloats];
#pragma unroll
            for (int q = 0; q < kRowFloats; ++q) r[q] = 0.f;
            r[0] = (float)(pix % p.img_width);
            r[1] = (float)(pix / p.img_width);
            r[2] = cx; r[3] = cxy; r[4] = cy;
            r[8] = lp; r[9] = 0.f; r[10] = lp;
            store_row(m, g, r);
        }
        base += chunk;
    }
    if (tid == 0) {
        stats[GI2D_STAT_ADDED] = (double)s_total;
        if (p.dynamic_points) stats[GI2D_STAT_NUM_POINTS] = (double)n_new;
    }
}

__global__ void set_stat_kernel(double *stats, int slot, double v) { stats[slot] = v; }

// ------------------------------------------------------------------------------- export of the binning
// The bucketed layout keeps a tile's (sorted) keys at tile * C; the reference's arrays -- isect_ids_sorted /
// gaussian_ids_sorted as ONE ascending key array and tile_bins (forward.cu:211-233) -- are produced on demand:
// one CTA scans the counts of the last forward into tile ranges, then one CTA per tile copies its keys.
__global__ void __launch_bounds__(1024)
export_ranges_kernel(int num_tiles, int bucket_cap, const int32_t *__restrict__ count, int32_t *__restrict__ tile_bins) {
    __shared__ int s_warp[32];
    int carry = 0;
    for (int base = 0; base < num_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int c = t < num_tiles ? min(__ldcg(count + t), bucket_cap) : 0;
        int tot;
        const int incl = block_scan_inclusive<1024>(c, s_warp, &tot);
        if (t < num_tiles)
            reinterpret_cast<int2 *>(tile_bins)[t] = c ? make_int2(carry + incl - c, carry + incl) : make_int2(0, 0);
        carry += tot;
    }
}

__global__ void __launch_bounds__(64)
export_keys_kernel(int bucket_cap, const int32_t *__restrict__ tile_bins, const uint64_t *__restrict__ bucket_keys,
                   uint64_t *__restrict__ keys_out) {
    const int tile = blockIdx.x;
    const int2 range = __ldcg(reinterpret_cast<const int2 *>(tile_bins) + tile);
    const int n = range.y - range.x;
    const uint64_t *src = bucket_keys + (size_t)tile * bucket_cap;
    if (n <= 960) {   // (sorted by the rasterizer)
        for (int e = threadIdx.x; e < n; e += 64) keys_out[range.x + e] = __ldcg(src + e);
        return;
    }
    // a tile the rasterizer only SELECTED its 256 smallest ids from: rank here (a diagnostic path)
    for (int e = threadIdx.x; e < n; e += 64) {
        const uint64_t key = __ldcg(src + e);
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (__ldcg(src + j) < key) ? 1 : 0;
        keys_out[range.x + rank] = key;
    }
}

int validate(const gi2d_fit_params *p, const gi2d_fit_buffers *b) {
    GI2D_REQUIRE(p && b, "null params");
    GI2D_REQUIRE(p->num_points >= 0 && p->img_width > 0 && p->img_height > 0, "bad sizes");
    GI2D_REQUIRE(p->tiles_x == cdiv(p->img_width, kTile) && p->tiles_y == cdiv(p->img_height, kTile),
                 "tile grid must be ceil(size/16)");
    GI2D_REQUIRE(p->tiles_x <= 65535 && p->tiles_y <= 65535, "image too large");
    GI2D_REQUIRE(0 <= p->tile_row_begin && p->tile_row_begin <= p->tile_row_end && p->tile_row_end <= p->tiles_y,
                 "bad tile row band");
    GI2D_REQUIRE(p->isect_capacity > 0, "isect_capacity must be positive");
    GI2D_REQUIRE(p->loss_ssim_weight == 0.f || (p->img_width >= 11 && p->img_height >= 11),
                 "the SSIM window needs an image of at least 11x11 pixels");
    GI2D_REQUIRE(p->loss_msssim_weight == 0.f || p->loss_ssim_weight == 0.f, "one of the SSIM / MS-SSIM terms at a time");
    GI2D_REQUIRE(p->loss_msssim_weight == 0.f || p->loss_msssim_win == 11 || p->loss_msssim_win == 5,
                 "loss_msssim_win must be 11 or 5");
    GI2D_REQUIRE(p->loss_msssim_weight == 0.f ||
                     ((p->img_width < p->img_height ? p->img_width : p->img_height) > (p->loss_msssim_win - 1) * 16 &&
                      p->loss_scale == 0.f),
                 "MS-SSIM needs the smaller image side to exceed (win - 1) * 16 pixels, and goes with an l1 term only");
    GI2D_REQUIRE(p->loss_msssim_weight == 0.f || (p->tile_row_begin == 0 && p->tile_row_end == p->tiles_y),
                 "SSIM losses are not available for a tile-row band (the window crosses band borders)");
    GI2D_REQUIRE(p->loss_ssim_weight == 0.f || (p->tile_row_begin == 0 && p->tile_row_end == p->tiles_y),
                 "SSIM losses are not available for a tile-row band (the window crosses band borders)");
    GI2D_REQUIRE(b->stats && b->workspace && b->proj && b->sorted_keys && b->tile_bins, "null buffer");
    return GI2D_OK;
}

struct Marks {            // optional per-kernel timing marks (gi2d_fit_profile)
    cudaEvent_t ev[8];
    int n = 0;
    bool on = false;
    void mark(cudaStream_t st) { if (on && n < 8) cudaEventRecord(ev[n++], st); }
};

int fit_forward_backward_impl(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int with_backward,
                              cudaStream_t st, Marks *mk, const gi2d_tilerow *tr = nullptr) {
    const int rc = validate(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(p->num_points == 0 || (b->xyz && b->cov && b->cov_bound && b->rgb), "null parameter buffer");
    GI2D_REQUIRE(!with_backward || (b->m_xyz && b->v_xyz && b->m_cov && b->v_cov && b->m_rgb && b->v_rgb),
                 "a training step needs the Adam moment buffers");
    GI2D_REQUIRE(!with_backward || (b->grads && (b->gt_hwc || b->gt_u8_hwc)), "fit needs grads and a target image");
    const Plan pl = make_plan(*p);
    const Workspace w = carve(*p, pl, b->workspace);
    if (b->workspace_bytes < w.total) {
        set_error("gi2d_fit_forward_backward: workspace too small (%zu < %zu)", b->workspace_bytes, w.total);
        return GI2D_ERR_WORKSPACE;
    }
    const int num_tiles = pl.num_tiles;
    if (mk) mk->mark(st);
    const AdamPtrs ap{b->xyz, b->cov, b->rgb, b->m_xyz, b->v_xyz, b->m_cov, b->v_cov, b->m_rgb, b->v_rgb};
    const int proj_threads = p->num_points <= (1 << 16) ? 64 : kProjThreads;
    if (tr)   // tile-row split: the records and boxes came from their owners; count this band's overlaps
        tr_count_kernel<<<max(1, cdiv(p->num_points, 256)), 256, 0, st>>>(
            *p, *tr, (const ushort4 *)tr->peer_boxes[tr->rank], w.boxes, w.tile_count, (float4 *)b->grads,
            with_backward);
    else if (pl.bucket_cap)
        launch_pdl(fit_project_kernel<true>, dim3(max(1, cdiv(p->num_points, proj_threads)) + (pl.ordered ? 1 : 0)),
            dim3(proj_threads), 0, st,
            *p, ap, b->cov_bound, (float4 *)b->proj, (float4 *)b->grads, w.boxes, w.tile_count, b->stats, with_backward,
            (float4 *)b->best, (with_backward && !p->external_optimizer) ? 1 : 0, b->best_bound, w.tile_fill,
            b->sorted_keys, w.records, pl.bucket_cap, pl.ordered ? (int32_t *)w.tile_work : (int32_t *)nullptr);
    else
        launch_pdl(fit_project_kernel<false>, dim3(max(1, cdiv(p->num_points, proj_threads))), dim3(proj_threads), 0, st,
            *p, ap, b->cov_bound, (float4 *)b->proj, (float4 *)b->grads, w.boxes, w.tile_count, b->stats, with_backward,
            (float4 *)b->best, (with_backward && !p->external_optimizer) ? 1 : 0, b->best_bound, (int32_t *)nullptr,
            (uint64_t *)nullptr, (float4 *)nullptr, 0, (int32_t *)nullptr);
    if (mk) mk->mark(st);
    if (!pl.bucket_cap && !pl.smem_scan) {
        // more tiles than one CTA scans in shared memory: device-wide inclusive prefix sum of the counts
        const int r2 = cumsum_i32_launch(num_tiles, w.tile_count, w.tile_incl, nullptr, w.scan_ws, st);
        if (r2 != GI2D_OK) return r2;
    }
    if (mk) mk->mark(st);
    if (pl.bucket_cap) {
        // (no scan, no placement kernel)
    } else if (pl.smem_scan)
        launch_pdl(fit_place_kernel<true>, dim3(pl.nblocks), dim3(kPlaceThreads), 0, st,
            *p, with_backward, pl.gpb, num_tiles, w.boxes, w.tile_count, w.tile_incl, w.tile_fill, b->sorted_keys,
            (const float4 *)b->proj, w.records, b->tile_bins, w.n_isect, b->stats, pl.ordered ? w.tile_work : nullptr);
    else
        launch_pdl(fit_place_kernel<false>, dim3(pl.nblocks), dim3(kPlaceThreads), 0, st,
            *p, with_backward, pl.gpb, num_tiles, w.boxes, w.tile_count, w.tile_incl, w.tile_fill, b->sorted_keys,
            (const float4 *)b->proj, w.records, b->tile_bins, w.n_isect, b->stats, pl.ordered ? w.tile_work : nullptr);
    if (mk) mk->mark(st);
    const int band = p->tile_row_end - p->tile_row_begin;
    if (band > 0) {
        dim3 grid(p->tiles_x, band);
        if (with_backward && (p->loss_ssim_weight != 0.f || p->loss_msssim_weight != 0.f)) {
            // SSIM couples pixels across tile borders: forward everywhere, then the loss gradient image, then
            // the backward half.  (Band-split multi-GPU runs would need a halo exchange of the render.)
            launch_raster<RasterMode::FitForward>(false, grid, st, *p, b->sorted_keys, w.keys_tmp, b->tile_bins,
                w.tile_count, w.tile_fill, (const float4 *)w.records, b->gt_hwc, b->gt_u8_hwc, w.loss_render, nullptr,
                b->stats, b->err_map, nullptr, pl.ordered ? w.tile_work : nullptr, pl.bucket_cap);
            if (b->out_img)
                cudaMemcpyAsync(b->out_img, w.loss_render, (size_t)p->img_width * p->img_height * 12,
                                cudaMemcpyDeviceToDevice, st);
            if (p->loss_msssim_weight != 0.f)
                msssim_grad_launch(p->img_height, p->img_width, p->loss_msssim_win, w.loss_render, b->gt_hwc,
                                   b->gt_u8_hwc, w.loss_dm, p->loss_msssim_weight, p->loss_l1_scale, w.loss_vout,
                                   b->stats + GI2D_STAT_MSSSIM, st);
            else
                ssim_grad_launch(p->img_height, p->img_width, w.loss_render, b->gt_hwc, b->gt_u8_hwc, w.loss_dm,
                                 p->loss_ssim_weight, p->loss_scale, p->loss_l1_scale, w.loss_vout,
                                 b->stats + GI2D_STAT_SSIM_SUM, st);
            launch_raster<RasterMode::FitBackward>(false, grid, st, *p, b->sorted_keys, w.keys_tmp, b->tile_bins,
                w.tile_count, w.tile_fill, (const float4 *)w.records, nullptr, nullptr, nullptr, b->grads, b->stats,
                nullptr, w.loss_vout, pl.ordered ? w.tile_work : nullptr, pl.bucket_cap);
        } else if (with_backward)
            launch_raster<RasterMode::Fit>(true, grid, st, *p, b->sorted_keys, w.keys_tmp, b->tile_bins, w.tile_count,
                w.tile_fill, (const float4 *)w.records, b->gt_hwc, b->gt_u8_hwc, b->out_img, b->grads, b->stats,
                b->err_map, nullptr, pl.ordered ? w.tile_work : nullptr, pl.bucket_cap);
        else
            launch_raster<RasterMode::Render>(true, grid, st, *p, b->sorted_keys, w.keys_tmp, b->tile_bins,
                w.tile_count, w.tile_fill, (const float4 *)w.records, nullptr, nullptr, b->out_img, nullptr, b->stats,
                nullptr, nullptr, pl.ordered ? w.tile_work : nullptr, pl.bucket_cap);
    }
    if (mk) mk->mark(st);
    return check_launch("gi2d_fit_forward_backward");
}

// FP32 peak microbenchmark: 16 independent chains of packed FMAs (fma.rn.f32x2 -> SASS FFMA2), 512 per loop trip,
// so that loop overhead is < 1 % of the issue slots.  tools/ubench/fp32_issue.cu measured on the B200: FFMA2 74.1
// TFLOP/s = 0.995 of nominal (148 SMs x 128 lanes x 2 x 1.965 GHz = 74.45); scalar FFMA 71.3 (0.958, register-bank
// limited).  (The round-1 kernel -- 8 scalar chains, 8 FFMAs per trip -- read 64.1 and inflated every fraction.)
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float m, float c) {
    constexpr int kChains = 16;
    f32x2 a[kChains];
    const f32x2 mm = pk2(m, m), cc = pk2(c, c);
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = pk2((float)(threadIdx.x + i), (float)(threadIdx.x - i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 32; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fma2(a[i], mm, cc);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        float x, y;
        unpk2(a[i], x, y);
        s += x + y;
    }
    if (s == 12345.678f) out[0] = s;
}

}  // namespace
}  // namespace gi2d

using namespace gi2d;

extern "C" size_t gi2d_fit_workspace_size(const gi2d_fit_params *p) {
    if (!p) return 0;
    const Plan pl = make_plan(*p);
    return carve(*p, pl, nullptr).total;
}

extern "C" int gi2d_fit_launch_count(const gi2d_fit_params *p, int with_backward) {
    if (!p) return 0;
    const Plan pl = make_plan(*p);
    int n = pl.bucket_cap ? 2 : 3;  // project(+Adam)[+place], [place,] raster
    if (!pl.bucket_cap && !pl.smem_scan) n += cdiv(pl.num_tiles, 2048) > 1 ? 3 : 1;  // device-wide scan of the tile counts
    if (with_backward && p->loss_ssim_weight != 0.f) n += 3;  // forward / SSIM stats / SSIM gradient / backward
    if (with_backward && p->loss_msssim_weight != 0.f) n += 21; // ... / memset + 19 MS-SSIM launches / ...
    return n;
}

extern "C" int gi2d_fit_bucket_capacity(const gi2d_fit_params *p) {
    if (!p) return 0;
    return make_plan(*p).bucket_cap;
}

extern "C" int gi2d_fit_export_binning(const gi2d_fit_params *p, const gi2d_fit_buffers *b, uint64_t *sorted_keys_out,
                                       int32_t *tile_bins_out, gi2d_stream_t stream) {
    const int rc = validate(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(sorted_keys_out && tile_bins_out, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    const Plan pl = make_plan(*p);
    const Workspace w = carve(*p, pl, b->workspace);
    if (!pl.bucket_cap) {   // the compact arrays are what the step produces
        cudaMemcpyAsync(sorted_keys_out, b->sorted_keys, (size_t)p->isect_capacity * 8, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(tile_bins_out, b->tile_bins, (size_t)pl.num_tiles * 8, cudaMemcpyDeviceToDevice, st);
        return check_launch(__func__);
    }
    double bank = 0.0;   // (host read of one flag: this is a diagnostic / test entry point, it synchronises)
    cudaStreamSynchronize(st);
    cudaMemcpy(&bank, b->stats + kStatBank, sizeof(double), cudaMemcpyDeviceToHost);
    const int32_t *count = bank != 0.0 ? w.tile_fill : w.tile_count;
    export_ranges_kernel<<<1, 1024, 0, st>>>(pl.num_tiles, pl.bucket_cap, count, tile_bins_out);
    export_keys_kernel<<<pl.num_tiles, 64, 0, st>>>(pl.bucket_cap, tile_bins_out, b->sorted_keys, sorted_keys_out);
    return check_launch(__func__);
}

extern "C" int gi2d_fit_reset(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int step,
                              gi2d_stream_t stream) {
    GI2D_REQUIRE(p && b && b->stats, "null stats");
    if (b->workspace) {  // the per-tile counters must start from zero (K3 hands them back zeroed after every step)
        const Plan pl = make_plan(*p);
        const Workspace w = carve(*p, pl, b->workspace);
        if (b->workspace_bytes >= w.total)
            cudaMemsetAsync(w.tile_count, 0, (char *)w.tile_fill - (char *)w.tile_count + (size_t)pl.num_tiles * 4,
                            (cudaStream_t)stream);
    }
    fit_reset_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(*p, b->stats, step);
    return check_launch(__func__);
}

extern "C" int gi2d_fit_forward_backward(const gi2d_fit_params *p, const gi2d_fit_buffers *b,
                                         int with_backward, gi2d_stream_t stream) {
    return fit_forward_backward_impl(p, b, with_backward, (cudaStream_t)stream, nullptr);
}

extern "C" int gi2d_fit_adam(const gi2d_fit_params *p, const gi2d_fit_buffers *b, gi2d_stream_t stream) {
    const int rc = validate(p, b);
    if (rc != GI2D_OK) return rc;
    if (p->num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(b->xyz && b->cov && b->cov_bound && b->rgb && b->m_xyz && b->v_xyz && b->m_cov && b->v_cov &&
                     b->m_rgb && b->v_rgb && b->grads,
                 "null buffer");
    const AdamPtrs ap{b->xyz, b->cov, b->rgb, b->m_xyz, b->v_xyz, b->m_cov, b->v_cov, b->m_rgb, b->v_rgb};
    launch_pdl(fit_adam_kernel, dim3(cdiv(p->num_points, 256)), dim3(256), 0, (cudaStream_t)stream, *p, ap,
               (const float *)b->cov_bound, (const float4 *)b->proj, (const float4 *)b->grads, b->stats,
               (float4 *)b->best, b->best_bound);
    fit_clear_pending_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b->stats);
    return check_launch(__func__);
}

extern "C" int gi2d_fit_set_num_points(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int n,
                                       gi2d_stream_t stream) {
    GI2D_REQUIRE(p && b && b->stats, "null stats");
    GI2D_REQUIRE(n >= 0 && n <= p->num_points, "count outside [0, capacity]");
    set_stat_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(b->stats, GI2D_STAT_NUM_POINTS, (double)n);
    return check_launch(__func__);
}

static ModelPtrs model_ptrs(const gi2d_fit_buffers *b) {
    return ModelPtrs{b->xyz, b->cov, b->rgb, b->cov_bound, b->m_xyz, b->v_xyz, b->m_cov, b->v_cov, b->m_rgb, b->v_rgb};
}

extern "C" size_t gi2d_fit_prune_workspace_size(int capacity) {
    const size_t n = (size_t)(capacity > 0 ? capacity : 1);
    return 2 * align_up(n * 4) + align_up(cumsum_i32_workspace((int)n)) + align_up(n * kRowFloats * 4);
}

extern "C" int gi2d_fit_prune(const gi2d_fit_params *p, const gi2d_fit_buffers *b, void *workspace,
                              size_t workspace_bytes, gi2d_stream_t stream) {
    int rc = gi2d_fit_adam(p, b, stream);   // (validates; applies a pending step; counts the non-PSD rows)
    if (rc != GI2D_OK) return rc;
    if (p->num_points == 0) return GI2D_OK;
    if (!workspace || workspace_bytes < gi2d_fit_prune_workspace_size(p->num_points)) {
        set_error("%s: workspace too small", __func__);
        return GI2D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)p->num_points;
    char *c = (char *)workspace;
    int32_t *keep = (int32_t *)c;   c += align_up(n * 4);
    int32_t *incl = (int32_t *)c;   c += align_up(n * 4);
    int32_t *scan_ws = (int32_t *)c; c += align_up(cumsum_i32_workspace((int)n));
    float *scratch = (float *)c;
    const ModelPtrs m = model_ptrs(b);
    const int grid = cdiv(p->num_points, 256);
    prune_flag_kernel<<<grid, 256, 0, st>>>(*p, m, b->stats, keep);
    rc = cumsum_i32_launch(p->num_points, keep, incl, nullptr, scan_ws, st);
    if (rc != GI2D_OK) return rc;
    prune_gather_kernel<<<grid, 256, 0, st>>>(*p, m, b->stats, keep, incl, scratch);
    prune_scatter_kernel<<<grid, 256, 0, st>>>(*p, m, b->stats, incl, scratch);
    return check_launch(__func__);
}

extern "C" size_t gi2d_fit_densify_workspace_size(int img_height, int img_width) {
    const size_t n = (size_t)img_height * img_width;
    return 2 * align_up(n * 8) + 2 * align_up(n * 4) + align_up(gi2d_sort_workspace_size((int)n));
}

extern "C" int gi2d_fit_densify(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int k_rows,
                                const float *new_cov2d, int slv, void *workspace, size_t workspace_bytes,
                                gi2d_stream_t stream) {
    int rc = gi2d_fit_adam(p, b, stream);   // (validates; nothing may be pending while rows are appended)
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(p->dynamic_points, "densification needs dynamic_points (capacity-sized arrays)");
    GI2D_REQUIRE(b->err_map, "no error map bound (a training step writes it when b->err_map is set)");
    GI2D_REQUIRE(k_rows >= 0 && (k_rows == 0 || new_cov2d), "bad candidate block");
    const size_t n = (size_t)p->img_height * p->img_width;
    GI2D_REQUIRE(n < ((size_t)1 << 31), "image too large for 32-bit pixel indices");
    if (k_rows == 0) return GI2D_OK;
    if (!workspace || workspace_bytes < gi2d_fit_densify_workspace_size(p->img_height, p->img_width)) {
        set_error("%s: workspace too small", __func__);
        return GI2D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *c = (char *)workspace;
    int64_t *keys = (int64_t *)c;     c += align_up(n * 8);
    int64_t *keys_s = (int64_t *)c;   c += align_up(n * 8);
    int32_t *vals = (int32_t *)c;     c += align_up(n * 4);
    int32_t *vals_s = (int32_t *)c;   c += align_up(n * 4);
    void *sort_ws = c;
    densify_keys_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>((int)n, b->err_map, keys, vals);
    rc = gi2d_sort_pairs_i64((int)n, keys, vals, keys_s, vals_s, 0, 63, sort_ws, gi2d_sort_workspace_size((int)n), stream);
    if (rc != GI2D_OK) return rc;
    densify_append_kernel<<<1, 1024, 0, st>>>(*p, model_ptrs(b), b->stats, vals_s, new_cov2d, k_rows, slv);
    return check_launch(__func__);
}

extern "C" size_t gi2d_bin_sort_workspace_size(int num_points, int tiles_x, int tiles_y, int capacity) {
    const size_t T = (size_t)(tiles_x > 0 && tiles_y > 0 ? (size_t)tiles_x * tiles_y : 1);
    const size_t n = (size_t)(num_points > 0 ? num_points : 1);
    return 3 * align_up(T * 4) + align_up(cumsum_i32_workspace((int)T)) + align_up(n * 8) +
           align_up((size_t)(capacity > 0 ? capacity : 1) * 8);
}

extern "C" int gi2d_bin_sort(int num_points, const float *xys, const float *depths, const int32_t *radii,
                             int tiles_x, int tiles_y, float radius_clip, int capacity,
                             int64_t *isect_ids_sorted, int32_t *gaussian_ids_sorted, int32_t *tile_bins,
                             int32_t *info, void *workspace, size_t workspace_bytes, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0 && tiles_x > 0 && tiles_y > 0 && capacity > 0, "bad sizes");
    GI2D_REQUIRE(tiles_x <= 65535 && tiles_y <= 65535, "image too large");
    GI2D_REQUIRE(isect_ids_sorted && gaussian_ids_sorted && tile_bins && info, "null output");
    GI2D_REQUIRE(num_points == 0 || (xys && radii), "null input");
    if (!workspace || workspace_bytes < gi2d_bin_sort_workspace_size(num_points, tiles_x, tiles_y, capacity)) {
        set_error("%s: workspace too small", __func__);
        return GI2D_ERR_WORKSPACE;
    }
    (void)depths;   // every depth shares one bit pattern by contract (0.0 from the 2-D projections): low key word 0
    cudaStream_t st = (cudaStream_t)stream;
    const int T = tiles_x * tiles_y;
    char *c = (char *)workspace;
    int32_t *tile_count = (int32_t *)c;  c += align_up((size_t)T * 4);
    int32_t *tile_fill = (int32_t *)c;   c += align_up((size_t)T * 4);
    int32_t *tile_incl = (int32_t *)c;   c += align_up((size_t)T * 4);
    int32_t *scan_ws = (int32_t *)c;     c += align_up(cumsum_i32_workspace(T));
    ushort4 *boxes = (ushort4 *)c;       c += align_up((size_t)(num_points > 0 ? num_points : 1) * 8);
    uint64_t *keys = (uint64_t *)c;
    cudaMemsetAsync(tile_count, 0, (char *)tile_incl - (char *)tile_count, st);   // counts + cursors
    gi2d_fit_params p{};
    p.num_points = num_points;
    p.tiles_x = tiles_x;
    p.tiles_y = tiles_y;
    p.tile_row_end = tiles_y;
    p.isect_capacity = capacity;
    if (num_points > 0)
        bs_count_kernel<<<cdiv(num_points, 256), 256, 0, st>>>(num_points, xys, radii, tiles_x, tiles_y, radius_clip,
                                                               boxes, tile_count);
    const Plan pl = make_plan(p);
    if (pl.smem_scan) {
        fit_place_kernel<true, false><<<pl.nblocks, kPlaceThreads, 0, st>>>(
            p, 0, pl.gpb, T, boxes, tile_count, tile_incl, tile_fill, keys, nullptr, nullptr, tile_bins, info, nullptr, nullptr);
    } else {
        const int rc = cumsum_i32_launch(T, tile_count, tile_incl, nullptr, scan_ws, st);
        if (rc != GI2D_OK) return rc;
        fit_place_kernel<false, false><<<pl.nblocks, kPlaceThreads, 0, st>>>(
            p, 0, pl.gpb, T, boxes, tile_count, tile_incl, tile_fill, keys, nullptr, nullptr, tile_bins, info, nullptr, nullptr);
    }
    bs_tile_sort_kernel<<<T, 64, 0, st>>>(capacity, tile_bins, keys, (int64_t)0, isect_ids_sorted, gaussian_ids_sorted);
    return check_launch(__func__);
}

// Measurement utility (bench.py): one full fit step with a CUDA event between the kernels, on
// `stream`; SYNCHRONISES.  ms[0..4] = adam(prev)+project, tile-count scan (0 up to 2048 tiles), place, raster, 0.
extern "C" int gi2d_fit_profile(const gi2d_fit_params *p, const gi2d_fit_buffers *b, float *ms_host,
                                gi2d_stream_t stream) {
    GI2D_REQUIRE(ms_host, "null ms_host");
    cudaStream_t st = (cudaStream_t)stream;
    Marks mk;
    mk.on = true;
    for (int i = 0; i < 8; ++i) cudaEventCreate(&mk.ev[i]);
    int rc = fit_forward_backward_impl(p, b, 1, st, &mk);
    mk.mark(st);
    cudaStreamSynchronize(st);
    for (int i = 0; i < 5; ++i) {
        ms_host[i] = 0.f;
        if (rc == GI2D_OK && i + 1 < mk.n) cudaEventElapsedTime(&ms_host[i], mk.ev[i], mk.ev[i + 1]);
    }
    for (int i = 0; i < 8; ++i) cudaEventDestroy(mk.ev[i]);
    return rc;
}

extern "C" int gi2d_fit_profile_raster(const gi2d_fit_params *p, const gi2d_fit_buffers *b, int reps, float *ms_host,
                                       gi2d_stream_t stream) {
    GI2D_REQUIRE(ms_host && reps > 0, "bad arguments");
    GI2D_REQUIRE(p && p->loss_ssim_weight == 0.f && p->loss_msssim_weight == 0.f, "profiles the single-launch rasterizer (no SSIM term)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = fit_forward_backward_impl(p, b, 1, st, nullptr);
    if (rc != GI2D_OK) return rc;
    const Plan pl = make_plan(*p);
    const Workspace w = carve(*p, pl, b->workspace);
    const int band = p->tile_row_end - p->tile_row_begin;
    GI2D_REQUIRE(band > 0, "empty band");
    const dim3 grid(p->tiles_x, band);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; ++i)
        launch_raster<RasterMode::Fit>(true, grid, st, *p, b->sorted_keys, w.keys_tmp, b->tile_bins, w.tile_count,
            w.tile_fill, (const float4 *)w.records, b->gt_hwc, b->gt_u8_hwc, nullptr, b->grads, b->stats, nullptr,
            nullptr, pl.ordered ? w.tile_work : nullptr, pl.bucket_cap);
    cudaEventRecord(e1, st);
    // nothing pending (the accumulated gradient is garbage), loss accumulators back to zero
    cudaMemsetAsync(b->stats + kStatPending, 0, sizeof(double), st);
    cudaMemsetAsync(b->stats + GI2D_STAT_SSE, 0, GI2D_STAT_SSE_SLOTS * sizeof(double), st);
    cudaMemsetAsync(b->stats + GI2D_STAT_ABS_SUM, 0, sizeof(double), st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_host = ms / (float)reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamSynchronize(st);
    return check_launch(__func__);
}

// Measurement utility: FP32 FMA issue peak of this GPU (the roofline denominator of the raster
// kernels; MEASURED_PEAKS.json only carries HBM and bf16 tensor figures).  Returns TFLOP/s.
extern "C" int gi2d_measure_fp32_peak(float *tflops_host, gi2d_stream_t stream) {
    GI2D_REQUIRE(tflops_host, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    float *d = nullptr;
    cudaMalloc(&d, 4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 1 << 10, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, st);
        fma_peak_kernel<<<grid, 256, 0, st>>>(d, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // 512 packed FMAs per trip and thread = 2048 FLOP
        const float tf = (float)((double)grid * 256.0 * iters * 512.0 * 4.0 / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_host = best;
    return check_launch(__func__);
}

extern "C" int gi2d_fit_input_grads(const gi2d_fit_params *p, const gi2d_fit_buffers *b, float *out,
                                    gi2d_stream_t stream) {
    const int rc = validate(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(b->grads && out, "null buffer");
    if (p->num_points == 0) return GI2D_OK;
    launch_pdl(fit_input_grads_kernel, dim3(cdiv(p->num_points, 256)), dim3(256), 0, (cudaStream_t)stream,
               p->num_points, (const float4 *)b->proj, (const float4 *)b->grads, (float4 *)out,
               (const double *)b->stats, p->dynamic_points);
    return check_launch(__func__);
}

// ---- tile-row split entry points (kernels: tr_count_kernel / tr_exchange_kernel above)
static int validate_tilerow(const gi2d_fit_params *p, const gi2d_fit_buffers *b, const gi2d_tilerow *tr) {
    const int rc = validate(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(tr, "null tile-row descriptor");
    GI2D_REQUIRE(tr->world >= 1 && tr->world <= GI2D_MAX_RANKS && tr->rank >= 0 && tr->rank < tr->world,
                 "bad rank/world (<= 8 GPUs of one box)");
    GI2D_REQUIRE(p->external_optimizer == 2, "set external_optimizer = 2 (the exchange kernel owns the optimiser)");
    GI2D_REQUIRE(tr->band_edge[0] == 0 && tr->band_edge[tr->world] == p->tiles_y, "bands must cover the image");
    for (int q = 0; q < tr->world; ++q) {
        GI2D_REQUIRE(tr->band_edge[q] <= tr->band_edge[q + 1], "band edges must ascend");
        GI2D_REQUIRE(tr->peer_grads[q] && tr->peer_proj[q] && tr->peer_boxes[q] && tr->peer_flags[q], "null peer pointer");
    }
    GI2D_REQUIRE(p->tile_row_begin == tr->band_edge[tr->rank] && p->tile_row_end == tr->band_edge[tr->rank + 1],
                 "tile_row_begin/end must be this rank's band");
    GI2D_REQUIRE(0 <= tr->own_begin && tr->own_begin <= tr->own_end && tr->own_end <= p->num_points, "bad owned slice");
    GI2D_REQUIRE(tr->ctrl, "null control block");
    GI2D_REQUIRE(b->grads == tr->peer_grads[tr->rank] && b->proj == tr->peer_proj[tr->rank],
                 "b->grads / b->proj must be this rank's peer-visible buffers");
    GI2D_REQUIRE(b->xyz && b->cov && b->cov_bound && b->rgb && b->m_xyz && b->v_xyz && b->m_cov && b->v_cov &&
                     b->m_rgb && b->v_rgb, "null parameter / moment buffer");
    return GI2D_OK;
}

extern "C" int gi2d_tilerow_init(const gi2d_fit_params *p, const gi2d_fit_buffers *b, const gi2d_tilerow *tr,
                                 gi2d_stream_t stream) {
    const int rc = validate_tilerow(p, b, tr);
    if (rc != GI2D_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // control block: epoch 1, tickets 0.  (The peer-visible buffers were zeroed by the caller BEFORE its barrier:
    // a peer's scatter below may reach our records / boxes before this call even starts.)
    static const uint32_t ctrl0[8] = {1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    cudaMemcpyAsync(tr->ctrl, ctrl0, sizeof(ctrl0), cudaMemcpyHostToDevice, st);
    const int n = tr->own_end - tr->own_begin;
    if (n > 0) {
        const AdamPtrs ap{b->xyz, b->cov, b->rgb, b->m_xyz, b->v_xyz, b->m_cov, b->v_cov, b->m_rgb, b->v_rgb};
        tr_exchange_kernel<<<cdiv(n, 256), 256, 0, st>>>(*p, *tr, ap, b->cov_bound, b->stats, 1);
    }
    return check_launch(__func__);
}

extern "C" int gi2d_tilerow_step(const gi2d_fit_params *p, const gi2d_fit_buffers *b, const gi2d_tilerow *tr,
                                 int with_backward, int phase, gi2d_stream_t stream) {
    int rc = validate_tilerow(p, b, tr);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(phase >= 1 && phase <= 3, "phase: 1 = band step, 2 = exchange, 3 = both");
    GI2D_REQUIRE(p->loss_ssim_weight == 0.f && p->loss_msssim_weight == 0.f, "SSIM losses are not available for a tile-row band");
    cudaStream_t st = (cudaStream_t)stream;
    if (phase & 1) {
        rc = fit_forward_backward_impl(p, b, with_backward, st, nullptr, tr);
        if (rc != GI2D_OK) return rc;
        if (!tr->sync && with_backward) tr_publish_kernel<<<1, 1, 0, st>>>(*tr, b->stats);
    }
    if ((phase & 2) && with_backward) {
        const int n = tr->own_end - tr->own_begin;
        const AdamPtrs ap{b->xyz, b->cov, b->rgb, b->m_xyz, b->v_xyz, b->m_cov, b->v_cov, b->m_rgb, b->v_rgb};
        // (every rank launches at least one CTA: it takes part in the flag exchange even with an empty slice)
        tr_exchange_kernel<<<max(1, cdiv(n, 256)), 256, 0, st>>>(*p, *tr, ap, b->cov_bound, b->stats, 0);
    }
    return check_launch(__func__);
}

// ------------------------------------------------------------------------------ host-fed steps
struct gi2d_host_pipe {
    cudaStream_t copy;
    cudaEvent_t copied[2], stats[2];
    cudaEvent_t read[2];        // recorded after the last step that read device target buffer buf[i]
    const void *buf[2];
    bool read_valid[2];
    unsigned long long calls;
};

extern "C" int gi2d_host_pipe_create(gi2d_host_pipe **out) {
    GI2D_REQUIRE(out, "null output");
    gi2d_host_pipe *pp = new gi2d_host_pipe();
    cudaStreamCreateWithFlags(&pp->copy, cudaStreamNonBlocking);
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&pp->copied[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&pp->stats[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&pp->read[i], cudaEventDisableTiming);
        pp->buf[i] = nullptr;
        pp->read_valid[i] = false;
    }
    pp->calls = 0;
    *out = pp;
    return check_launch(__func__);
}

extern "C" int gi2d_host_pipe_destroy(gi2d_host_pipe *pp) {
    if (!pp) return GI2D_OK;
    cudaStreamSynchronize(pp->copy);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(pp->copied[i]);
        cudaEventDestroy(pp->stats[i]);
        cudaEventDestroy(pp->read[i]);
    }
    cudaStreamDestroy(pp->copy);
    delete pp;
    return GI2D_OK;
}

extern "C" int gi2d_fit_step_host(gi2d_host_pipe *pp, const gi2d_fit_params *p, const gi2d_fit_buffers *b,
                                  const void *host_target, size_t target_bytes, double *host_stats,
                                  gi2d_stream_t stream, int *slot_out) {
    GI2D_REQUIRE(pp && p && b && host_target && host_stats && slot_out, "null argument");
    void *dst = b->gt_u8_hwc ? (void *)b->gt_u8_hwc : (void *)b->gt_hwc;
    GI2D_REQUIRE(dst, "the buffers carry no target image");
    const size_t want = (size_t)p->img_width * p->img_height * 3 * (b->gt_u8_hwc ? 1 : 4);
    GI2D_REQUIRE(target_bytes == want, "target_bytes does not match the image size / dtype");
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = (int)(pp->calls & 1);
    // which of the (at most two) device target buffers is this?  (a third evicts the older entry)
    int e = pp->buf[0] == dst ? 0 : (pp->buf[1] == dst ? 1 : -1);
    if (e < 0) {
        e = pp->buf[0] == nullptr ? 0 : (pp->buf[1] == nullptr ? 1 : slot);
        pp->buf[e] = dst;
        pp->read_valid[e] = false;
    }
    // upload: after the last step that read this buffer, concurrently with whatever `stream` is running now
    if (pp->read_valid[e]) cudaStreamWaitEvent(pp->copy, pp->read[e], 0);
    cudaMemcpyAsync(dst, host_target, target_bytes, cudaMemcpyHostToDevice, pp->copy);
    cudaEventRecord(pp->copied[slot], pp->copy);
    cudaStreamWaitEvent(st, pp->copied[slot], 0);
    const int rc = fit_forward_backward_impl(p, b, 1, st, nullptr);
    if (rc != GI2D_OK) return rc;
    cudaEventRecord(pp->read[e], st);
    pp->read_valid[e] = true;
    cudaMemcpyAsync(host_stats, b->stats, GI2D_STAT_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaEventRecord(pp->stats[slot], st);
    *slot_out = slot;
    pp->calls++;
    return check_launch(__func__);
}

extern "C" int gi2d_host_pipe_wait(gi2d_host_pipe *pp, int slot) {
    GI2D_REQUIRE(pp && (slot == 0 || slot == 1), "bad slot");
    const cudaError_t err = cudaEventSynchronize(pp->stats[slot]);
    if (err != cudaSuccess) {
        set_error("gi2d_host_pipe_wait: %s", cudaGetErrorString(err));
        return GI2D_ERR_CUDA;
    }
    return GI2D_OK;
}
