// gi2d_raster.cu -- stand-alone rasterize-sum forward / backward:
// the kernels behind `rasterize_gaussians_plus` / `rasterize_gaussians_sum`.
// Bounded by FP32 SIMT issue (22 / 43 algorithmic FLOP per pixel x Gaussian pair) plus one
// MUFU.EX2 per pair; see gi2d_raster_core.cuh for the work decomposition.
#include "gi2d_raster_core.cuh"

namespace gi2d {
namespace {

__device__ __forceinline__ int2 load_range(const int32_t *__restrict__ tile_bins, int tile_id, int rows) {
    // tiles beyond the rows the caller owns are empty (the reference reads out of
    // bounds there when num_intersects < #tiles)
    if (tile_id >= rows) return make_int2(0, 0);
    return __ldg(reinterpret_cast<const int2 *>(tile_bins) + tile_id);
}

__global__ void __launch_bounds__(kTilePixels)
raster_fwd_kernel(int tiles_x, int img_w, int img_h, const int32_t *__restrict__ gids,
                  const int32_t *__restrict__ tile_bins, int rows, const float *__restrict__ xys,
                  const float *__restrict__ conics, const float *__restrict__ colors,
                  const float *__restrict__ opacities, float *__restrict__ out_img,
                  float *__restrict__ final_Ts, int32_t *__restrict__ final_idx,
                  const int32_t *__restrict__ num_intersects_dev, const float *__restrict__ background) {
    __shared__ TileGaussians sg;
    const int tid = threadIdx.x;
    const int tile_id = blockIdx.y * tiles_x + blockIdx.x;
    const int blk = tid >> 5;  // warp <-> 8x4 sub-block
    const int j = blockIdx.x * kTile + blk_x(blk, tid & 31);
    const int i = blockIdx.y * kTile + blk_y(blk, tid & 31);
    const int2 range = load_range(tile_bins, tile_id, rows);
    // forward.cu:673 -- only the first 256-batch is ever processed
    const int cnt = max(0, min(kMaxPerTile, range.y - range.x));
    if (tid < cnt)
        stage_gaussian(sg, tid, __ldg(gids + range.x + tid), (float)(blockIdx.x * kTile),
                       (float)(blockIdx.y * kTile), xys, conics, colors, opacities);
    __syncthreads();
    const bool inside = i < img_h && j < img_w;
    float r = 0.f, g = 0.f, b = 0.f;
    int last = -1;
    forward_sweep(sg, cnt, blk, inside, (float)j, (float)i, r, g, b, last);
    // rasterize_sum_plus.py:110-118: with no intersection at all the operator returns ones * background; the
    // caller that kept num_intersects on the device (gi2d_bin_sort) lets the kernel take that branch
    if (num_intersects_dev && __ldg(num_intersects_dev) < 1) {
        r = background ? __ldg(background) : 1.f;
        g = background ? __ldg(background + 1) : 1.f;
        b = background ? __ldg(background + 2) : 1.f;
    }
    if (inside) {
        const size_t pix = (size_t)i * img_w + j;
        out_img[3 * pix] = r;
        out_img[3 * pix + 1] = g;
        out_img[3 * pix + 2] = b;
        if (final_Ts) final_Ts[pix] = 1.f;                                   // forward.cu:682 (T stays 1)
        if (final_idx) final_idx[pix] = last < 0 ? 0 : range.x + last;       // forward.cu:619,669,684
    }
}

template <int kWarps, bool kOpacity>
__global__ void __launch_bounds__(kWarps * 32, 4)
raster_bwd_kernel(int tiles_x, int img_w, int img_h, const int32_t *__restrict__ gids,
                  const int32_t *__restrict__ tile_bins, int rows, const float *__restrict__ xys,
                  const float *__restrict__ conics, const float *__restrict__ colors,
                  const float *__restrict__ opacities, const float *__restrict__ v_output,
                  float *__restrict__ v_xy, float *__restrict__ v_conic,
                  float *__restrict__ v_colors, float *__restrict__ v_opacity) {
    static_assert(kWarps * 32 == kTilePixels, "one thread per pixel stages v_out");
    __shared__ TileGaussians sg;
    __shared__ TileGrad tg;
    __shared__ int s_ids[kMaxPerTile];
    const int tid = threadIdx.x;
    const int tile_id = blockIdx.y * tiles_x + blockIdx.x;
    const int2 range = load_range(tile_bins, tile_id, rows);
    const int cnt = max(0, min(kMaxPerTile, range.y - range.x));
    if (cnt == 0) return;
    {   // stage dL/d(out) of this tile (0 outside the image)
        const int j = blockIdx.x * kTile + (tid & 15), i = blockIdx.y * kTile + (tid >> 4);
        float vr = 0.f, vg = 0.f, vb = 0.f;
        if (i < img_h && j < img_w) {
            const size_t pix = (size_t)i * img_w + j;
            vr = __ldg(v_output + 3 * pix);
            vg = __ldg(v_output + 3 * pix + 1);
            vb = __ldg(v_output + 3 * pix + 2);
        }
        const int gi = grad_index(tid & 15, tid >> 4);
        tg.v[0][gi] = vr;
        tg.v[1][gi] = vg;
        tg.v[2][gi] = vb;
    }
    if (tid < cnt) {
        const int g = __ldg(gids + range.x + tid);
        s_ids[tid] = g;
        stage_gaussian(sg, tid, g, (float)(blockIdx.x * kTile), (float)(blockIdx.y * kTile), xys, conics, colors,
                       opacities);
    }
    __syncthreads();
    const LanePixels lp = lane_pixels(blockIdx.x, blockIdx.y, img_w, img_h);
    backward_tile<kOpacity, kWarps, false>(sg, s_ids, cnt, lp, tg,
                                    [&](int g, int k) -> float * {
                                        return k < 2 ? v_xy + 2 * g + k
                                                     : (k < 5 ? v_conic + 3 * g + (k - 2) : v_colors + 3 * g + (k - 5));
                                    },
                                    v_opacity);
}

}  // namespace
}  // namespace gi2d

using namespace gi2d;

extern "C" int gi2d_rasterize_sum_fwd_dev(int tiles_x, int tiles_y, int img_width, int img_height,
                                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                          int num_bins_rows, const float *xys, const float *conics,
                                          const float *colors, const float *opacities, float *out_img,
                                          float *final_Ts, int32_t *final_idx, const int32_t *num_intersects_dev,
                                          const float *background, gi2d_stream_t stream);

extern "C" int gi2d_rasterize_sum_fwd(int tiles_x, int tiles_y, int img_width, int img_height,
                                      const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                      int num_bins_rows, const float *xys, const float *conics,
                                      const float *colors, const float *opacities, float *out_img,
                                      float *final_Ts, int32_t *final_idx, gi2d_stream_t stream) {
    return gi2d_rasterize_sum_fwd_dev(tiles_x, tiles_y, img_width, img_height, gaussian_ids_sorted, tile_bins,
                                      num_bins_rows, xys, conics, colors, opacities, out_img, final_Ts, final_idx,
                                      nullptr, nullptr, stream);
}

extern "C" int gi2d_rasterize_sum_fwd_dev(int tiles_x, int tiles_y, int img_width, int img_height,
                                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                          int num_bins_rows, const float *xys, const float *conics,
                                          const float *colors, const float *opacities, float *out_img,
                                          float *final_Ts, int32_t *final_idx, const int32_t *num_intersects_dev,
                                          const float *background, gi2d_stream_t stream) {
    GI2D_REQUIRE(tiles_x >= 0 && tiles_y >= 0 && img_width >= 0 && img_height >= 0, "negative size");
    GI2D_REQUIRE(num_bins_rows >= 0, "negative num_bins_rows");
    if (tiles_x == 0 || tiles_y == 0 || img_width == 0 || img_height == 0) return GI2D_OK;
    GI2D_REQUIRE(tiles_x * kTile >= img_width && tiles_y * kTile >= img_height,
                 "tile grid does not cover the image (tiles are 16x16)");
    GI2D_REQUIRE(out_img, "null out_img");
    GI2D_REQUIRE(num_bins_rows == 0 || (gaussian_ids_sorted && tile_bins && xys && conics && colors),
                 "null pointer");
    dim3 grid(tiles_x, tiles_y);
    raster_fwd_kernel<<<grid, kTilePixels, 0, (cudaStream_t)stream>>>(
        tiles_x, img_width, img_height, gaussian_ids_sorted, tile_bins, num_bins_rows, xys, conics,
        colors, opacities, out_img, final_Ts, final_idx, num_intersects_dev, background);
    return check_launch(__func__);
}

extern "C" int gi2d_rasterize_sum_bwd(int num_points, int tiles_x, int tiles_y, int img_width,
                                      int img_height, const int32_t *gaussian_ids_sorted,
                                      const int32_t *tile_bins, int num_bins_rows, const float *xys,
                                      const float *conics, const float *colors,
                                      const float *opacities, const int32_t *final_idx,
                                      const float *v_output, float *v_xy, float *v_conic,
                                      float *v_colors, float *v_opacity, gi2d_stream_t stream) {
    (void)final_idx;  // implied by the 256-per-tile cap; accepted for signature parity
    GI2D_REQUIRE(num_points >= 0 && tiles_x >= 0 && tiles_y >= 0 && num_bins_rows >= 0, "negative size");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(v_xy && v_conic && v_colors, "null gradient output");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)num_points;
    if (v_conic == v_xy + 2 * n && v_colors == v_conic + 3 * n && (!v_opacity || v_opacity == v_colors + 3 * n)) {
        // one block (the binding allocates the four outputs back to back): one fill instead of four
        cudaMemsetAsync(v_xy, 0, n * (v_opacity ? 9 : 8) * sizeof(float), st);
    } else {
        cudaMemsetAsync(v_xy, 0, n * 2 * sizeof(float), st);
        cudaMemsetAsync(v_conic, 0, n * 3 * sizeof(float), st);
        cudaMemsetAsync(v_colors, 0, n * 3 * sizeof(float), st);
        if (v_opacity) cudaMemsetAsync(v_opacity, 0, n * sizeof(float), st);
    }
    if (tiles_x == 0 || tiles_y == 0 || num_bins_rows == 0) return check_launch(__func__);
    GI2D_REQUIRE(gaussian_ids_sorted && tile_bins && xys && conics && colors && v_output, "null pointer");
    GI2D_REQUIRE(tiles_x * kTile >= img_width && tiles_y * kTile >= img_height,
                 "tile grid does not cover the image (tiles are 16x16)");
    dim3 grid(tiles_x, tiles_y);
    if (v_opacity)
        raster_bwd_kernel<8, true><<<grid, 256, 0, st>>>(tiles_x, img_width, img_height,
                                                         gaussian_ids_sorted, tile_bins, num_bins_rows,
                                                         xys, conics, colors, opacities, v_output, v_xy,
                                                         v_conic, v_colors, v_opacity);
    else
        raster_bwd_kernel<8, false><<<grid, 256, 0, st>>>(tiles_x, img_width, img_height,
                                                          gaussian_ids_sorted, tile_bins, num_bins_rows,
                                                          xys, conics, colors, opacities, v_output, v_xy,
                                                          v_conic, v_colors, v_opacity);
    return check_launch(__func__);
}
