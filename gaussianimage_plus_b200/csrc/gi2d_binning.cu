// gi2d_binning.cu -- integer-exact binning primitives: prefix sum,
// intersection key emit, stable LSD radix sort of 64-bit keys with 32-bit payload, tile edges.
// These are the stand-alone, reference-shaped entry points (`compute_cumulative_intersects`,
// `map_gaussian_to_intersects`, torch.sort/gather, `get_tile_bin_edges`).  The fused,
// sync-free binning used by the fit step lives in gi2d_fit.cu and shares the ranking code.
//
// All of it is HBM/latency bound integer work: coalesced 128-bit accesses where the layout
// allows, shared-memory privatised histograms, warp match_any ranking -- no tensor cores.
#include "gi2d_scan.cuh"

namespace gi2d {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                       // per thread
constexpr int kTileItems = kThreads * kItems;   // 2048 per CTA

// ------------------------------------------------------------------ prefix sum (inclusive)
__global__ void __launch_bounds__(kThreads) scan_reduce_kernel(int n, const int32_t *__restrict__ in,
                                                               int32_t *__restrict__ block_sums) {
    __shared__ int s_warp[kThreads / 32];
    const int base = blockIdx.x * kTileItems + threadIdx.x * kItems;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) sum += in[base + k];
    int total;
    block_scan_inclusive<kThreads>(sum, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: in-place exclusive scan of block_sums[0..m)
__global__ void __launch_bounds__(1024) scan_sums_kernel(int m, int32_t *__restrict__ block_sums) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int start = 0; start < m; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = i < m ? block_sums[i] : 0;
        int total;
        const int incl = block_scan_inclusive<1024>(v, s_warp, &total);
        const int carry = s_carry;
        if (i < m) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) scan_apply_kernel(int n, const int32_t *__restrict__ in,
                                                              const int32_t *__restrict__ block_offsets,
                                                              int32_t *__restrict__ out,
                                                              int32_t *__restrict__ total_out) {
    __shared__ int s_warp[kThreads / 32];
    const int base = blockIdx.x * kTileItems + threadIdx.x * kItems;
    int v[kItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
        v[k] = sum;
    }
    const int incl = block_scan_inclusive<kThreads>(sum, s_warp, nullptr);
    const int off = (block_offsets ? block_offsets[blockIdx.x] : 0) + incl - sum;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) {
            out[base + k] = v[k] + off;
            if (total_out && base + k == n - 1) *total_out = v[k] + off;
        }
}

int launch_cumsum(int n, const int32_t *in, int32_t *out, int32_t *total, int32_t *block_sums,
                  cudaStream_t st) {
    const int nb = cdiv(n, kTileItems);
    if (nb > 1) {
        scan_reduce_kernel<<<nb, kThreads, 0, st>>>(n, in, block_sums);
        scan_sums_kernel<<<1, 1024, 0, st>>>(nb, block_sums);
    }
    scan_apply_kernel<<<nb, kThreads, 0, st>>>(n, in, nb > 1 ? block_sums : nullptr, out, total);
    return check_launch("cumsum");
}

// ------------------------------------------------------------------ key emit
// forward.cu:141-206: one thread per Gaussian walks its tile bbox row-major.
__global__ void __launch_bounds__(kThreads)
map_isects_kernel(int n, const float *__restrict__ xys, const float *__restrict__ depths,
                  const int32_t *__restrict__ radii, const int32_t *__restrict__ cum_tiles_hit,
                  int tiles_x, int tiles_y, float radius_clip, int64_t *__restrict__ isect_ids,
                  int32_t *__restrict__ gaussian_ids) {
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= n) return;
    const int r = radii[idx];
    if ((float)r < radius_clip) return;  // int radius promoted to float, forward.cu:161
    const float2 c = reinterpret_cast<const float2 *>(xys)[idx];
    const TileBox box = tile_bbox(c.x, c.y, (float)r, tiles_x, tiles_y);
    int cur = idx == 0 ? 0 : cum_tiles_hit[idx - 1];
    // (int64_t)*(int32_t*)&depth : sign-extended bit pattern, forward.cu:187
    const int64_t depth_id = (int64_t)__float_as_int(depths[idx]);
    for (int ty = box.y0; ty < box.y1; ++ty)
        for (int tx = box.x0; tx < box.x1; ++tx) {
            const int64_t tile_id = (int64_t)(ty * tiles_x + tx);
            isect_ids[cur] = (tile_id << 32) | depth_id;
            gaussian_ids[cur] = idx;
            ++cur;
        }
}

// ------------------------------------------------------------------ radix sort (8-bit digits)
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr uint64_t kSignFlip = 0x8000000000000000ull;

__device__ __forceinline__ int digit_of(int64_t key, int shift, int bits) {
    return (int)(((uint64_t)key ^ kSignFlip) >> shift) & ((1 << bits) - 1);
}

// counts[d * nblocks + b] = number of keys of CTA b whose digit is d
__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(int n, const int32_t *__restrict__ n_dev, const int64_t *__restrict__ keys, int shift,
                  int bits, int nblocks, int32_t *__restrict__ counts) {
    __shared__ int s_hist[kRadix];
    if (n_dev) n = min(n, *n_dev);
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kTileItems;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int i = base + k * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_hist[digit_of(keys[i], shift, bits)], 1);
    }
    __syncthreads();
    counts[threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// Stable scatter.  Items are taken warp-striped (warp w owns items [w*256, w*256+256) of the CTA
// tile, 32 consecutive per round), ranked with match_any inside the round, per-warp digit
// counters across rounds, then an exclusive scan across warps per digit.
__global__ void __launch_bounds__(kThreads)
radix_scatter_kernel(int n, const int32_t *__restrict__ n_dev, const int64_t *__restrict__ keys_in,
                     const int32_t *__restrict__ vals_in, int64_t *__restrict__ keys_out,
                     int32_t *__restrict__ vals_out, int shift, int bits, int nblocks,
                     const int32_t *__restrict__ counts, const int32_t *__restrict__ counts_incl) {
    constexpr int kWarps = kThreads / 32;
    if (n_dev) n = min(n, *n_dev);
    __shared__ int s_cnt[kWarps][kRadix];
    __shared__ int s_base[kRadix];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kWarps * kRadix; i += kThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int wbase = blockIdx.x * kTileItems + warp * (kItems * 32);
    int64_t key[kItems];
    int rank[kItems];
    int dig[kItems];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        rank[r] = 0;
        dig[r] = 0;
        if (valid) {
            key[r] = keys_in[i];
            const int d = digit_of(key[r], shift, bits);
            dig[r] = d;
            const unsigned peers = __match_any_sync(act, d);
            const int leader = __ffs(peers) - 1;
            int old = 0;
            if (lane == leader) {
                old = s_cnt[warp][d];
                s_cnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(peers, old, leader);
            rank[r] = old + __popc(peers & lt_mask);
        }
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: exclusive scan over warps for digit d, plus the global base of (d, CTA)
        const int d = tid;
        int run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = s_cnt[w][d];
            s_cnt[w][d] = run;
            run += c;
        }
        const int gi = d * nblocks + blockIdx.x;
        s_base[d] = counts_incl[gi] - counts[gi];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int i = wbase + r * 32 + lane;
        if (i < n) {
            const int pos = s_base[dig[r]] + s_cnt[warp][dig[r]] + rank[r];
            keys_out[pos] = key[r];
            if (vals_in) vals_out[pos] = vals_in[i];
        }
    }
}

// ------------------------------------------------------------------ tile edges
// forward.cu:211-233
__global__ void __launch_bounds__(kThreads)
tile_edges_kernel(int n, const int32_t *__restrict__ n_dev, const int64_t *__restrict__ sorted,
                  int32_t *__restrict__ tile_bins, int rows) {
    if (n_dev) n = min(n, *n_dev);
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= n) return;
    const int cur = (int)(sorted[idx] >> 32);
    const bool cur_ok = cur >= 0 && cur < rows;
    if (idx == 0 && cur_ok) tile_bins[2 * cur] = 0;
    if (idx == n - 1 && cur_ok) tile_bins[2 * cur + 1] = n;
    if (idx == 0) return;
    const int prev = (int)(sorted[idx - 1] >> 32);
    if (prev != cur) {
        if (prev >= 0 && prev < rows) tile_bins[2 * prev + 1] = idx;
        if (cur_ok) tile_bins[2 * cur] = idx;
    }
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// inclusive prefix sum for the fit step's per-tile overlap counts (gi2d_fit.cu)
int cumsum_i32_launch(int n, const int32_t *in, int32_t *out, int32_t *total, int32_t *block_sums,
                      cudaStream_t st) {
    return launch_cumsum(n, in, out, total, block_sums, st);
}
size_t cumsum_i32_workspace(int n) { return gi2d_scan_workspace_size(n); }

}  // namespace gi2d

using namespace gi2d;

extern "C" size_t gi2d_scan_workspace_size(int num_points) {
    if (num_points <= 0) return 256;
    return align_up((size_t)cdiv(num_points, kTileItems) * sizeof(int32_t));
}

extern "C" int gi2d_cumsum_i32(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit,
                               int32_t *total, void *workspace, size_t workspace_bytes,
                               gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_points == 0) {
        if (total) cudaMemsetAsync(total, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    GI2D_REQUIRE(num_tiles_hit && cum_tiles_hit, "null pointer");
    if (workspace_bytes < gi2d_scan_workspace_size(num_points) || !workspace) {
        set_error("%s: workspace too small", __func__);
        return GI2D_ERR_WORKSPACE;
    }
    return launch_cumsum(num_points, num_tiles_hit, cum_tiles_hit, total, (int32_t *)workspace, st);
}

extern "C" int gi2d_map_gaussian_to_intersects(int num_points, const float *xys, const float *depths,
                                               const int32_t *radii, const int32_t *cum_tiles_hit,
                                               int tiles_x, int tiles_y, float radius_clip,
                                               int64_t *isect_ids, int32_t *gaussian_ids,
                                               gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(xys && depths && radii && cum_tiles_hit && isect_ids && gaussian_ids, "null pointer");
    map_isects_kernel<<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        num_points, xys, depths, radii, cum_tiles_hit, tiles_x, tiles_y, radius_clip, isect_ids,
        gaussian_ids);
    return check_launch(__func__);
}

extern "C" size_t gi2d_sort_workspace_size(int num_items) {
    if (num_items <= 0) return 256;
    const size_t nb = (size_t)cdiv(num_items, kTileItems);
    const size_t mat = align_up(nb * kRadix * sizeof(int32_t));
    return 2 * mat + gi2d_scan_workspace_size((int)(nb * kRadix)) +
           align_up((size_t)num_items * sizeof(int64_t)) + align_up((size_t)num_items * sizeof(int32_t));
}

extern "C" int gi2d_sort_pairs_i64(int num_items, const int64_t *keys_in, const int32_t *vals_in,
                                   int64_t *keys_out, int32_t *vals_out, int begin_bit, int end_bit,
                                   void *workspace, size_t workspace_bytes, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_items >= 0, "num_items < 0");
    GI2D_REQUIRE(begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "bad bit range");
    if (num_items == 0) return GI2D_OK;
    GI2D_REQUIRE(keys_in && vals_in && keys_out && vals_out, "null pointer");
    if (!workspace || workspace_bytes < gi2d_sort_workspace_size(num_items)) {
        set_error("%s: workspace too small", __func__);
        return GI2D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = cdiv(num_items, kTileItems);
    const size_t mat = align_up((size_t)nb * kRadix * sizeof(int32_t));
    char *w = (char *)workspace;
    int32_t *counts = (int32_t *)w;  w += mat;
    int32_t *incl = (int32_t *)w;    w += mat;
    int32_t *scan_ws = (int32_t *)w; w += gi2d_scan_workspace_size(nb * kRadix);
    int64_t *tmp_keys = (int64_t *)w; w += align_up((size_t)num_items * sizeof(int64_t));
    int32_t *tmp_vals = (int32_t *)w;
    const int passes = cdiv(end_bit - begin_bit, kRadixBits);
    if (passes == 0) {
        cudaMemcpyAsync(keys_out, keys_in, (size_t)num_items * sizeof(int64_t), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(vals_out, vals_in, (size_t)num_items * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
        return check_launch(__func__);
    }
    const int64_t *src_k = keys_in;
    const int32_t *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = (end_bit - shift) < kRadixBits ? (end_bit - shift) : kRadixBits;
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        int64_t *dst_k = to_out ? keys_out : tmp_keys;
        int32_t *dst_v = to_out ? vals_out : tmp_vals;
        radix_hist_kernel<<<nb, kThreads, 0, st>>>(num_items, nullptr, src_k, shift, bits, nb, counts);
        const int rc = launch_cumsum(nb * kRadix, counts, incl, nullptr, scan_ws, st);
        if (rc != GI2D_OK) return rc;
        radix_scatter_kernel<<<nb, kThreads, 0, st>>>(num_items, nullptr, src_k, src_v, dst_k, dst_v, shift,
                                                     bits, nb, counts, incl);
        src_k = dst_k;
        src_v = dst_v;
    }
    return check_launch(__func__);
}

extern "C" int gi2d_get_tile_bin_edges(int num_intersects, const int64_t *isect_ids_sorted,
                                       int32_t *tile_bins, int num_bins_rows, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_intersects >= 0 && num_bins_rows >= 0, "negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_bins_rows > 0) {
        GI2D_REQUIRE(tile_bins, "null tile_bins");
        cudaMemsetAsync(tile_bins, 0, (size_t)num_bins_rows * 2 * sizeof(int32_t), st);
    }
    if (num_intersects > 0 && num_bins_rows > 0) {
        GI2D_REQUIRE(isect_ids_sorted, "null keys");
        tile_edges_kernel<<<cdiv(num_intersects, kThreads), kThreads, 0, st>>>(
            num_intersects, nullptr, isect_ids_sorted, tile_bins, num_bins_rows);
    }
    return check_launch(__func__);
}
