// gi2d_scan.cuh -- warp/block prefix-sum building blocks (integer exact).
#pragma once
#include "gi2d_common.cuh"

namespace gi2d {

__device__ __forceinline__ int warp_scan_inclusive(int v) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// Inclusive scan of one value per thread across a CTA of kThreads (multiple of 32, <= 1024).
// `s_warp` must hold kThreads/32 ints.  Returns the inclusive prefix; *total gets the CTA sum.
template <int kThreads>
__device__ __forceinline__ int block_scan_inclusive(int v, int *s_warp, int *total) {
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_scan_inclusive(v);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kWarps ? s_warp[lane] : 0;
        w = warp_scan_inclusive(w);
        if (lane < kWarps) s_warp[lane] = w;
    }
    __syncthreads();
    if (warp > 0) incl += s_warp[warp - 1];
    if (total) *total = s_warp[kWarps - 1];
    __syncthreads();  // s_warp may be reused by the caller
    return incl;
}

}  // namespace gi2d
