// gi2d_common.cuh -- shared device helpers for libgi2d (sm_100a).
//
// Floating-point discipline: every operation that decides an INTEGER (radius, tile bbox,
// cull tests) is spelled with explicit round-to-nearest intrinsics in the exact order the
// reference's -O3 build executes them (SASS of csrc/foward2d.cu:192-288 and
// csrc/helpers.cuh:16-50,179-206 for sm_100a), so the compiler can neither contract nor
// re-associate them.  That is what makes radii / num_tiles_hit / tile ranges bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/gi2d.h"

namespace gi2d {

constexpr int kTile = GI2D_TILE;
constexpr int kTilePixels = kTile * kTile;
constexpr int kMaxPerTile = GI2D_MAX_PER_TILE;

void set_error(const char *fmt, ...);
int check_launch(const char *what);

#define GI2D_REQUIRE(cond, msg)                         \
    do {                                                \
        if (!(cond)) {                                  \
            gi2d::set_error("%s: %s", __func__, msg);   \
            return GI2D_ERR_INVALID;                    \
        }                                               \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- (int) cast of the reference: cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0). SURVEY Q9.
__device__ __forceinline__ int f2i_rz(float v) { return __float2int_rz(v); }

// ---- helpers.cuh:179-206 compute_cov2d_bounds, op-for-op as the -O3 SASS:
//   det = fma(x, z, -(y*y));  inv = 1/det (IEEE);  conic = (z*inv, y*(-inv), x*inv)
//   b = (x+z)*0.5;  t = max(fma(b,b,-det), 0.1);  s = sqrt(t);  v1 = b+s, v2 = b-s
//   radius = ceil(clip * sqrt(max|min(v1,v2)))
struct Cov2dBounds {
    float a, b, c;   // conic
    float rx, ry;    // radius.x (major), radius.y (minor); may be NaN / inf
    bool ok;         // false when det == 0
};

__device__ __forceinline__ Cov2dBounds cov2d_bounds(float sx, float sxy, float sy, float clip_coe) {
    Cov2dBounds o;
    const float det = __fmaf_rn(sx, sy, -__fmul_rn(sxy, sxy));
    o.ok = !(det == 0.f);
    if (!o.ok) {
        o.a = o.b = o.c = 0.f;
        o.rx = o.ry = 0.f;
        return o;
    }
    const float inv = __frcp_rn(det);
    o.a = __fmul_rn(sy, inv);
    o.b = __fmul_rn(sxy, -inv);
    o.c = __fmul_rn(sx, inv);
    const float hb = __fmul_rn(__fadd_rn(sx, sy), 0.5f);
    const float t = fmaxf(__fmaf_rn(hb, hb, -det), 0.1f);
    const float s = __fsqrt_rn(t);
    const float v1 = __fadd_rn(hb, s);
    const float v2 = __fsub_rn(hb, s);
    o.rx = ceilf(__fmul_rn(__fsqrt_rn(fmaxf(v1, v2)), clip_coe));
    o.ry = ceilf(__fmul_rn(__fsqrt_rn(fminf(v1, v2)), clip_coe));
    return o;
}

// ---- helpers.cuh:16-50 get_bbox/get_tile_bbox: /16 is an exact scaling, so
//   min = clamp((int)(c/16 - r/16), 0, tb),  max = clamp((int)((c/16 + r/16) + 1), 0, tb)
struct TileBox {
    int x0, y0, x1, y1;  // inclusive min, exclusive max, in tiles
};

__device__ __forceinline__ TileBox tile_bbox(float cx, float cy, float radius, int tiles_x, int tiles_y) {
    const float r16 = __fmul_rn(radius, 0.0625f);
    const float tx = __fmul_rn(cx, 0.0625f), ty = __fmul_rn(cy, 0.0625f);
    TileBox t;
    t.x0 = min(max(0, f2i_rz(__fsub_rn(tx, r16))), tiles_x);
    t.x1 = min(max(0, f2i_rz(__fadd_rn(__fadd_rn(tx, r16), 1.f))), tiles_x);
    t.y0 = min(max(0, f2i_rz(__fsub_rn(ty, r16))), tiles_y);
    t.y1 = min(max(0, f2i_rz(__fadd_rn(__fadd_rn(ty, r16), 1.f))), tiles_y);
    return t;
}

// ---- the pair evaluation of csrc/forward.cu:652-661, op-for-op as the -O3 SASS:
//   sigma = fma(dy, b*dx, 0.5*fma(dx, a*dx, dy*(c*dy)));  vis = ex2(-sigma*log2e)
// ex2.approx.ftz differs from the reference's guarded ex2.approx only for results below
// 2^-126, which both fail alpha >= 1/255.
__device__ __forceinline__ float pair_sigma(float a, float b, float c, float dx, float dy) {
    const float q = __fmaf_rn(dx, __fmul_rn(a, dx), __fmul_rn(dy, __fmul_rn(c, dy)));
    return __fmaf_rn(dy, __fmul_rn(b, dx), __fmul_rn(q, 0.5f));
}

__device__ __forceinline__ float fast_exp_neg(float sigma) {
    float r;
    const float t = __fmul_rn(sigma, -1.4426950216293334961f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

constexpr float kAlphaMin = 1.f / 255.f;

// ---- 8-bit target sample -> float exactly as torchvision's ToTensor (u8 / 255, IEEE division;
// utils.py:21-27): one Newton step on x * fl(1/255).  Equal to __fdiv_rn(x, 255.f) for all 256 inputs
// (checked exhaustively, tests/test_oracle_golden.py) without the divide's special-case branch.
__device__ __forceinline__ float u8_to_unit(uint8_t v) {
    const float x = (float)v, r = 0.0039215688593685626984f;
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q, 255.f, x), r, q);
}

__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// ---- programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may start while
// its predecessor on the stream is still draining; everything before pdl_wait() (index math, shared
// memory zeroing, loads of data no predecessor writes) overlaps the predecessor's tail.  pdl_wait()
// returns once the predecessor grid has completed and its writes are visible.  Every thread calls it
// before touching dependent data (and before any early return, which keeps the chain transitive).
// RULE: data produced by a predecessor in the chain must be read with COHERENT loads (__ldcg / plain
// ld.global), never through the read-only path (__ldg, or a `const T *__restrict__` the compiler may turn
// into ld.global.nc): the kernel's lifetime starts BEFORE those writes, so "read-only for the lifetime of
// the kernel" does not hold and a stale L1/texture line can be served (seen as a once-in-many-steps glitch
// whenever the intersection list shifted between two steps).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Captured launches carry the attribute too: with the 3-kernel step the programmatic edges make a graph
    // replay 5-6 % faster (28.6 vs 30.2 us at 768x512 / 5k Gaussians, driver 580.159).  (With the earlier
    // 5-kernel step they were slower, 42.9 vs 37.5 us, and had been left out of captures.)
    // GI2D_NO_PDL: fully serialised launches everywhere (debugging); GI2D_GRAPH_NO_PDL: only in captures.
    static const bool no_pdl = getenv("GI2D_NO_PDL") != nullptr;
    static const bool graph_no_pdl = getenv("GI2D_GRAPH_NO_PDL") != nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (graph_no_pdl) cudaStreamIsCapturing(st, &cap);
    cfg.numAttrs = (no_pdl || cap != cudaStreamCaptureStatusNone) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace gi2d
