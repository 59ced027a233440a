// gi2d_raster_core.cuh -- per-tile rasterization building blocks shared by the stand-alone
// rasterize kernels (gi2d_raster.cu) and the fused fit step (gi2d_fit.cu).
//
// Geometry : a 16x16 tile is split into 8 sub-blocks of 8x4 pixels (2 across, 4 down).  A warp
//            always works on one sub-block at a time: lane L <-> pixel (L&7, L>>3) of the block.
//            While a Gaussian is staged, one thread computes an 8-bit mask of the sub-blocks its
//            alpha >= 1/255 ellipse can reach (axis-aligned bounding box of {sigma <= ln 255},
//            inflated by 0.1 % + 1e-3 px, so it is conservative under float rounding); sweeps skip
//            (Gaussian, sub-block) combinations whose bit is clear.  Those pairs would all have
//            failed the reference's per-pair test (forward.cu:659), so results do not change, but
//            35-60 % of the pair evaluations never issue.
// Forward  : pixel-parallel.  One thread per pixel; the tile's (at most 256) Gaussians are
//            staged once in shared memory as 2 x float4 (+ float) so the inner loop is two
//            broadcast LDS.128 per pair and the pair math of forward.cu:652-668, in the
//            reference's operation order.  Accumulation runs in ascending sorted order, exactly
//            like the reference, so the image is bit-reproducible.
// Backward : Gaussian-parallel.  Each warp owns whole Gaussians (round-robin over the tile's
//            list) and sweeps the tile's 8 sub-blocks (v_out staged planar in shared memory),
//            accumulating the 8 (9) gradient components in registers; only then does it reduce
//            across lanes -- once per (tile, Gaussian) instead of once per (warp, Gaussian) as
//            csrc/backward.cu:1322-1345 does -- with a transposed (reduce-scatter) butterfly:
//            9 shuffles for the 8 components instead of 40, after which 8 lanes each issue ONE
//            red.global.add: 8 atomics per (tile, Gaussian) versus the reference's 72.
#pragma once
#include "gi2d_common.cuh"

namespace gi2d {

constexpr int kBlocksPerTile = 8;   // 8x4-pixel sub-blocks
constexpr int kGradStride = 24;     // padded row stride of TileGrad: 4 rows of a sub-block hit distinct banks

// pixel of lane `lane` in sub-block `blk`, relative to the tile origin
__device__ __forceinline__ int blk_x(int blk, int lane) { return ((blk & 1) << 3) + (lane & 7); }
__device__ __forceinline__ int blk_y(int blk, int lane) { return ((blk >> 1) << 2) + (lane >> 3); }

// One staged Gaussian as the sweeps see it.
struct GaussRec {
    float x, y, a, b, c, op, r, g, bl;
};

// Sub-block reach mask of one Gaussian (see header).  (gx,gy) relative to the tile origin.
__device__ __forceinline__ unsigned reach_mask(float gx, float gy, float a, float b, float c, float op) {
    const float det = fmaf(a, c, -b * b);
    // sigma <= ln(255 * op)  <=>  alpha = op * exp(-sigma) >= 1/255   (alpha's min(1,.) never lowers it)
    const float L = __logf(255.f * op) * 1.001f + 1e-3f;
    // indefinite / NaN / extremely anisotropic (det would carry cancellation error): no culling
    if (!(det > 1e-3f * a * c && a > 0.f && c > 0.f && L > 0.f && L < 1e30f)) return 0xFFu;
    // (approximate divide / sqrt, <= 2 ulp each: far inside the 0.1 % + 1e-3 px inflation)
    const float k = __fdividef(2.f * L, det);
    const float hx = sqrt_approx(k * c) * 1.001f + 1e-3f;
    const float hy = sqrt_approx(k * a) * 1.001f + 1e-3f;
    if (!(hx < 1e30f && hy < 1e30f)) return 0xFFu;
    const float x0 = gx - hx, x1 = gx + hx, y0 = gy - hy, y1 = gy + hy;
    unsigned m = 0;
#pragma unroll
    for (int blk = 0; blk < kBlocksPerTile; ++blk) {
        const float bx0 = (float)((blk & 1) << 3), by0 = (float)((blk >> 1) << 2);
        if (x0 <= bx0 + 7.f && x1 >= bx0 && y0 <= by0 + 3.f && y1 >= by0) m |= 1u << blk;
    }
    return m;
}

// Shared-memory staging, general form (stand-alone API: arbitrary opacity).
struct TileGaussians {
    float4 xyab[kMaxPerTile];   // x, y, conic.a, conic.b
    float4 corg[kMaxPerTile];   // conic.c, opacity, r, g
    float bl[kMaxPerTile];      // b
    unsigned char mask[kMaxPerTile];
    static constexpr bool kUnitOpacity = false;
    __device__ __forceinline__ GaussRec get(int t) const {
        const float4 p0 = xyab[t], p1 = corg[t];
        return GaussRec{p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, bl[t]};
    }
};

// Shared-memory staging, fit-step form: the 32-byte projected record {x,y,a,b | c,r,g,b} of
// gi2d_fit.cu copied verbatim (opacity is the constant 1 of the live model, so 1*vis == vis and
// the results are bit-identical to the general form).
struct TileRecords {
    float4 xyab[kMaxPerTile];
    float4 crgb[kMaxPerTile];
    unsigned char mask[kMaxPerTile];
    // opacity == 1: alpha = min(1, vis) == vis for every accepted pair (sigma >= 0 => vis <= 1)
    static constexpr bool kUnitOpacity = true;
    __device__ __forceinline__ GaussRec get(int t) const {
        const float4 p0 = xyab[t], p1 = crgb[t];
        return GaussRec{p0.x, p0.y, p0.z, p0.w, p1.x, 1.f, p1.y, p1.z, p1.w};
    }
};

// Stage one Gaussian by gathering the reference-layout arrays.  The caller syncs.
__device__ __forceinline__ void stage_gaussian(TileGaussians &s, int slot, int g, float tile_x0, float tile_y0,
                                               const float *__restrict__ xys,
                                               const float *__restrict__ conics,
                                               const float *__restrict__ colors,
                                               const float *__restrict__ opacities) {
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(xys) + g);
    const float a = __ldg(conics + 3 * g), b = __ldg(conics + 3 * g + 1), c = __ldg(conics + 3 * g + 2);
    const float r = __ldg(colors + 3 * g), gg = __ldg(colors + 3 * g + 1), bb = __ldg(colors + 3 * g + 2);
    const float op = opacities ? __ldg(opacities + g) : 1.f;
    s.xyab[slot] = make_float4(xy.x, xy.y, a, b);
    s.corg[slot] = make_float4(c, op, r, gg);
    s.bl[slot] = bb;
    s.mask[slot] = (unsigned char)reach_mask(xy.x - tile_x0, xy.y - tile_y0, a, b, c, op);
}

__device__ __forceinline__ void stage_record(TileRecords &s, int slot, float4 p0, float4 p1, float tile_x0,
                                             float tile_y0) {
    s.xyab[slot] = p0;
    s.crgb[slot] = p1;
    s.mask[slot] = (unsigned char)reach_mask(p0.x - tile_x0, p0.y - tile_y0, p0.z, p0.w, p1.x, 1.f);
}

// Forward sweep for one pixel (of sub-block `blk`, a warp-uniform value) over the staged Gaussians.
// The warp first compacts, 32 Gaussians at a time, the ones whose reach mask has bit `blk` set
// (one LDS.U8 + ballot), then walks the set bits: Gaussians out of reach cost nothing, and the
// walk is warp-uniform (uniform-datapath integer work).  `last` receives the rank of the last
// contributor (unchanged if none).  Must be called by all 32 lanes of the warp.
template <class Store>
__device__ __forceinline__ void forward_sweep(const Store &s, int cnt, int blk, bool active, float px, float py,
                                              float &r, float &g, float &b, int &last) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < cnt; base += 32) {
        const unsigned m = (base + lane < cnt) ? s.mask[base + lane] : 0u;
        unsigned hit = __ballot_sync(0xffffffffu, (m >> blk) & 1u);
        while (hit) {
            const int t = base + __ffs(hit) - 1;
            hit &= hit - 1;
            const GaussRec q = s.get(t);
            const float dx = __fsub_rn(q.x, px), dy = __fsub_rn(q.y, py);
            const float sigma = pair_sigma(q.a, q.b, q.c, dx, dy);
            const float vis = fast_exp_neg(sigma);
            const float alpha = Store::kUnitOpacity ? vis : fminf(1.f, __fmul_rn(q.op, vis));
            if (!active || sigma < 0.f || alpha < kAlphaMin) continue;
            r = __fmaf_rn(alpha, q.r, r);
            g = __fmaf_rn(alpha, q.g, g);
            b = __fmaf_rn(alpha, q.bl, b);
            last = t;
        }
    }
}

// Reduce-scatter of 8 per-lane values over the warp in 9 shuffles (4+2+1 transposed steps over
// lane bits 4,3,2, then a 2-step butterfly over bits 1,0): on return EVERY lane holds the warp
// total of component (lane >> 2).  A plain butterfly would take 8 x 5 = 40.
__device__ __forceinline__ float warp_reduce_scatter8(float (&v)[8]) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
        const bool upper = (lane & (half << 2)) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half << 2);
        }
    }
    float r = v[0];
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// dL/d(out) of the tile's 256 pixels, planar in shared memory: v[c][y*24 + x]; pixels outside the
// image hold 0 and are flagged in the lane's `inside` mask.  Stride 24: the 4 rows a warp reads
// per sub-block start at banks 0, 24, 16, 8 -> conflict free.
struct TileGrad {
    float v[3][kTile * kGradStride];
};

__device__ __forceinline__ int grad_index(int x, int y) { return y * kGradStride + x; }

// Per-lane constants of the backward sweep.
struct LanePixels {
    unsigned inside;   // bit `blk` set when this lane's pixel of sub-block blk is inside the image
    float px0, py0;    // pixel coordinates of this lane's pixel in sub-block 0
    int base;          // grad_index of that pixel
};

__device__ __forceinline__ LanePixels lane_pixels(int tile_x, int tile_y, int img_w, int img_h) {
    const int lane = threadIdx.x & 31;
    const int px = tile_x * kTile + (lane & 7);
    const int py = tile_y * kTile + (lane >> 3);
    LanePixels lp;
    lp.px0 = (float)px;
    lp.py0 = (float)py;
    lp.base = grad_index(lane & 7, lane >> 3);
    lp.inside = 0;
#pragma unroll
    for (int blk = 0; blk < kBlocksPerTile; ++blk)
        if (px + ((blk & 1) << 3) < img_w && py + ((blk >> 1) << 2) < img_h) lp.inside |= 1u << blk;
    return lp;
}

// Accumulate the gradient of ONE staged Gaussian over the tile into acc[0..8):
//   acc = { v_x, v_y, v_a, v_b, v_c, v_r, v_g, v_b }  (+ *acc_op for opacity when kOpacity)
// Same validity rule as backward.cu:1273-1283 (sigma>=0, alpha>=1/255); the `<= final_idx`
// condition of the reference is implied by the 256-per-tile cap.
template <bool kOpacity, bool kFullTile, class Store>
__device__ __forceinline__ void backward_accumulate(const Store &s, int t, const LanePixels &lp,
                                                    const TileGrad &tg, float (&acc)[8], float *acc_op) {
    const GaussRec q = s.get(t);
    const unsigned reach = s.mask[t];
    // the lane's pixel column takes two values (left / right sub-block); its x-only products are shared
    const float dx0 = __fsub_rn(q.x, lp.px0), dx1 = __fsub_rn(q.x, lp.px0 + 8.f);
    const float adx0 = __fmul_rn(q.a, dx0), adx1 = __fmul_rn(q.a, dx1);
    const float bdx0 = __fmul_rn(q.b, dx0), bdx1 = __fmul_rn(q.b, dx1);
    float ax = 0.f, ay = 0.f, aa = 0.f, ab = 0.f, ac = 0.f, ar = 0.f, ag = 0.f, abl = 0.f, ao = 0.f;
#pragma unroll
    for (int blk = 0; blk < kBlocksPerTile; ++blk) {
        if (!((reach >> blk) & 1u)) continue;  // warp-uniform
        const float dx = (blk & 1) ? dx1 : dx0, adx = (blk & 1) ? adx1 : adx0, bdx = (blk & 1) ? bdx1 : bdx0;
        const float py = lp.py0 + (float)((blk >> 1) << 2);
        const float dy = __fsub_rn(q.y, py);
        const float qq = __fmaf_rn(dx, adx, __fmul_rn(dy, __fmul_rn(q.c, dy)));
        const float sigma = __fmaf_rn(dy, bdx, __fmul_rn(qq, 0.5f));
        const float vis = fast_exp_neg(sigma);
        const float alpha = Store::kUnitOpacity ? vis : fminf(1.f, __fmul_rn(q.op, vis));
        const bool valid = (kFullTile || ((lp.inside >> blk) & 1u)) && !(sigma < 0.f || alpha < kAlphaMin);
        if (!__any_sync(0xffffffffu, valid)) continue;
        if (valid) {
            const int pi = lp.base + ((blk & 1) << 3) + ((blk >> 1) << 2) * kGradStride;
            const float vr = tg.v[0][pi], vg = tg.v[1][pi], vb = tg.v[2][pi];
            ar = fmaf(alpha, vr, ar);
            ag = fmaf(alpha, vg, ag);
            abl = fmaf(alpha, vb, abl);
            const float v_alpha = fmaf(q.bl, vb, fmaf(q.g, vg, q.r * vr));
            const float vva = vis * v_alpha;
            const float v_sigma = Store::kUnitOpacity ? -vva : -q.op * vva;
            if (kOpacity) ao += vva;
            const float t1 = v_sigma * dx, t2 = v_sigma * dy;
            aa = fmaf(t1, dx, aa);
            ab = fmaf(t1, dy, ab);
            ac = fmaf(t2, dy, ac);
            // v_xy = sum v_sigma * (C delta) = C * sum(v_sigma * delta): accumulate the two sums only
            ax += t1;
            ay += t2;
        }
    }
    acc[0] = fmaf(q.a, ax, q.b * ay);
    acc[1] = fmaf(q.b, ax, q.c * ay);
    acc[2] = 0.5f * aa;
    acc[3] = 0.5f * ab;
    acc[4] = 0.5f * ac;
    acc[5] = ar;
    acc[6] = ag;
    acc[7] = abl;
    if (kOpacity) *acc_op = ao;
}

// Backward of one tile: warps take the staged Gaussians round-robin, one at a time.  After the
// reduce-scatter lane L holds component L>>2; lanes with (L&3)==0 issue the red.global.add.
// grad_of(g, k) returns the address of component k (0..7) of Gaussian g.
template <bool kOpacity, int kWarps, bool kFullTile, class Store, class GradAddr>
__device__ __forceinline__ void backward_tile(const Store &s, const int *s_ids, int cnt, const LanePixels &lp,
                                              const TileGrad &tg, GradAddr grad_of, float *v_opacity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t < cnt; t += kWarps) {
        if (s.mask[t] == 0) continue;  // the alpha >= 1/255 ellipse misses the tile altogether (warp-uniform)
        float acc[8];
        float op = 0.f;
        backward_accumulate<kOpacity, kFullTile>(s, t, lp, tg, acc, &op);
        const float total = warp_reduce_scatter8(acc);
        if ((lane & 3) == 0 && total != 0.f) atomicAdd(grad_of(s_ids[t], lane >> 2), total);
        if (kOpacity) {
            op = warp_sum(op);
            if (lane == 0 && op != 0.f) atomicAdd(v_opacity + s_ids[t], op);
        }
    }
}

}  // namespace gi2d
