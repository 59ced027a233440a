// gi2d_raster_core.cuh -- per-tile rasterization building blocks shared by the stand-alone
// rasterize kernels (gi2d_raster.cu) and the fused fit step (gi2d_fit.cu).
//
// Forward  : pixel-parallel.  One thread per pixel of the 16x16 tile; the tile's (at most 256)
//            Gaussians are staged once in shared memory as 2 x float4 + float so the inner
//            loop is two broadcast LDS.128 per pair and the pair math of forward.cu:652-668.
//            Accumulation runs in ascending sorted order, exactly like the reference, so the
//            image is bit-reproducible.
// Backward : Gaussian-parallel.  Each warp owns whole Gaussians (round-robin over the tile's
//            list); its 32 lanes own 8 pixels each of the tile (v_out staged planar in shared
//            memory), sweep the 256 pixels in 8 steps accumulating the 8 (9) gradient components
//            in registers, and only then reduce across lanes -- once per (tile, Gaussian)
//            instead of once per (warp, Gaussian) as csrc/backward.cu:1322-1345 does -- with a
//            transposed (reduce-scatter) butterfly: 9 shuffles for the 8 components instead of
//            40, after which 8 lanes each issue ONE red.global.add: 8 atomics per
//            (tile, Gaussian) versus the reference's 9 per (warp, Gaussian) = 72.
#pragma once
#include "gi2d_common.cuh"

namespace gi2d {

// One staged Gaussian as the sweeps see it.
struct GaussRec {
    float x, y, a, b, c, op, r, g, bl;
};

// Shared-memory staging, general form (stand-alone API: arbitrary opacity).
struct TileGaussians {
    float4 xyab[kMaxPerTile];   // x, y, conic.a, conic.b
    float4 corg[kMaxPerTile];   // conic.c, opacity, r, g
    float bl[kMaxPerTile];      // b
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
    __device__ __forceinline__ GaussRec get(int t) const {
        const float4 p0 = xyab[t], p1 = crgb[t];
        return GaussRec{p0.x, p0.y, p0.z, p0.w, p1.x, 1.f, p1.y, p1.z, p1.w};
    }
};

// Stage one Gaussian by gathering the reference-layout arrays.  The caller syncs.
__device__ __forceinline__ void stage_gaussian(TileGaussians &s, int slot, int g,
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
}

// Forward sweep for one pixel over the staged Gaussians.  `last` receives the rank of the last
// contributor (unchanged if none).
template <class Store>
__device__ __forceinline__ void forward_sweep(const Store &s, int cnt, float px, float py,
                                              float &r, float &g, float &b, int &last) {
#pragma unroll 4
    for (int t = 0; t < cnt; ++t) {
        const GaussRec q = s.get(t);
        const float dx = __fsub_rn(q.x, px), dy = __fsub_rn(q.y, py);
        const float sigma = pair_sigma(q.a, q.b, q.c, dx, dy);
        const float alpha = fminf(1.f, __fmul_rn(q.op, fast_exp_neg(sigma)));
        if (sigma < 0.f || alpha < kAlphaMin) continue;
        r = __fmaf_rn(alpha, q.r, r);
        g = __fmaf_rn(alpha, q.g, g);
        b = __fmaf_rn(alpha, q.bl, b);
        last = t;
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

// dL/d(out) of the tile's 256 pixels, planar in shared memory: v[c][row*16 + col]; pixels outside
// the image hold 0 and are flagged in the lane's `inside` mask.
struct TileGrad {
    float v[3][kTilePixels];
};

// Pixel ownership of a lane in the backward sweep: column = lane & 15, rows (lane>>4) + 2*s, s=0..7.
// Consecutive lanes read consecutive shared-memory words: conflict free.
struct LanePixels {
    unsigned inside;   // bit s set when pixel s is inside the image
    float px, py0;     // pixel x, and y of step 0 (y of step s = py0 + 2 s)
    int base;          // (lane>>4)*16 + (lane&15): index of the step-0 pixel in TileGrad
};

__device__ __forceinline__ LanePixels lane_pixels(int tile_x, int tile_y, int img_w, int img_h) {
    const int lane = threadIdx.x & 31;
    const int px = tile_x * kTile + (lane & 15);
    const int py0 = tile_y * kTile + (lane >> 4);
    LanePixels lp;
    lp.px = (float)px;
    lp.py0 = (float)py0;
    lp.base = (lane >> 4) * kTile + (lane & 15);
    lp.inside = 0;
#pragma unroll
    for (int st = 0; st < 8; ++st)
        if (px < img_w && py0 + 2 * st < img_h) lp.inside |= 1u << st;
    return lp;
}

// Accumulate the gradient of ONE staged Gaussian over the lane's 8 pixels into acc[0..8):
//   acc = { v_x, v_y, v_a, v_b, v_c, v_r, v_g, v_b }  (+ *acc_op for opacity when kOpacity)
// Same validity rule as backward.cu:1273-1283 (sigma>=0, alpha>=1/255); the `<= final_idx`
// condition of the reference is implied by the 256-per-tile cap.
template <bool kOpacity, class Store>
__device__ __forceinline__ void backward_accumulate(const Store &s, int t, const LanePixels &lp,
                                                    const TileGrad &tg, float (&acc)[8], float *acc_op) {
    const GaussRec q = s.get(t);
    const float dx = __fsub_rn(q.x, lp.px);
    // dx-only subexpressions are shared by the 8 rows (bit-identical to recomputing them)
    const float adx = __fmul_rn(q.a, dx);
    const float bdx = __fmul_rn(q.b, dx);
    float ax = 0.f, ay = 0.f, aa = 0.f, ab = 0.f, ac = 0.f, ar = 0.f, ag = 0.f, abl = 0.f, ao = 0.f;
#pragma unroll
    for (int st = 0; st < 8; ++st) {
        const float py = lp.py0 + (float)(2 * st);
        const float dy = __fsub_rn(q.y, py);
        const float qq = __fmaf_rn(dx, adx, __fmul_rn(dy, __fmul_rn(q.c, dy)));
        const float sigma = __fmaf_rn(dy, bdx, __fmul_rn(qq, 0.5f));
        const float vis = fast_exp_neg(sigma);
        const float alpha = fminf(1.f, __fmul_rn(q.op, vis));
        const bool valid = ((lp.inside >> st) & 1u) && !(sigma < 0.f || alpha < kAlphaMin);
        if (!__any_sync(0xffffffffu, valid)) continue;
        if (valid) {
            const int pi = lp.base + 2 * kTile * st;
            const float vr = tg.v[0][pi], vg = tg.v[1][pi], vb = tg.v[2][pi];
            ar = fmaf(alpha, vr, ar);
            ag = fmaf(alpha, vg, ag);
            abl = fmaf(alpha, vb, abl);
            const float v_alpha = fmaf(q.bl, vb, fmaf(q.g, vg, q.r * vr));
            const float vva = vis * v_alpha;
            const float v_sigma = -q.op * vva;
            if (kOpacity) ao += vva;
            const float t1 = v_sigma * dx, t2 = v_sigma * dy;
            aa = fmaf(t1, dx, aa);
            ab = fmaf(t1, dy, ab);
            ac = fmaf(t2, dy, ac);
            ax = fmaf(v_sigma, fmaf(q.b, dy, adx), ax);
            ay = fmaf(v_sigma, fmaf(q.c, dy, bdx), ay);
        }
    }
    acc[0] = ax;
    acc[1] = ay;
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
template <bool kOpacity, int kWarps, class Store, class GradAddr>
__device__ __forceinline__ void backward_tile(const Store &s, const int *s_ids, int cnt, const LanePixels &lp,
                                              const TileGrad &tg, GradAddr grad_of, float *v_opacity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t < cnt; t += kWarps) {
        float acc[8];
        float op = 0.f;
        backward_accumulate<kOpacity>(s, t, lp, tg, acc, &op);
        const float total = warp_reduce_scatter8(acc);
        if ((lane & 3) == 0 && total != 0.f) atomicAdd(grad_of(s_ids[t], lane >> 2), total);
        if (kOpacity) {
            op = warp_sum(op);
            if (lane == 0 && op != 0.f) atomicAdd(v_opacity + s_ids[t], op);
        }
    }
}

}  // namespace gi2d
