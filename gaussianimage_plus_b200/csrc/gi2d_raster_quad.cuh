// gi2d_raster_quad.cuh -- warp-owned rasterization of a 16x16 tile for the fused fit step (round 2).
//
// What the ncu profile of the round-1 kernel said (profiles/r01_ncu_full_*_v8_raw.csv): 62 warp instructions per
// warp-pair, issue slots busy 68 % of the active cycles, FMA pipe 25 %, ALU pipe 37 %, block barriers the top
// stall.  And what tools/ubench/fp32_issue.cu measured on the B200: integer/compare/select (ALU-pipe)
// instructions issue at HALF rate, while the packed fma.rn.f32x2 (SASS FFMA2) delivers the full FP32 peak
// (74.1 of 74.4 TFLOP/s) with half the issue slots of FFMA.  Hence this design:
//
// Geometry : the tile is 4 QUADRANTS of 8x8 pixels.  A warp owns whole quadrants (all 4: one warp per tile;
//            or 2: the top / bottom half, two warps per tile); lane L owns the pixel PAIR (L&7, L>>3) and
//            (L&7, (L>>3)+4) of each of its quadrants and keeps it in ONE 64-bit register as f32x2, so every
//            floating-point instruction of the sweeps is packed and works on 64 pixels of a quadrant at once.
//            A Gaussian's record is staged in shared memory with every value DUPLICATED (x,x | y,y | ...):
//            4 broadcast LDS.128 deliver ready-made f32x2 operands, no register moves.
// Forward  : the warp walks the tile's list in ascending order (bit-reproducible image: per pixel the same
//            operations in the same order as forward.cu:652-668), skipping quadrants outside the Gaussian's
//            alpha >= 1/255 reach box (warp-uniform mask).  A rejected pair gets weight 0 instead of a branch.
// Loss     : the lane that rendered a pixel also evaluates its loss gradient; dL/d(out) of its 8 (4) pixels
//            never leaves the registers -- no shared-memory image, no block barrier between the passes.
// Backward : the same walk again; per Gaussian 8 packed accumulators, then the 9-shuffle transposed
//            reduce-scatter (gi2d_raster_core.cuh) and 8 red.global per (warp, Gaussian).
// The only block-wide synchronisation is the one after the list has been rank-sorted into shared memory.
#pragma once
#include "gi2d_raster_core.cuh"

namespace gi2d {

typedef unsigned long long f32x2;   // two floats in one 64-bit register: .x = pixel row r, .y = pixel row r+4

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

constexpr int kQuads = 4;   // 8x8-pixel quadrants per tile: bit q <-> (x0, y0) = (8*(q&1), 8*(q>>1))

// Quadrant reach mask of one Gaussian with opacity 1 (same conservative box as reach_mask(): the axis-aligned
// bounding box of {sigma <= ln 255}, inflated by 0.1 % + 1e-3 px).  (gx,gy) relative to the tile origin.
__device__ __forceinline__ unsigned reach_mask_quad(float gx, float gy, float a, float b, float c) {
    const float det = fmaf(a, c, -b * b);
    const float L = 5.5412635f * 1.001f + 1e-3f;   // ln(255)
    if (!(det > 1e-3f * a * c && a > 0.f && c > 0.f)) return 0xFu;   // indefinite / NaN / needle: no culling
    const float k = __fdividef(2.f * L, det);
    const float hx = sqrt_approx(k * c) * 1.001f + 1e-3f;
    const float hy = sqrt_approx(k * a) * 1.001f + 1e-3f;
    if (!(hx < 1e30f && hy < 1e30f)) return 0xFu;
    const float x0 = gx - hx, x1 = gx + hx, y0 = gy - hy, y1 = gy + hy;
    const unsigned cols = ((x0 <= 7.f && x1 >= 0.f) ? 1u : 0u) | ((x0 <= 15.f && x1 >= 8.f) ? 2u : 0u);
    const unsigned rows = ((y0 <= 7.f && y1 >= 0.f) ? 1u : 0u) | ((y0 <= 15.f && y1 >= 8.f) ? 2u : 0u);
    // bit q = col bit (q&1) AND row bit (q>>1)
    return ((rows & 1u) ? cols : 0u) | ((rows & 2u) ? (cols << 2) : 0u);
}

// Shared-memory staging: every value duplicated, so that one LDS.128 yields two f32x2 operands.
struct QuadRecords {
    float4 xy[kMaxPerTile];   // x x y y
    float4 ab[kMaxPerTile];   // a a b b
    float4 cr[kMaxPerTile];   // c c r r
    float4 gb[kMaxPerTile];   // g g b b   (colour g, colour b)
    unsigned char mask[kMaxPerTile];
};

__device__ __forceinline__ void stage_quad(QuadRecords &s, int slot, float4 p0, float4 p1, float tile_x0,
                                           float tile_y0) {
    s.xy[slot] = make_float4(p0.x, p0.x, p0.y, p0.y);
    s.ab[slot] = make_float4(p0.z, p0.z, p0.w, p0.w);
    s.cr[slot] = make_float4(p1.x, p1.x, p1.y, p1.y);
    s.gb[slot] = make_float4(p1.z, p1.z, p1.w, p1.w);
    s.mask[slot] = (unsigned char)reach_mask_quad(p0.x - tile_x0, p0.y - tile_y0, p0.z, p0.w, p1.x);
}

__device__ __forceinline__ void lds_pair(const float4 *p, f32x2 &u, f32x2 &v) {
    const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(p);
    u = w.x;
    v = w.y;
}

// Per-lane constants: minus the pixel coordinates of the lane's pairs.  kNQ quadrants owned by the warp
// (4: all; 2: the row of quadrants `half`), i.e. 2 columns x kNQ/2 quadrant rows.
template <int kNQ>
struct QuadLane {
    f32x2 npx[2];         // (-px, -px) for the left / right quadrant column
    f32x2 npy[kNQ / 2];   // (-py, -(py+4)) per owned quadrant row
};

template <int kNQ>
__device__ __forceinline__ QuadLane<kNQ> quad_lane(int tile_px0, int tile_py0, int half) {
    const int lane = threadIdx.x & 31;
    QuadLane<kNQ> g;
    const float px = (float)(tile_px0 + (lane & 7));
#pragma unroll
    for (int k = 0; k < 2; ++k) g.npx[k] = pk2(-(px + 8.f * k), -(px + 8.f * k));
#pragma unroll
    for (int j = 0; j < kNQ / 2; ++j) {
        const float py = (float)(tile_py0 + 8 * (half + j) + (lane >> 3));
        g.npy[j] = pk2(-py, -(py + 4.f));
    }
    return g;
}

// Everything of one staged Gaussian the sweeps share between the quadrants.
template <int kNQ>
struct QuadGauss {
    f32x2 dx[2], adx[2], bdx[2];
    f32x2 dy[kNQ / 2], dycdy[kNQ / 2];
    f32x2 r, g, b;
};

template <int kNQ>
__device__ __forceinline__ void quad_load(const QuadRecords &s, int t, const QuadLane<kNQ> &ln, QuadGauss<kNQ> &q) {
    f32x2 xx, yy, aa, bb, cc;
    lds_pair(&s.xy[t], xx, yy);
    lds_pair(&s.ab[t], aa, bb);
    lds_pair(&s.cr[t], cc, q.r);
    lds_pair(&s.gb[t], q.g, q.b);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        q.dx[k] = add2(xx, ln.npx[k]);        // x - px   (forward.cu:654)
        q.adx[k] = mul2(aa, q.dx[k]);
        q.bdx[k] = mul2(bb, q.dx[k]);
    }
#pragma unroll
    for (int j = 0; j < kNQ / 2; ++j) {
        q.dy[j] = add2(yy, ln.npy[j]);
        q.dycdy[j] = mul2(q.dy[j], mul2(cc, q.dy[j]));
    }
}

// weight of one pair: vis when sigma >= 0 and vis >= 1/255 (NaN passes both, like the reference's
// `if (sigma < 0.f || alpha < 1.f / 255.f) continue;`), else 0.  Two chained predicate compares + one select:
// compare / select instructions go through the half-rate ALU pipe, so every one of them counts.
__device__ __forceinline__ float accept_weight(float sigma, float vis) {
    float w;
    asm("{\n .reg .pred p, q;\n setp.geu.f32 q, %1, 0f00000000;\n setp.geu.and.f32 p, %2, 0f3B808081, q;\n"
        " selp.f32 %0, %2, 0f00000000, p;\n}"
        : "=f"(w)
        : "f"(sigma), "f"(vis));
    return w;
}

// sigma and the masked weight of the lane's pixel pair in quadrant (column k, row j):
//   sigma = fma(dy, b*dx, 0.5*fma(dx, a*dx, dy*(c*dy)))  (the -O3 SASS order of forward.cu:655-657),
//   vis = ex2(-sigma*log2e);  weight = vis when sigma >= 0 and vis >= 1/255 (forward.cu:659), else 0.
// kGuard: lanes whose pixel lies outside the image (bits of `outside`: 2*quadrant + pair element) get weight 0,
// so that a NaN of a pixel that does not exist cannot reach the gradient sums.
template <int kNQ, bool kGuard>
__device__ __forceinline__ f32x2 quad_weight(const QuadGauss<kNQ> &q, int k, int j, unsigned outside, int qi) {
    const f32x2 half2 = pk2(0.5f, 0.5f), nl2e = pk2(-1.4426950216293334961f, -1.4426950216293334961f);
    const f32x2 qq = fma2(q.dx[k], q.adx[k], q.dycdy[j]);
    const f32x2 sg = fma2(q.dy[j], q.bdx[k], mul2(qq, half2));
    const f32x2 tt = mul2(sg, nl2e);
    float s0, s1, t0, t1, v0, v1;
    unpk2(sg, s0, s1);
    unpk2(tt, t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v1) : "f"(t1));
    float w0 = accept_weight(s0, v0), w1 = accept_weight(s1, v1);
    if (kGuard) {
        if ((outside >> (2 * qi)) & 1u) w0 = 0.f;
        if ((outside >> (2 * qi + 1)) & 1u) w1 = 0.f;
    }
    return pk2(w0, w1);
}

// Forward sweep of a warp over the staged list: acc*[qi] += weight * colour for its kNQ quadrants.
// `qshift`: bit position of the warp's first quadrant in the reach masks (0, or 2 for the bottom half).
template <int kNQ, bool kGuard>
__device__ __forceinline__ void quad_forward(const QuadRecords &s, int cnt, const QuadLane<kNQ> &ln, int qshift,
                                             unsigned outside, f32x2 (&accR)[kNQ], f32x2 (&accG)[kNQ],
                                             f32x2 (&accB)[kNQ]) {
    for (int t = 0; t < cnt; ++t) {
        const unsigned m = ((unsigned)s.mask[t] >> qshift) & ((1u << kNQ) - 1u);
        if (!m) continue;   // warp-uniform
        QuadGauss<kNQ> q;
        quad_load<kNQ>(s, t, ln, q);
#pragma unroll
        for (int qi = 0; qi < kNQ; ++qi) {
            if (!((m >> qi) & 1u)) continue;   // warp-uniform
            const f32x2 w = quad_weight<kNQ, kGuard>(q, qi & 1, qi >> 1, outside, qi);
            accR[qi] = fma2(w, q.r, accR[qi]);   // out += alpha * rgb, ascending order (forward.cu:662-666)
            accG[qi] = fma2(w, q.g, accG[qi]);
            accB[qi] = fma2(w, q.b, accB[qi]);
        }
    }
}

// Backward sweep: per staged Gaussian accumulate over the warp's quadrants, reduce across the warp, and add the
// 8 components {v_x, v_y, v_a, v_b, v_c, v_r, v_g, v_b} to grads[8*id + k] (backward.cu:1273-1345).
// vR/vG/vB: dL/d(out) of the lane's pixel pairs (0 outside the image).
template <int kNQ, bool kGuard>
__device__ __forceinline__ void quad_backward(const QuadRecords &s, const int *s_ids, int cnt,
                                              const QuadLane<kNQ> &ln, int qshift, unsigned outside,
                                              const f32x2 (&vR)[kNQ], const f32x2 (&vG)[kNQ],
                                              const f32x2 (&vB)[kNQ], float *__restrict__ grads) {
    const int lane = threadIdx.x & 31;
    for (int t = 0; t < cnt; ++t) {
        const unsigned m = ((unsigned)s.mask[t] >> qshift) & ((1u << kNQ) - 1u);
        if (!m) continue;   // warp-uniform
        QuadGauss<kNQ> q;
        quad_load<kNQ>(s, t, ln, q);
        f32x2 ax = 0ull, ay = 0ull, aa = 0ull, ab = 0ull, ac = 0ull, ar = 0ull, ag = 0ull, abl = 0ull;
#pragma unroll
        for (int qi = 0; qi < kNQ; ++qi) {
            if (!((m >> qi) & 1u)) continue;   // warp-uniform
            const int k = qi & 1, j = qi >> 1;
            const f32x2 w = quad_weight<kNQ, kGuard>(q, k, j, outside, qi);
            ar = fma2(w, vR[qi], ar);     // v_rgb += alpha * v_out
            ag = fma2(w, vG[qi], ag);
            abl = fma2(w, vB[qi], abl);
            const f32x2 va = fma2(q.b, vB[qi], fma2(q.g, vG[qi], mul2(q.r, vR[qi])));   // v_alpha = rgb . v_out
            const f32x2 u = mul2(w, va);  // vis * v_alpha = -v_sigma
            const f32x2 t1 = mul2(u, q.dx[k]), t2 = mul2(u, q.dy[j]);
            aa = fma2(t1, q.dx[k], aa);
            ab = fma2(t1, q.dy[j], ab);
            ac = fma2(t2, q.dy[j], ac);
            ax = add2(ax, t1);
            ay = add2(ay, t2);
        }
        float l0, l1;
        float acc[8];
        // sums over the pair, with the sign of v_sigma = -u restored
        unpk2(ax, l0, l1); const float sx = -(l0 + l1);
        unpk2(ay, l0, l1); const float sy = -(l0 + l1);
        // conic a, b, c of this Gaussian: re-read (broadcast LDS) instead of holding six more registers
        const float ca = s.ab[t].x, cb = s.ab[t].z, cc = s.cr[t].x;
        acc[0] = fmaf(ca, sx, cb * sy);   // v_xy = C * sum(v_sigma * delta)
        acc[1] = fmaf(cb, sx, cc * sy);
        unpk2(aa, l0, l1); acc[2] = -0.5f * (l0 + l1);
        unpk2(ab, l0, l1); acc[3] = -0.5f * (l0 + l1);
        unpk2(ac, l0, l1); acc[4] = -0.5f * (l0 + l1);
        unpk2(ar, l0, l1); acc[5] = l0 + l1;
        unpk2(ag, l0, l1); acc[6] = l0 + l1;
        unpk2(abl, l0, l1); acc[7] = l0 + l1;
        const float total = warp_reduce_scatter8(acc);
        if ((lane & 3) == 0 && total != 0.f) atomicAdd(grads + 8 * (size_t)s_ids[t] + (lane >> 2), total);
    }
}

}  // namespace gi2d
