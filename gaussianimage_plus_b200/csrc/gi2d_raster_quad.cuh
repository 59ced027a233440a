// gi2d_raster_quad.cuh -- warp-owned rasterization of a 16x16 tile for the fused fit step (round 2).
//
// What the ncu profile of the round-1 kernel said (profiles/r01_ncu_full_*_v8_raw.csv): 62 warp instructions per
// warp-pair, issue slots busy 68 % of the active cycles, FMA pipe 25 %, ALU pipe 37 %, block barriers the top
// stall.  And what tools/ubench/fp32_issue.cu measured on the B200: integer/compare/select (ALU-pipe)
// instructions issue at HALF rate, while the packed fma.rn.f32x2 (SASS FFMA2) delivers the full FP32 peak
// (74.1 of 74.4 TFLOP/s) with half the issue slots of FFMA.  Hence this design:
//
// Geometry : the tile is 4 QUADRANTS of 8x8 pixels.  A warp owns whole quadrants (all 4: one warp per tile;
//            2: the top / bottom half, two warps per tile; 1: four warps per tile); lane L owns the pixel PAIR (L&7, L>>3) and
//            (L&7, (L>>3)+4) of each of its quadrants and keeps it in ONE 64-bit register as f32x2, so every
//            floating-point instruction of the sweeps is packed and works on 64 pixels of a quadrant at once.
//            A Gaussian's record is staged in shared memory as it comes (32 bytes, 2 broadcast LDS.128); a
//            scalar enters the packed math through the scalar-broadcast operand form of FADD2/FMUL2/FFMA2.
// Forward  : the warp walks the tile's list in ascending order (bit-reproducible image: per pixel the same
//            operations in the same order as forward.cu:652-668), skipping quadrants outside the Gaussian's
//            alpha >= 1/255 reach box (warp-uniform mask).  A rejected pair gets weight 0 instead of a branch.
// Loss     : the lane that rendered a pixel also evaluates its loss gradient; dL/d(out) of its 8 (4) pixels
//            never leaves the registers -- no shared-memory image, no block barrier between the passes.
// Backward : FOUR Gaussians per warp at a time, 8 lanes each, sweeping pixel rows (see quad_backward4 below): the
//            cross-lane reduction spans 8 lanes and serves 4 Gaussians at once, every lane ends with exactly one
//            (Gaussian, component) total -> one RED instruction per group of four.
// Block-wide synchronisation: one barrier after the list has been rank-sorted into shared memory (and, with four
// warps per tile, one more where dL/d(out) changes hands between the pixel-owning and the Gaussian-owning warps).
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

// Reach of one Gaussian with opacity 1 inside a tile (same conservative box as reach_mask(): the axis-aligned
// bounding box of {sigma <= ln 255}, inflated by 0.1 % + 1e-3 px): the quadrants it can touch (bits 0..3) and
// the first / last tile row (0..15) it can touch.  (gx,gy) relative to the tile origin.
struct QuadReach {
    unsigned mask;
    int row_lo, row_hi;
    int col_lo, col_hi;
};

__device__ __forceinline__ QuadReach reach_quad(float gx, float gy, float a, float b, float c) {
    QuadReach o;
    o.mask = 0xFu;
    o.row_lo = 0;
    o.row_hi = kTile - 1;
    o.col_lo = 0;
    o.col_hi = kTile - 1;
    const float det = fmaf(a, c, -b * b);
    const float L = 5.5412635f * 1.001f + 1e-3f;   // ln(255)
    if (!(det > 1e-3f * a * c && a > 0.f && c > 0.f)) return o;   // indefinite / NaN / needle: no culling
    const float k = __fdividef(2.f * L, det);
    const float hx = sqrt_approx(k * c) * 1.001f + 1e-3f;
    const float hy = sqrt_approx(k * a) * 1.001f + 1e-3f;
    if (!(hx < 1e30f && hy < 1e30f)) return o;
    const float x0 = gx - hx, x1 = gx + hx, y0 = gy - hy, y1 = gy + hy;
    const unsigned cols = ((x0 <= 7.f && x1 >= 0.f) ? 1u : 0u) | ((x0 <= 15.f && x1 >= 8.f) ? 2u : 0u);
    const unsigned rows = ((y0 <= 7.f && y1 >= 0.f) ? 1u : 0u) | ((y0 <= 15.f && y1 >= 8.f) ? 2u : 0u);
    // bit q = col bit (q&1) AND row bit (q>>1)
    o.mask = ((rows & 1u) ? cols : 0u) | ((rows & 2u) ? (cols << 2) : 0u);
    // pixel rows r with y0 <= r <= y1 (pixel centres sit on integers, forward.cu:596-597)
    o.row_lo = (int)fminf(fmaxf(ceilf(y0), 0.f), 15.f);
    o.row_hi = (int)fminf(fmaxf(floorf(y1), 0.f), 15.f);
    o.col_lo = (int)fminf(fmaxf(ceilf(x0), 0.f), 15.f);
    o.col_hi = (int)fminf(fmaxf(floorf(x1), 0.f), 15.f);
    return o;
}

// Shared-memory staging of the tile's list in rank order: the 32-byte projected record of gi2d_fit.cu verbatim.
// The sweeps turn a scalar into an f32x2 operand with pk2(v, v): ptxas folds that into the scalar-broadcast
// operand form of FADD2 / FMUL2 / FFMA2, so no register is spent on the copy.
struct QuadRecords {
    float4 xyab[kMaxPerTile];   // x y a b
    float4 crgb[kMaxPerTile];   // c r g b
    unsigned char mask[kMaxPerTile];
    unsigned char rows[kMaxPerTile];   // row_lo | row_hi << 4
    unsigned char cols[kMaxPerTile];   // col_lo | col_hi << 4
    // tile-wide backward (4 warps per tile): the sweep shape of the entry over the whole tile, packed
    // lo | c0 << 4 | lg << 8 | trips << 10, and its place in the backward's list, bucket | position << 2 (0xffff: none)
    unsigned short shape[kMaxPerTile];
    unsigned short place[kMaxPerTile];
    int bucket_count[4];
};

__device__ __forceinline__ void stage_quad(QuadRecords &s, int slot, float4 p0, float4 p1, float tile_x0,
                                           float tile_y0) {
    s.xyab[slot] = p0;
    s.crgb[slot] = p1;
    const QuadReach rc = reach_quad(p0.x - tile_x0, p0.y - tile_y0, p0.z, p0.w, p1.x);
    s.mask[slot] = (unsigned char)rc.mask;
    s.rows[slot] = (unsigned char)(rc.row_lo | (rc.row_hi << 4));
    s.cols[slot] = (unsigned char)(rc.col_lo | (rc.col_hi << 4));
}

// Geometry of a warp that owns kNQ quadrants: kNQ = 4 all of them (2 columns x 2 rows), 2 one ROW of quadrants
// (2 columns), 1 a single quadrant.  Quadrant qi of the warp sits at column qi % kCols, row qi / kCols
// relative to the warp's first quadrant.
template <int kNQ>
struct QuadGeom {
    static constexpr int kCols = kNQ >= 2 ? 2 : 1;
    static constexpr int kRows = kNQ == 4 ? 2 : 1;
};

// Per-lane constants: minus the pixel coordinates of the lane's pairs.
template <int kNQ>
struct QuadLane {
    f32x2 npx[QuadGeom<kNQ>::kCols];   // (-px, -px) per quadrant column
    f32x2 npy[QuadGeom<kNQ>::kRows];   // (-py, -(py+4)) per quadrant row
};

// (px0, py0): pixel of this lane in the warp's first quadrant, pair element 0
template <int kNQ>
__device__ __forceinline__ QuadLane<kNQ> quad_lane(int px0, int py0) {
    QuadLane<kNQ> g;
#pragma unroll
    for (int k = 0; k < QuadGeom<kNQ>::kCols; ++k) {
        const float px = (float)(px0 + 8 * k);
        g.npx[k] = pk2(-px, -px);
    }
#pragma unroll
    for (int j = 0; j < QuadGeom<kNQ>::kRows; ++j) {
        const float py = (float)(py0 + 8 * j);
        g.npy[j] = pk2(-py, -(py + 4.f));
    }
    return g;
}

// Everything of one staged Gaussian the forward sweep shares between the warp's quadrants.
template <int kNQ>
struct QuadGauss {
    f32x2 dx[QuadGeom<kNQ>::kCols], adx[QuadGeom<kNQ>::kCols], bdx[QuadGeom<kNQ>::kCols];
    f32x2 dy[QuadGeom<kNQ>::kRows], dycdy[QuadGeom<kNQ>::kRows];
    f32x2 r, g, b;
};

template <int kNQ>
__device__ __forceinline__ void quad_setup(float4 p0, float4 p1, const QuadLane<kNQ> &ln, QuadGauss<kNQ> &q);

template <int kNQ>
__device__ __forceinline__ void quad_load(const QuadRecords &s, int t, const QuadLane<kNQ> &ln, QuadGauss<kNQ> &q) {
    quad_setup<kNQ>(s.xyab[t], s.crgb[t], ln, q);
}

template <int kNQ>
__device__ __forceinline__ void quad_setup(float4 p0, float4 p1, const QuadLane<kNQ> &ln, QuadGauss<kNQ> &q) {
    const f32x2 xx = pk2(p0.x, p0.x), yy = pk2(p0.y, p0.y), aa = pk2(p0.z, p0.z), bb = pk2(p0.w, p0.w);
    const f32x2 cc = pk2(p1.x, p1.x);
    q.r = pk2(p1.y, p1.y);
    q.g = pk2(p1.z, p1.z);
    q.b = pk2(p1.w, p1.w);
#pragma unroll
    for (int k = 0; k < QuadGeom<kNQ>::kCols; ++k) {
        q.dx[k] = add2(xx, ln.npx[k]);        // x - px   (forward.cu:654)
        q.adx[k] = mul2(aa, q.dx[k]);
        q.bdx[k] = mul2(bb, q.dx[k]);
    }
#pragma unroll
    for (int j = 0; j < QuadGeom<kNQ>::kRows; ++j) {
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

// sigma and the masked weight of a pixel pair:
//   sigma = fma(dy, b*dx, 0.5*fma(dx, a*dx, dy*(c*dy)))  (the -O3 SASS order of forward.cu:655-657),
//   vis = ex2(-sigma*log2e);  weight = vis when sigma >= 0 and vis >= 1/255 (forward.cu:659), else 0.
__device__ __forceinline__ f32x2 pair_weight(f32x2 dx, f32x2 adx, f32x2 bdx, f32x2 dy, f32x2 dycdy) {
    const f32x2 half2 = pk2(0.5f, 0.5f), nl2e = pk2(-1.4426950216293334961f, -1.4426950216293334961f);
    const f32x2 qq = fma2(dx, adx, dycdy);
    const f32x2 sg = fma2(dy, bdx, mul2(qq, half2));
    const f32x2 tt = mul2(sg, nl2e);
    float s0, s1, t0, t1, v0, v1;
    unpk2(sg, s0, s1);
    unpk2(tt, t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v1) : "f"(t1));
    return pk2(accept_weight(s0, v0), accept_weight(s1, v1));
}

// Forward sweep of a warp over the staged list: acc*[qi] += weight * colour for its kNQ quadrants.
// `qshift`: bit position of the warp's first quadrant in the reach masks.  (Pixels outside the image
// accumulate like any other; nobody reads them.)
template <int kNQ>
__device__ __forceinline__ void quad_forward(const QuadRecords &s, int cnt, const QuadLane<kNQ> &ln, int qshift,
                                             f32x2 (&accR)[kNQ], f32x2 (&accG)[kNQ], f32x2 (&accB)[kNQ]) {
    constexpr int kCols = QuadGeom<kNQ>::kCols;
    // (not unrolled: the kernel is instruction-cache bound when several CTAs of an SM sit in different phases --
    // ncu: 0.8 'no instruction' stalls per issue with the compiler's 4x unrolling; 2040x1356 -9 %, 8192^2 -13 %)
#ifndef GI2D_FWD_ILP
#define GI2D_FWD_ILP 64   // list length from which the 4-warp kernel's forward sweep takes two entries per trip (0: never)
#endif
#if GI2D_FWD_ILP > 0
    // long lists (densification clusters new Gaussians: one tile can hold 10x the average, and its CTA ends up
    // alone on its SM, one warp per scheduler, paying every Gaussian's ~150-cycle dependency chain in full): two
    // entries per trip, both evaluated whenever either reaches the quadrant (outside its reach box a Gaussian's
    // weight is 0 by the alpha test, and acc + 0 * colour == acc exactly), accumulated in list order
    // (measured with one 170-entry tile among 1536: rasterizer 35.4 -> 31.1 us; no change in the steady state)
    if (kNQ == 1 && cnt > GI2D_FWD_ILP) {
        int t = 0;
#pragma unroll 1
        for (; t + 1 < cnt; t += 2) {
            const unsigned m0 = ((unsigned)s.mask[t] >> qshift) & 1u, m1 = ((unsigned)s.mask[t + 1] >> qshift) & 1u;
            if (!(m0 | m1)) continue;   // warp-uniform
            QuadGauss<kNQ> q0, q1;
            quad_setup<kNQ>(s.xyab[t], s.crgb[t], ln, q0);
            quad_setup<kNQ>(s.xyab[t + 1], s.crgb[t + 1], ln, q1);
            const f32x2 w0 = pair_weight(q0.dx[0], q0.adx[0], q0.bdx[0], q0.dy[0], q0.dycdy[0]);
            const f32x2 w1 = pair_weight(q1.dx[0], q1.adx[0], q1.bdx[0], q1.dy[0], q1.dycdy[0]);
            accR[0] = fma2(w1, q1.r, fma2(w0, q0.r, accR[0]));
            accG[0] = fma2(w1, q1.g, fma2(w0, q0.g, accG[0]));
            accB[0] = fma2(w1, q1.b, fma2(w0, q0.b, accB[0]));
        }
        if (t < cnt && (((unsigned)s.mask[t] >> qshift) & 1u)) {
            QuadGauss<kNQ> q;
            quad_load<kNQ>(s, t, ln, q);
            const f32x2 w = pair_weight(q.dx[0], q.adx[0], q.bdx[0], q.dy[0], q.dycdy[0]);
            accR[0] = fma2(w, q.r, accR[0]);
            accG[0] = fma2(w, q.g, accG[0]);
            accB[0] = fma2(w, q.b, accB[0]);
        }
        return;
    }
#endif
#ifdef GI2D_FWD_PREFETCH
    // software pipeline: the next entry's mask and record are in flight while this one is evaluated (a warp
    // that is alone on its scheduler -- the long tile at the end of the grid -- otherwise pays the shared-memory
    // round trip on top of every Gaussian's dependency chain)
    if (cnt <= 0) return;
    unsigned m_next = s.mask[0];
    float4 p0n = s.xyab[0], p1n = s.crgb[0];
#pragma unroll 1
    for (int t = 0; t < cnt; ++t) {
        const unsigned m = (m_next >> qshift) & ((1u << kNQ) - 1u);
        const float4 p0 = p0n, p1 = p1n;
        const int tn = min(t + 1, cnt - 1);
        m_next = s.mask[tn];
        p0n = s.xyab[tn];
        p1n = s.crgb[tn];
        if (!m) continue;   // warp-uniform
        QuadGauss<kNQ> q;
        quad_setup<kNQ>(p0, p1, ln, q);
#else
#pragma unroll 1
    for (int t = 0; t < cnt; ++t) {
        const unsigned m = ((unsigned)s.mask[t] >> qshift) & ((1u << kNQ) - 1u);
        if (!m) continue;   // warp-uniform
        QuadGauss<kNQ> q;
        quad_load<kNQ>(s, t, ln, q);
#endif
#pragma unroll
        for (int qi = 0; qi < kNQ; ++qi) {
            if (kNQ > 1 && !((m >> qi) & 1u)) continue;   // warp-uniform
            const int k = qi % kCols, j = qi / kCols;
            const f32x2 w = pair_weight(q.dx[k], q.adx[k], q.bdx[k], q.dy[j], q.dycdy[j]);
            accR[qi] = fma2(w, q.r, accR[qi]);   // out += alpha * rgb, ascending order (forward.cu:662-666)
            accG[qi] = fma2(w, q.g, accG[qi]);
            accB[qi] = fma2(w, q.b, accB[qi]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Backward sweep, group layout: FOUR Gaussians per warp at a time, 8 lanes each.
//
// A warp that walks the list one Gaussian at a time with its lanes on pixels pays per (warp, Gaussian) a 32-lane
// reduction (9 shuffles + 14 selects + 7 adds, a chain of five dependent shuffle round trips), the record load,
// the loop control and the atomic set-up: ~95 instructions around ~45 of useful pair math at the list lengths of
// a 768x512 / 5000-Gaussian scene.  Here every 8-lane group (L>>3) works on its OWN Gaussian and sweeps that
// Gaussian's OWN reach box inside the region (kRows rows of 16 pixels): one instruction still evaluates 64
// pairs, but the fixed part is paid once per FOUR Gaussians, the reduction spans 8 lanes (7 shuffles for 4
// Gaussians at once), and after it every one of the 32 lanes holds exactly one (Gaussian, component) total: one
// RED instruction.
//
// Shape of a group's sweep (all of it per-lane DATA, so four different shapes run in one warp without a
// divergent branch): the 8 lanes hold 8 pixel pairs (x, x+1) laid out as
//     8 pairs x 1 row   (reach wider than 8 columns)          1 row  per trip
//     4 pairs x 2 rows  (reach within 8 columns)              2 rows per trip
//     2 pairs x 4 rows  (reach within 4 columns)              4 rows per trip
// starting at the reach box's first row / first (even) column.  The trip count of the warp is the largest of its
// four groups; the list is ordered by trip count (build_group_list) so that the four are alike.  A group that
// runs past its own rows reads dL/d(out) from an all-zero extra row (its weights there are below 1/255 anyway).
// At 768x512 / 5000 Gaussians a reach box covers ~9 of a tile's 16 rows and ~9 of its 16 columns: sweeping the
// union of four boxes over all 16 columns (the first version of this kernel) evaluated ~2x the pairs.
template <int kRows>
struct WarpGrad {                      // dL/d(out) of the region, planar, row stride 16; row kRows stays zero
    float v[3][kRows + 1][kTile];
};

struct GroupShape {
    int lo;      // first region row
    int c0;      // first column (even)
    int lg;      // log2(pairs per row): 3, 2, 1
    int trips;   // ceil(rows / rows per trip); 0: nothing to sweep
};

template <int kRows>
__device__ __forceinline__ GroupShape group_shape(unsigned rw, unsigned cw, int region_row0) {
    GroupShape g;
    g.lo = max((int)(rw & 15u) - region_row0, 0);
    const int hi = min((int)(rw >> 4) - region_row0, kRows - 1);
    const int chi = (int)(cw >> 4);
    g.c0 = (int)(cw & 14u);
    const int w = chi - g.c0;   // columns c0 .. c0 + w
    if (w < 4) { g.lg = 1; g.c0 = min(g.c0, 12); }
    else if (w < 8) { g.lg = 2; g.c0 = min(g.c0, 8); }
    else { g.lg = 3; g.c0 = 0; }
    g.trips = max(0, (hi - g.lo + (8 >> g.lg)) >> (3 - g.lg));
    return g;
}

// The staged Gaussians that reach the region (quadrant bits `region_bits`), longest sweep first: four buckets of
// trip counts (> 8, 5-8, 3-4, 1-2) filled by ballots.  Every lane of the warp calls it; returns the list length.
template <int kRows>
__device__ __forceinline__ int build_group_list(const QuadRecords &s, int cnt, unsigned region_bits, int region_row0,
                                                unsigned char *list) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    auto bucket_of = [&](int t) -> int {
        if (t >= cnt || !((unsigned)s.mask[t] & region_bits)) return -1;
        const int trips = group_shape<kRows>(s.rows[t], s.cols[t], region_row0).trips;
        return trips > 8 ? 0 : (trips > 4 ? 1 : (trips > 2 ? 2 : (trips > 0 ? 3 : -1)));
    };
    int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    int b_first = bucket_of(lane);   // (lists of up to 32 entries -- the common case -- classify once)
    for (int base = 0; base < cnt; base += 32) {
        const int b = base == 0 ? b_first : bucket_of(base + lane);
        n0 += __popc(__ballot_sync(0xffffffffu, b == 0));
        n1 += __popc(__ballot_sync(0xffffffffu, b == 1));
        n2 += __popc(__ballot_sync(0xffffffffu, b == 2));
        n3 += __popc(__ballot_sync(0xffffffffu, b == 3));
    }
    int o0 = 0, o1 = n0, o2 = n0 + n1, o3 = n0 + n1 + n2;
    const int n = o3 + n3;
    for (int base = 0; base < cnt; base += 32) {
        const int b = base == 0 ? b_first : bucket_of(base + lane);
        const unsigned m0 = __ballot_sync(0xffffffffu, b == 0), m1 = __ballot_sync(0xffffffffu, b == 1);
        const unsigned m2 = __ballot_sync(0xffffffffu, b == 2), m3 = __ballot_sync(0xffffffffu, b == 3);
        const unsigned mine = b == 0 ? m0 : (b == 1 ? m1 : (b == 2 ? m2 : m3));
        const int off = b == 0 ? o0 : (b == 1 ? o1 : (b == 2 ? o2 : o3));
        if (b >= 0) list[off + __popc(mine & lt)] = (unsigned char)(base + lane);
        o0 += __popc(m0); o1 += __popc(m1); o2 += __popc(m2); o3 += __popc(m3);
    }
    return n;
}

__device__ __forceinline__ float group_reduce_scatter8(float (&v)[8]) {
    // transposed butterfly over lane bits 2,1,0: on return lane j (= L&7) of every group holds component j
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

struct GroupAcc {
    f32x2 ax, ay, sa, sb, sc, ar, ag, abl;
};

// one trip of a lane: the pixel pair (px, px+1) of the row at byte offset `off` of the region (fy = its y)
template <int kRows, bool kGuard>
__device__ __forceinline__ void group_trip(GroupAcc &A, const char *wg_base, int off, int zoff, float fy, bool in0,
                                           bool in1, float y, float c, f32x2 dx, f32x2 adx, f32x2 bdx, f32x2 cr,
                                           f32x2 cg, f32x2 cb) {
    constexpr int kPlane = (kRows + 1) * kTile * 4;
    const float dyf = y - fy;
    const float dycdyf = dyf * (c * dyf);
    const f32x2 dy = pk2(dyf, dyf);
    f32x2 w = pair_weight(dx, adx, bdx, dy, pk2(dycdyf, dycdyf));
    if (kGuard) {   // a NaN of a pixel that does not exist must not reach the sums
        float w0, w1;
        unpk2(w, w0, w1);
        w = pk2(in0 ? w0 : 0.f, in1 ? w1 : 0.f);
    }
    const char *src = wg_base + min(off, zoff);   // past the region: the zero row
    const f32x2 vr = *reinterpret_cast<const f32x2 *>(src);
    const f32x2 vg = *reinterpret_cast<const f32x2 *>(src + kPlane);
    const f32x2 vb = *reinterpret_cast<const f32x2 *>(src + 2 * kPlane);
    A.ar = fma2(w, vr, A.ar);     // v_rgb += alpha * v_out
    A.ag = fma2(w, vg, A.ag);
    A.abl = fma2(w, vb, A.abl);
    const f32x2 va = fma2(cb, vb, fma2(cg, vg, mul2(cr, vr)));   // v_alpha = rgb . v_out
    const f32x2 u = mul2(w, va);   // vis * v_alpha = -v_sigma
    const f32x2 t1x = mul2(u, dx), t2y = mul2(u, dy);
    A.sa = fma2(t1x, dx, A.sa);
    A.sb = fma2(t1x, dy, A.sb);
    A.sc = fma2(t2y, dy, A.sc);
    A.ax = add2(A.ax, t1x);
    A.ay = add2(A.ay, t2y);
}

// Tile-wide variant of the staging (4 warps per tile): the entry's sweep shape over the whole tile is computed ONCE
// here (the backward reads it back instead of decoding rows / cols per group) and the entry takes its place in one
// of the four trip-count buckets of the CTA's ONE list (a shared-memory counter per bucket; order inside a bucket is
// irrelevant).  finish_wide_list() turns (bucket, position) into list slots once every entry has been staged.
__device__ __forceinline__ unsigned pack_shape(const GroupShape &g) {
    return (unsigned)g.lo | ((unsigned)g.c0 << 4) | ((unsigned)g.lg << 8) | ((unsigned)g.trips << 10);
}
__device__ __forceinline__ GroupShape unpack_shape(unsigned v) {
    GroupShape g;
    g.lo = (int)(v & 15u);
    g.c0 = (int)((v >> 4) & 15u);
    g.lg = (int)((v >> 8) & 3u);
    g.trips = (int)(v >> 10);
    return g;
}

__device__ __forceinline__ void stage_quad_wide(QuadRecords &s, int slot, float4 p0, float4 p1, float tile_x0,
                                                float tile_y0) {
    stage_quad(s, slot, p0, p1, tile_x0, tile_y0);
    const GroupShape g = group_shape<kTile>(s.rows[slot], s.cols[slot], 0);
    s.shape[slot] = (unsigned short)pack_shape(g);
    unsigned place = 0xffffu;
    if (s.mask[slot] != 0 && g.trips > 0) {
        const int b = g.trips > 8 ? 0 : (g.trips > 4 ? 1 : (g.trips > 2 ? 2 : 3));
        place = (unsigned)b | ((unsigned)atomicAdd(&s.bucket_count[b], 1) << 2);
    }
    s.place[slot] = (unsigned short)place;
}

// after the block barrier that ends the staging: every thread writes the list slots of ranks tid, tid + threads, ...
// Returns the list length (the same in every thread).
__device__ __forceinline__ int finish_wide_list(const QuadRecords &s, int cnt, unsigned char *list) {
    const int n0 = s.bucket_count[0], n1 = s.bucket_count[1], n2 = s.bucket_count[2], n3 = s.bucket_count[3];
    for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
        const unsigned pl = s.place[r];
        if (pl == 0xffffu) continue;
        const int b = (int)(pl & 3u), pos = (int)(pl >> 2);
        const int off = b == 0 ? 0 : (b == 1 ? n0 : (b == 2 ? n0 + n1 : n0 + n1 + n2));
        list[off + pos] = (unsigned char)r;
    }
    return n0 + n1 + n2 + n3;
}

// `list`: ranks of the staged Gaussians that reach the region (n of them, build_group_list order).  The warp
// takes the groups first_group, first_group + group_stride, ...  (region = the warp's own rows: 0, 1; region =
// the whole tile shared by the CTA's warps: warp, #warps).
template <int kRows, bool kGuard, bool kPacked = false>
__device__ __forceinline__ void quad_backward4(const QuadRecords &s, const int *s_ids, const unsigned char *list,
                                               int n, int first_group, int group_stride, int tile_px0,
                                               int region_py0, int region_row0, int img_w, int rows_inside,
                                               const WarpGrad<kRows> &wg, float *__restrict__ grads) {
    const int lane = threadIdx.x & 31, j = lane & 7, grp = lane >> 3;
    const char *wg_base = reinterpret_cast<const char *>(&wg.v[0][0][0]);
    for (int base = 4 * first_group; base < n; base += 4 * group_stride) {
        const bool valid = base + grp < n;
        const int t = list[valid ? base + grp : n - 1];
        const float4 p0 = s.xyab[t], p1 = s.crgb[t];
        GroupShape sh = kPacked ? unpack_shape(s.shape[t]) : group_shape<kRows>(s.rows[t], s.cols[t], region_row0);
        if (!valid) sh.trips = 0;
        const int T = __reduce_max_sync(0xffffffffu, sh.trips);   // warp-uniform
        const int rpt = 8 >> sh.lg;                                // rows per trip
        const int col = sh.c0 + 2 * (j & ((1 << sh.lg) - 1));
        int row = sh.lo + (j >> sh.lg);
        const int px = tile_px0 + col;
        bool in0 = true, in1 = true;
        if (kGuard) {
            in0 = px < img_w;
            in1 = px + 1 < img_w;
        }
        const f32x2 xx = pk2(p0.x, p0.x), aa = pk2(p0.z, p0.z), bb = pk2(p0.w, p0.w);
        const f32x2 cr = pk2(p1.y, p1.y), cg = pk2(p1.z, p1.z), cb = pk2(p1.w, p1.w);
        const f32x2 npx = pk2(-(float)px, -((float)px + 1.f));
        const f32x2 dx = add2(xx, npx), adx = mul2(aa, dx), bdx = mul2(bb, dx);
        float fy = (float)(region_py0 + row);
        const float frpt = (float)rpt;
        int off = (row * kTile + col) * 4;
        const int zoff = (kRows * kTile + col) * 4, doff = rpt * kTile * 4;
#ifdef GI2D_BWD_SINGLE_ACC
        // (variant: ONE accumulator set -- 16 registers less, one dependency chain per warp)
        GroupAcc A{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
        const GroupAcc B{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
#pragma unroll 1
        for (int i = 0; i < T; ++i) {
            bool y0 = in0, y1 = in1;
            if (kGuard) {
                const bool ra = row < rows_inside;
                y0 = in0 && ra; y1 = in1 && ra;
                row += rpt;
            }
            group_trip<kRows, kGuard>(A, wg_base, off, zoff, fy, y0, y1, p0.y, p1.x, dx, adx, bdx, cr, cg, cb);
            off += doff;
            fy += frpt;
        }
#else
        // two trips per loop iteration with separate accumulators: two independent dependency chains per warp
        GroupAcc A{0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull}, B = A;
        int i = 0;
#pragma unroll 1
        for (; i + 1 < T; i += 2) {   // warp-uniform (kept rolled: instruction cache, see quad_forward)
            bool y0 = in0, y1 = in1, z0 = in0, z1 = in1;
            if (kGuard) {
                const bool ra = row < rows_inside, rb = row + rpt < rows_inside;
                y0 = in0 && ra; y1 = in1 && ra; z0 = in0 && rb; z1 = in1 && rb;
                row += 2 * rpt;
            }
            group_trip<kRows, kGuard>(A, wg_base, off, zoff, fy, y0, y1, p0.y, p1.x, dx, adx, bdx, cr, cg, cb);
            group_trip<kRows, kGuard>(B, wg_base, off + doff, zoff, fy + frpt, z0, z1, p0.y, p1.x, dx, adx, bdx, cr,
                                      cg, cb);
            off += 2 * doff;
            fy += 2.f * frpt;
        }
        if (i < T) {
            bool y0 = in0, y1 = in1;
            if (kGuard) {
                const bool ra = row < rows_inside;
                y0 = in0 && ra; y1 = in1 && ra;
            }
            group_trip<kRows, kGuard>(A, wg_base, off, zoff, fy, y0, y1, p0.y, p1.x, dx, adx, bdx, cr, cg, cb);
        }
#endif
        float l0, l1, m0, m1, acc[8];
        unpk2(A.ax, l0, l1); unpk2(B.ax, m0, m1); const float sx = -((l0 + l1) + (m0 + m1));
        unpk2(A.ay, l0, l1); unpk2(B.ay, m0, m1); const float sy = -((l0 + l1) + (m0 + m1));
        acc[0] = fmaf(p0.z, sx, p0.w * sy);   // v_xy = C * sum(v_sigma * delta)
        acc[1] = fmaf(p0.w, sx, p1.x * sy);
        unpk2(A.sa, l0, l1); unpk2(B.sa, m0, m1); acc[2] = -0.5f * ((l0 + l1) + (m0 + m1));
        unpk2(A.sb, l0, l1); unpk2(B.sb, m0, m1); acc[3] = -0.5f * ((l0 + l1) + (m0 + m1));
        unpk2(A.sc, l0, l1); unpk2(B.sc, m0, m1); acc[4] = -0.5f * ((l0 + l1) + (m0 + m1));
        unpk2(A.ar, l0, l1); unpk2(B.ar, m0, m1); acc[5] = (l0 + l1) + (m0 + m1);
        unpk2(A.ag, l0, l1); unpk2(B.ag, m0, m1); acc[6] = (l0 + l1) + (m0 + m1);
        unpk2(A.abl, l0, l1); unpk2(B.abl, m0, m1); acc[7] = (l0 + l1) + (m0 + m1);
        const float total = group_reduce_scatter8(acc);
        if (valid && total != 0.f) atomicAdd(grads + 8 * (size_t)s_ids[t] + j, total);
    }
}

}  // namespace gi2d
