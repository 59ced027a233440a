// gi2d_loss.cu -- gradient of the SSIM-family image losses of models/utils.py:60-80
// (`SSIM`, `Fusion1` = 0.7 mse + 0.3 (1-ssim), `Fusion2` = 0.7 l1 + 0.3 (1-ssim)) with respect to the
// UNCLAMPED rasterizer output, i.e. everything autograd does between `out_img` and the loss in
// gaussianimage_covariance.py:210,252-253: torch.clamp(0,1) -> pytorch_msssim.ssim(data_range=1,
// size_average=True) / mse / l1 -> backward.  The pointwise losses (L2, L1, Fusion3) never come here:
// the rasterizer evaluates them inline (gi2d_fit.cu, K4).
//
// pytorch_msssim is a third-party dependency that is NOT vendored in /root/reference and is unpinned
// (requirements.txt: "pytorch-msssim"); the algorithm restated here is the one of its `ssim()` as
// published (v1.0.0, pytorch_msssim/ssim.py `_fspecial_gauss_1d`, `gaussian_filter`, `_ssim`):
//   window  : 11 taps, sigma 1.5, exp(-(k-5)^2 / (2 sigma^2)) normalised in float32;
//   filter  : separable, VALID (no padding): the SSIM map has (H-10) x (W-10) entries per channel;
//   moments : mu1 = g*X, mu2 = g*Y, s1 = g*X^2 - mu1^2, s2 = g*Y^2 - mu2^2, s12 = g*XY - mu1 mu2;
//   map     : ((2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1)) * ((2 s12 + C2) / (s1 + s2 + C2)),
//             C1 = 0.01^2, C2 = 0.03^2;  ssim = mean over channels of the spatial means.
//
// Two kernels, both per 16x16 tile with a 5-pixel halo staged in shared memory (HBM-bound: the image,
// the target and three derivative planes per channel are each read once plus halo, 2.6x):
//   S1 ssim_stats_kernel : the 5 filtered moments -> SSIM map value (summed into the stats block)
//                          and its partials  M = d map/d mu1 (total, through s1 and s12 as well),
//                          S = d map/d s1,  T = d map/d s12  at every valid window, 0 elsewhere;
//   S2 ssim_grad_kernel  : d loss/d X[p] = coef * sum_w g[w-p] (M[w] + 2 X[p] S[w] + Y[p] T[w])
//                          (the transpose of the three filters), plus the pointwise mse / l1 terms,
//                          times the clamp mask (0 <= out <= 1); written as v_out f32[H,W,3].
#include "gi2d_common.cuh"

namespace gi2d {
namespace {

constexpr int kWin = 11;                // pytorch_msssim's default window; 5 for `Fusion_hinerv` (win_size=5)
constexpr int kLT = 16;                 // output tile edge
template <int kTaps>
struct Geo {
    static constexpr int kHalo = kTaps / 2;
    static constexpr int kLIn = kLT + 2 * kHalo;   // staged edge: 26 (11 taps), 20 (5 taps)
    static constexpr int kLPad = kLIn + 1;
};

// float32 values of torch: g = exp(-(arange(11)-5)^2 / (2*1.5^2)); g /= g.sum()
__constant__ float c_win[kWin] = {
    0.0010283803567290306f, 0.0075987582094967365f, 0.036000773310661316f, 0.10936068743467331f,
    0.21300552785396576f,   0.26601171493530273f,   0.21300552785396576f,  0.10936068743467331f,
    0.036000773310661316f,  0.0075987582094967365f, 0.0010283803567290306f};
// the same for win_size = 5 (sigma stays 1.5): exp(-(arange(5)-2)^2 / 4.5) normalised in float32
__constant__ float c_win5[5] = {0.12007837742567062f, 0.23388074338436127f, 0.29208171367645264f,
                                0.23388074338436127f, 0.12007837742567062f};
template <int kTaps>
__device__ __forceinline__ float win_tap(int k) { return kTaps == 11 ? c_win[k] : c_win5[k]; }

constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

__device__ __forceinline__ float target_at(const float *gt, const uint8_t *gt_u8, size_t idx) {
    return gt ? __ldg(gt + idx) : u8_to_unit(__ldg(gt_u8 + idx));
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// S1.  dm: 9 planes f32[H*W], plane 3*c+q for channel c and q in {M,S,T}.
// kEval: no derivative planes.  chan_sums (optional): the per-channel sums of the SSIM map and of the
// contrast-structure map (`cs` of pytorch_msssim `_ssim`) go to chan_sums[2*c], chan_sums[2*c+1] -- one level of
// MS-SSIM.  cs_only: the derivative planes are those of the cs map (levels 0..3 of MS-SSIM), not of the SSIM map.
// clamp_x: the first image is an unclamped render (level 0); pooled levels are inside [0,1] already.
template <bool kEval, int kTaps>
__global__ void __launch_bounds__(256)
ssim_stats_kernel(int H, int W, const float *__restrict__ render, const float *__restrict__ gt,
                  const uint8_t *__restrict__ gt_u8, float *__restrict__ dm, double *__restrict__ ssim_sum,
                  double *__restrict__ chan_sums, int cs_only, int clamp_x) {
    constexpr int kHalo = Geo<kTaps>::kHalo, kLIn = Geo<kTaps>::kLIn, kLPad = Geo<kTaps>::kLPad;
    __shared__ float sx[kLIn][kLPad], sy[kLIn][kLPad];
    __shared__ float sh[5][kLIn][kLT];
    __shared__ float s_red[8];
    const int tid = threadIdx.x, lx = tid & 15, ly = tid >> 4;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const int px = x0 + lx, py = y0 + ly;
    const size_t plane = (size_t)H * W;
    // a window is valid when all of its pixels are inside the image
    const bool valid = px >= kHalo && px < W - kHalo && py >= kHalo && py < H - kHalo;
    float acc = 0.f;
    for (int c = 0; c < 3; ++c) {
        for (int i = tid; i < kLIn * kLIn; i += 256) {
            const int r = i / kLIn, q = i - r * kLIn;
            const int yy = y0 - kHalo + r, xx = x0 - kHalo + q;
            float a = 0.f, b = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t idx = 3 * ((size_t)yy * W + xx) + c;
                a = __ldg(render + idx);
                if (clamp_x) a = clamp01(a);
                b = target_at(gt, gt_u8, idx);
            }
            sx[r][q] = a;
            sy[r][q] = b;
        }
        __syncthreads();
        // horizontal pass: kLIn rows x 16 columns x 5 moments
        for (int i = tid; i < kLIn * kLT; i += 256) {
            const int r = i >> 4, q = i & 15;
            float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
            for (int k = 0; k < kTaps; ++k) {
                const float w = win_tap<kTaps>(k), a = sx[r][q + k], b = sy[r][q + k];
                m1 = fmaf(w, a, m1);
                m2 = fmaf(w, b, m2);
                xx = fmaf(w, a * a, xx);
                yy = fmaf(w, b * b, yy);
                xy = fmaf(w, a * b, xy);
            }
            sh[0][r][q] = m1; sh[1][r][q] = m2; sh[2][r][q] = xx; sh[3][r][q] = yy; sh[4][r][q] = xy;
        }
        __syncthreads();
        float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const float w = win_tap<kTaps>(k);
            m1 = fmaf(w, sh[0][ly + k][lx], m1);
            m2 = fmaf(w, sh[1][ly + k][lx], m2);
            xx = fmaf(w, sh[2][ly + k][lx], xx);
            yy = fmaf(w, sh[3][ly + k][lx], yy);
            xy = fmaf(w, sh[4][ly + k][lx], xy);
        }
        float M = 0.f, S = 0.f, T = 0.f;
        float ev_ssim = 0.f, ev_cs = 0.f;
        if (valid) {
            const float s1 = xx - m1 * m1, s2 = yy - m2 * m2, s12 = xy - m1 * m2;
            const float A1 = fmaf(2.f * m1, m2, kC1), A2 = fmaf(2.f, s12, kC2);
            const float B1 = fmaf(m1, m1, fmaf(m2, m2, kC1)), B2 = s1 + s2 + kC2;
            const float iB1 = 1.f / B1, iB2 = 1.f / B2;
            const float lum = A1 * iB1, cs = A2 * iB2;
            acc += lum * cs;
            ev_ssim = lum * cs;
            ev_cs = cs;
            // partials at fixed (s1, s12), then the chain through s1 = E[x^2] - mu1^2, s12 = E[xy] - mu1 mu2
            if (cs_only) {
                S = -cs * iB2;
                T = 2.f * iB2;
                M = -2.f * m1 * S - m2 * T;
            } else {
                S = -lum * cs * iB2;
                T = 2.f * lum * iB2;
                const float dmu = 2.f * (m2 - m1 * lum) * iB1 * cs;
                M = dmu - 2.f * m1 * S - m2 * T;
            }
        }
        if (!kEval && px < W && py < H) {
            const size_t pix = (size_t)py * W + px;
            dm[(3 * c + 0) * plane + pix] = M;
            dm[(3 * c + 1) * plane + pix] = S;
            dm[(3 * c + 2) * plane + pix] = T;
        }
        __syncthreads();  // sx/sy/sh are rewritten by the next channel
        if (chan_sums) {   // (kernel-uniform)
#pragma unroll
            for (int q = 0; q < 2; ++q) {   // this channel's sums of the SSIM map and of the cs map
                float v = q ? ev_cs : ev_ssim;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if ((tid & 31) == 0) s_red[tid >> 5] = v;
                __syncthreads();
                if (tid == 0) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) t += s_red[w];
                    atomicAdd(chan_sums + 2 * c + q, (double)t);
                }
                __syncthreads();
            }
        }
    }
    if (ssim_sum) {   // (kernel-uniform)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if ((tid & 31) == 0) s_red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += s_red[w];
            atomicAdd(ssim_sum, (double)t);
        }
    }
}

// F.avg_pool2d(x, kernel_size=2, padding=(H%2, W%2)) of pytorch_msssim.ms_ssim on an HWC image: stride 2, the
// zero padding counts in the average.  `clamp`/`u8`: level 0 reads the unclamped render / the 8-bit target.
__global__ void __launch_bounds__(256)
avgpool2_kernel(int H, int W, const float *__restrict__ in, const uint8_t *__restrict__ in_u8, int clamp, int H2,
                int W2, float *__restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= H2 * W2 * 3) return;
    const int c = i % 3, x = (i / 3) % W2, y = i / (3 * W2);
    const int y0 = 2 * y - (H & 1), x0 = 2 * x - (W & 1);
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int yy = y0 + dy, xx = x0 + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t idx = 3 * ((size_t)yy * W + xx) + c;
                float v = in ? __ldg(in + idx) : u8_to_unit(__ldg(in_u8 + idx));
                if (clamp) v = clamp01(v);
                s += v;
            }
        }
    out[i] = 0.25f * s;
}

// S2.  v_out f32[H,W,3] = mask * (coef * filter^T(M,S,T) + up(deeper) + l2_scale * d + l1_scale * sign(d)),
// d = clamp(out) - gt.  coef: `ssim_coef` for every channel, or (chan_coef != nullptr) one per channel read from the
// device (MS-SSIM: they depend on all five levels).  deeper (optional, f32[H2,W2,3]): the gradient with respect
// to the next pyramid level, pulled back through avg_pool2d (every pixel belongs to exactly one 2x2 window:
// 0.25 * deeper[(y + H%2) / 2][(x + W%2) / 2]).  level0: x = clamp(render) with torch.clamp's backward mask; the
// pooled levels are plain images.
template <int kTaps>
__global__ void __launch_bounds__(256)
ssim_grad_kernel(int H, int W, const float *__restrict__ render, const float *__restrict__ gt,
                 const uint8_t *__restrict__ gt_u8, const float *__restrict__ dm, float ssim_coef,
                 const float *__restrict__ chan_coef, float l2_scale, float l1_scale,
                 const float *__restrict__ deeper, int H2, int W2, int level0, float *__restrict__ v_out) {
    constexpr int kHalo = Geo<kTaps>::kHalo, kLIn = Geo<kTaps>::kLIn, kLPad = Geo<kTaps>::kLPad;
    __shared__ float sd[3][kLIn][kLPad];
    __shared__ float sh[3][kLIn][kLT];
    const int tid = threadIdx.x, lx = tid & 15, ly = tid >> 4;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
    const int px = x0 + lx, py = y0 + ly;
    const size_t plane = (size_t)H * W;
    const bool inside = px < W && py < H;
    for (int c = 0; c < 3; ++c) {
        for (int i = tid; i < kLIn * kLIn; i += 256) {
            const int r = i / kLIn, q = i - r * kLIn;
            const int yy = y0 - kHalo + r, xx = x0 - kHalo + q;
            const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const size_t pix = ok ? (size_t)yy * W + xx : 0;
#pragma unroll
            for (int m = 0; m < 3; ++m) sd[m][r][q] = ok ? __ldg(dm + (3 * c + m) * plane + pix) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kLIn * kLT; i += 256) {
            const int r = i >> 4, q = i & 15;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int k = 0; k < kTaps; ++k) {
                const float w = win_tap<kTaps>(k);
                a0 = fmaf(w, sd[0][r][q + k], a0);
                a1 = fmaf(w, sd[1][r][q + k], a1);
                a2 = fmaf(w, sd[2][r][q + k], a2);
            }
            sh[0][r][q] = a0; sh[1][r][q] = a1; sh[2][r][q] = a2;
        }
        __syncthreads();
        float fM = 0.f, fS = 0.f, fT = 0.f;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const float w = win_tap<kTaps>(k);
            fM = fmaf(w, sh[0][ly + k][lx], fM);
            fS = fmaf(w, sh[1][ly + k][lx], fS);
            fT = fmaf(w, sh[2][ly + k][lx], fT);
        }
        if (inside) {
            const size_t idx = 3 * ((size_t)py * W + px) + c;
            const float o = __ldg(render + idx);
            const float x = level0 ? clamp01(o) : o, y = target_at(gt, gt_u8, idx);
            const float d = x - y;
            const float coef = chan_coef ? __ldcg(chan_coef + c) : ssim_coef;
            float v = coef * fmaf(2.f * x, fS, fmaf(y, fT, fM));
            if (deeper) {
                const int yy = (py + (H & 1)) >> 1, xx = (px + (W & 1)) >> 1;
                v = fmaf(0.25f, __ldcg(deeper + 3 * ((size_t)yy * W2 + xx) + c), v);
            }
            v = fmaf(l2_scale, d, v);
            v = fmaf(l1_scale, (float)((d > 0.f) - (d < 0.f)), v);
            v_out[idx] = (!level0 || (o >= 0.f && o <= 1.f)) ? v : 0.f;  // torch.clamp backward
        }
        __syncthreads();
    }
}

// MS-SSIM (pytorch_msssim.ms_ssim): per channel c the product over the 5 levels of relu(mean cs_l)^w_l (levels 0..3)
// and relu(mean ssim_4)^w_4; the result is the mean over the channels.  level_sums: [5][6] = per level and channel
// (sum ssim map, sum cs map); counts[l] = windows per channel of level l.  Writes the value and, for the loss term
// weight * (1 - ms_ssim), d loss / d (map sum of level l, channel c) into coef[3*l + c].
struct LevelCounts { float v[5]; };
__global__ void ms_coef_kernel(const double *__restrict__ level_sums, LevelCounts cnt, float weight,
                               float *__restrict__ coef, double *__restrict__ value_out) {
    const float *counts = cnt.v;
    const double w[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
    if (threadIdx.x != 0) return;
    double ms = 0.0;
    for (int c = 0; c < 3; ++c) {
        double v[5], prod = 1.0;
        for (int l = 0; l < 5; ++l) {
            const double mean = level_sums[6 * l + 2 * c + (l < 4 ? 1 : 0)] / (double)counts[l];
            v[l] = mean > 0.0 ? mean : 0.0;
            prod *= pow(v[l], w[l]);
        }
        ms += prod / 3.0;
        for (int l = 0; l < 5; ++l)
            coef[3 * l + c] = v[l] > 0.0 ? (float)(-(double)weight / 3.0 * w[l] * prod / v[l] / (double)counts[l]) : 0.f;
    }
    if (value_out) *value_out = ms;
}

template <int kTaps>
void launch_stats(bool eval, int H, int W, const float *x, const float *y, const uint8_t *y8, float *dm,
                  double *ssim_sum, double *chan_sums, int cs_only, int clamp_x, cudaStream_t st) {
    const dim3 grid(cdiv(W, kLT), cdiv(H, kLT));
    if (eval) ssim_stats_kernel<true, kTaps><<<grid, 256, 0, st>>>(H, W, x, y, y8, dm, ssim_sum, chan_sums, cs_only, clamp_x);
    else ssim_stats_kernel<false, kTaps><<<grid, 256, 0, st>>>(H, W, x, y, y8, dm, ssim_sum, chan_sums, cs_only, clamp_x);
}

inline int pooled(int s) { return (s + 2 * (s & 1) - 2) / 2 + 1; }

struct Pyramid {
    int h[5], w[5];
    size_t px[5];
    size_t levels_px;   // sum over levels 1..4
};

inline Pyramid pyramid(int H, int W) {
    Pyramid p;
    p.h[0] = H; p.w[0] = W; p.px[0] = (size_t)H * W;
    p.levels_px = 0;
    for (int l = 1; l < 5; ++l) {
        p.h[l] = pooled(p.h[l - 1]);
        p.w[l] = pooled(p.w[l - 1]);
        p.px[l] = (size_t)p.h[l] * p.w[l];
        p.levels_px += p.px[l];
    }
    return p;
}

}  // namespace

// Internal launcher shared with the fused fit step.  dm_ws: 9 * H * W floats.
int ssim_grad_launch(int H, int W, const float *render, const float *gt, const uint8_t *gt_u8, float *dm_ws,
                     float ssim_weight, float l2_scale, float l1_scale, float *v_out, double *ssim_sum,
                     cudaStream_t st) {
    constexpr int kHalo = kWin / 2;
    const dim3 grid(cdiv(W, kLT), cdiv(H, kLT));
    launch_stats<kWin>(false, H, W, render, gt, gt_u8, dm_ws, ssim_sum, nullptr, 0, 1, st);
    // loss term = ssim_weight * (1 - mean(map)), mean over 3 channels x (H-10)(W-10) windows
    const float coef = -ssim_weight / (3.f * (float)(H - 2 * kHalo) * (float)(W - 2 * kHalo));
    ssim_grad_kernel<kWin><<<grid, 256, 0, st>>>(H, W, render, gt, gt_u8, dm_ws, coef, nullptr, l2_scale, l1_scale,
                                                 nullptr, 0, 0, 1, v_out);
    return GI2D_OK;
}

// Workspace of the MS-SSIM loss gradient, in floats: both pyramids (levels 1..4), derivative planes of all five
// levels, the gradient images of levels 1..4, and (as 64 floats) the level sums + coefficients + window counts.
size_t msssim_grad_workspace_floats(int H, int W) {
    const Pyramid p = pyramid(H, W);
    return p.levels_px * 6 + (p.px[0] + p.levels_px) * 9 + p.levels_px * 3 + 4 + 128;
}

// d [ weight * (1 - ms_ssim(clamp(render), gt)) + l1 term ] / d render -> v_out; the value of ms_ssim -> *value
// (device pointer, optional).  19 launches: 5 x (moments + derivative planes), 8 pools, the coefficients, 5
// transposed filters from the coarsest level up.
template <int kTaps>
int msssim_grad_launch_t(int H, int W, const float *render, const float *gt, const uint8_t *gt_u8, float *ws,
                         float weight, float l1_scale, float *v_out, double *value, cudaStream_t st) {
    const Pyramid p = pyramid(H, W);
    float *xs[5], *ys[5], *dm[5], *gl[5];
    float *c = ws;
    xs[0] = nullptr; ys[0] = nullptr; gl[0] = v_out;
    for (int l = 1; l < 5; ++l) { xs[l] = c; c += p.px[l] * 3; ys[l] = c; c += p.px[l] * 3; }
    for (int l = 0; l < 5; ++l) { dm[l] = c; c += p.px[l] * 9; }
    for (int l = 1; l < 5; ++l) { gl[l] = c; c += p.px[l] * 3; }
    c = ws + (((size_t)(c - ws) + 3) & ~(size_t)3);             // (the doubles below want 8-byte alignment)
    double *level_sums = reinterpret_cast<double *>(c);          // 30 doubles = 60 floats
    float *coef = c + 64;                                        // 15 floats
    cudaMemsetAsync(level_sums, 0, 32 * sizeof(double), st);
    LevelCounts cnt;   // windows per channel of every level
    for (int l = 0; l < 5; ++l) cnt.v[l] = (float)(p.h[l] - (kTaps - 1)) * (float)(p.w[l] - (kTaps - 1));
    for (int l = 0; l < 5; ++l) {
        const float *x = l ? xs[l] : render, *y = l ? ys[l] : gt;
        const uint8_t *y8 = (l == 0 && !gt) ? gt_u8 : nullptr;
        launch_stats<kTaps>(false, p.h[l], p.w[l], x, y, y8, dm[l], nullptr, level_sums + 6 * l, l < 4, l == 0, st);
        if (l == 4) break;
        const int n = (int)(p.px[l + 1] * 3);
        avgpool2_kernel<<<cdiv(n, 256), 256, 0, st>>>(p.h[l], p.w[l], x, nullptr, l == 0, p.h[l + 1], p.w[l + 1], xs[l + 1]);
        avgpool2_kernel<<<cdiv(n, 256), 256, 0, st>>>(p.h[l], p.w[l], y, y8, 0, p.h[l + 1], p.w[l + 1], ys[l + 1]);
    }
    ms_coef_kernel<<<1, 32, 0, st>>>(level_sums, cnt, weight, coef, value);
    for (int l = 4; l >= 0; --l) {
        const float *x = l ? xs[l] : render, *y = l ? ys[l] : gt;
        const uint8_t *y8 = (l == 0 && !gt) ? gt_u8 : nullptr;
        const dim3 grid(cdiv(p.w[l], kLT), cdiv(p.h[l], kLT));
        ssim_grad_kernel<kTaps><<<grid, 256, 0, st>>>(p.h[l], p.w[l], x, y, y8, dm[l], 0.f, coef + 3 * l, 0.f,
                                                      l == 0 ? l1_scale : 0.f, l < 4 ? gl[l + 1] : nullptr,
                                                      l < 4 ? p.h[l + 1] : 0, l < 4 ? p.w[l + 1] : 0, l == 0, gl[l]);
    }
    return GI2D_OK;
}

int msssim_grad_launch(int H, int W, int win, const float *render, const float *gt, const uint8_t *gt_u8, float *ws,
                       float weight, float l1_scale, float *v_out, double *value, cudaStream_t st) {
    return win == 5 ? msssim_grad_launch_t<5>(H, W, render, gt, gt_u8, ws, weight, l1_scale, v_out, value, st)
                    : msssim_grad_launch_t<11>(H, W, render, gt, gt_u8, ws, weight, l1_scale, v_out, value, st);
}

}  // namespace gi2d

using namespace gi2d;

extern "C" size_t gi2d_ssim_workspace_size(int img_height, int img_width) {
    return (size_t)9 * img_height * img_width * sizeof(float);
}

extern "C" int gi2d_image_loss_grad(int img_height, int img_width, const float *render_hwc, const float *gt_hwc,
                                    const uint8_t *gt_u8_hwc, float ssim_weight, float l2_scale, float l1_scale,
                                    float *v_out_hwc, double *ssim_sum, void *workspace, size_t workspace_bytes,
                                    gi2d_stream_t stream) {
    GI2D_REQUIRE(img_height >= kWin && img_width >= kWin, "SSIM needs an image of at least 11x11 pixels");
    GI2D_REQUIRE(render_hwc && (gt_hwc || gt_u8_hwc) && v_out_hwc && workspace, "null buffer");
    if (workspace_bytes < gi2d_ssim_workspace_size(img_height, img_width)) {
        set_error("gi2d_image_loss_grad: workspace too small");
        return GI2D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (ssim_sum) cudaMemsetAsync(ssim_sum, 0, sizeof(double), st);
    ssim_grad_launch(img_height, img_width, render_hwc, gt_hwc, gt_u8_hwc, (float *)workspace, ssim_weight, l2_scale,
                     l1_scale, v_out_hwc, ssim_sum, st);
    return check_launch(__func__);
}

// ---- MS-SSIM (evaluation metric of train.py:190 / train_quantize.py:214; pytorch_msssim.ms_ssim, 5 levels)
extern "C" size_t gi2d_ms_ssim_workspace_size(int img_height, int img_width) {
    return pyramid(img_height, img_width).levels_px * 3 * sizeof(float) * 2;   // both images, levels 1..4
}

extern "C" int gi2d_ms_ssim(int img_height, int img_width, const float *render_hwc, const float *gt_hwc,
                            const uint8_t *gt_u8_hwc, double *level_sums, void *workspace, size_t workspace_bytes,
                            gi2d_stream_t stream) {
    GI2D_REQUIRE(render_hwc && (gt_hwc || gt_u8_hwc) && level_sums && workspace, "null buffer");
    GI2D_REQUIRE((img_height < img_width ? img_height : img_width) > (kWin - 1) * 16,
                 "MS-SSIM needs the smaller image side to exceed 160 pixels (pytorch_msssim's own assertion)");
    if (workspace_bytes < gi2d_ms_ssim_workspace_size(img_height, img_width)) {
        set_error("gi2d_ms_ssim: workspace too small");
        return GI2D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(level_sums, 0, 5 * 6 * sizeof(double), st);
    float *ws = (float *)workspace;
    const float *x = render_hwc, *y = gt_hwc;
    const uint8_t *y8 = gt_hwc ? nullptr : gt_u8_hwc;
    int h = img_height, w = img_width;
    for (int l = 0; l < 5; ++l) {
        launch_stats<kWin>(true, h, w, x, y, y8, nullptr, nullptr, level_sums + 6 * l, 0, 1, st);
        if (l == 4) break;
        const int h2 = pooled(h), w2 = pooled(w);
        float *x2 = ws, *y2 = ws + (size_t)h2 * w2 * 3;
        ws += (size_t)h2 * w2 * 6;
        const int n = h2 * w2 * 3;
        avgpool2_kernel<<<cdiv(n, 256), 256, 0, st>>>(h, w, x, nullptr, l == 0, h2, w2, x2);
        avgpool2_kernel<<<cdiv(n, 256), 256, 0, st>>>(h, w, y, y8, 0, h2, w2, y2);
        x = x2; y = y2; y8 = nullptr; h = h2; w = w2;
    }
    return check_launch(__func__);
}

// ---- MS-SSIM as a TRAINING loss (`Fusion4`, `Fusion_hinerv` of models/utils.py:76-79):
//   loss = l1_weight * l1 + msssim_weight * (1 - ms_ssim(clamp(render), gt, win_size = win))
// v_out = d loss / d render (f32[H,W,3]), *ms_value (DEVICE pointer, optional) = ms_ssim.
extern "C" size_t gi2d_msssim_grad_workspace_size(int img_height, int img_width) {
    return msssim_grad_workspace_floats(img_height, img_width) * sizeof(float);
}

extern "C" int gi2d_image_msssim_loss_grad(int img_height, int img_width, int win, const float *render_hwc,
                                           const float *gt_hwc, const uint8_t *gt_u8_hwc, float msssim_weight,
                                           float l1_scale, float *v_out_hwc, double *ms_value, void *workspace,
                                           size_t workspace_bytes, gi2d_stream_t stream) {
    GI2D_REQUIRE(win == 11 || win == 5, "win_size must be 11 (ms_ssim's default) or 5 (Fusion_hinerv)");
    GI2D_REQUIRE(render_hwc && (gt_hwc || gt_u8_hwc) && v_out_hwc && workspace, "null buffer");
    GI2D_REQUIRE((img_height < img_width ? img_height : img_width) > (win - 1) * 16,
                 "MS-SSIM needs the smaller image side to exceed (win_size - 1) * 16 pixels (pytorch_msssim's assertion)");
    if (workspace_bytes < gi2d_msssim_grad_workspace_size(img_height, img_width)) {
        set_error("gi2d_image_msssim_loss_grad: workspace too small");
        return GI2D_ERR_WORKSPACE;
    }
    msssim_grad_launch(img_height, img_width, win, render_hwc, gt_hwc, gt_u8_hwc, (float *)workspace, msssim_weight,
                       l1_scale, v_out_hwc, ms_value, (cudaStream_t)stream);
    return check_launch(__func__);
}
