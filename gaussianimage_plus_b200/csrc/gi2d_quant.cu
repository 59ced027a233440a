// gi2d_quant.cu -- the quantisation-aware iteration of the compression pass as kernels.
//
// Reference: GaussianImage_Covariance.forward_quantize / train_iter_quantize (models/gaussianimage_covariance.py:
// 219-247, 384-410) with the default `lsq` quantisers of quantize.py --
//   positions  UniformQuantizer(bits=12, learned, 2 channels)   (LSQ+, quantize.py:39-156)
//   covariance HybirdQuant: variances (columns 0, 2) through LogQuantizer(learned=False) whose range is re-derived
//              from the data ON EVERY CALL (min / max of log(|x|+1e-6) over the whole tensor, :223-235, gradients
//              flowing through min and max), covariance (column 1) through UniformQuantizer(1 channel)  (:336-389)
//   colours    UniformQuantizer(bits=6, learned, 3 channels)
// and four torch.optim.Adam instances (attributes; one per quantiser) each with a StepLR.
//
// In the reference that is ~60 elementwise autograd nodes + ~45 optimiser launches per iteration around the
// rasterizer; round 1 replayed the equivalent ~40 torch launches from a graph (0.54 ms per iteration against
// 0.03 ms for the fit step they surround).  Here it is four kernels around the fused fit step:
//   quant_range_kernel   min / max (+ tie counts) of log(|var|+1e-6) over the 2N variance elements   [1 CTA]
//   quant_forward_kernel per Gaussian: LSQ / log quantise-dequantise the 8 attributes -> the fit step's inputs
//   (gi2d_fit_forward_backward + gi2d_fit_input_grads: projection, binning, rasterize fwd + loss + bwd, -X G X)
//   quant_reduce_kernel  the 14 sums the quantiser parameters' gradients and the log range's gradient need
//   quant_update_kernel  per Gaussian: straight-through gradients of the 8 raw attributes (log quantiser: through
//                        exp / log and, for the elements that ARE the min / max, through the range) + Adam;
//                        block 0 also steps the 12 quantiser parameters and the iteration counter
// all asynchronous on one stream, no host round trip: one CUDA-graph replay per iteration.
// Arithmetic follows torch's float32 ops (IEEE divide, round-half-even, logf / expf); the sums are accumulated in
// double (torch sums in float with its own tree order: the parity tests compare to 1e-5 relative).
#include "gi2d_common.cuh"

namespace gi2d {
namespace {

// f64 slots of b->qstats
constexpr int kQStep = 0, kQMin = 1, kQMax = 2, kQNMin = 3, kQNMax = 4;
constexpr int kQSumLogScale = 5, kQSumLogBeta = 6;    // sum gq (r - m code), sum gq (1 - m)
constexpr int kQSumS = 7, kQSumB = 13;                // 6 LSQ channels each: xyz 0,1 | cov 2 | rgb 3,4,5
constexpr int kQTicket = 19;

struct Lsq {
    float y, code, r;
    bool inside;
};

// UniformQuantizer.forward (quantize.py:118-133): code = (x - beta) / s clamped to [qmin, qm], rounded (STE)
__device__ __forceinline__ Lsq lsq(float x, float s, float beta, float qmax) {
    Lsq o;
    o.code = __fdiv_rn(__fsub_rn(x, beta), s);
    o.inside = o.code >= 0.f && o.code <= qmax;
    o.r = rintf(fminf(fmaxf(o.code, 0.f), qmax));
    o.y = __fadd_rn(__fmul_rn(o.r, s), beta);   // torch: r * scale + beta, two roundings
    return o;
}

__device__ __forceinline__ float sigmoid_(float v) { return 1.f / (1.f + expf(-v)); }

// ---- log range: ONE CTA (the reduction is over 2N floats: 40 KB at N = 5000)
__global__ void __launch_bounds__(1024)
quant_range_kernel(gi2d_quant_params p, const float *__restrict__ cov, const float *__restrict__ bound,
                   double *__restrict__ qstats) {
    __shared__ float s_min[32], s_max[32];
    __shared__ int s_cnt[2][32];
    __shared__ float s_mm[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mn = __int_as_float(0x7f800000), mx = -mn;
    for (int i = tid; i < 2 * p.num_points; i += 1024) {
        const int g = i >> 1, k = (i & 1) * 2;
        const float L = logf(fabsf(__fadd_rn(cov[3 * g + k], bound[3 * g + k])) + 1e-6f);
        mn = fminf(mn, L);
        mx = fmaxf(mx, L);
    }
    for (int d = 16; d >= 1; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = s_min[lane]; mx = s_max[lane];
        for (int d = 16; d >= 1; d >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if (lane == 0) { s_mm[0] = mn; s_mm[1] = mx; }
    }
    __syncthreads();
    mn = s_mm[0]; mx = s_mm[1];
    // how many elements ARE the min / the max: torch.min() / torch.max() spread their gradient over ties
    int cmin = 0, cmax = 0;
    for (int i = tid; i < 2 * p.num_points; i += 1024) {
        const int g = i >> 1, k = (i & 1) * 2;
        const float L = logf(fabsf(__fadd_rn(cov[3 * g + k], bound[3 * g + k])) + 1e-6f);
        cmin += L == mn;
        cmax += L == mx;
    }
    for (int d = 16; d >= 1; d >>= 1) {
        cmin += __shfl_xor_sync(0xffffffffu, cmin, d);
        cmax += __shfl_xor_sync(0xffffffffu, cmax, d);
    }
    if (lane == 0) { s_cnt[0][warp] = cmin; s_cnt[1][warp] = cmax; }
    __syncthreads();
    if (tid == 0) {
        int a = 0, b2 = 0;
        for (int w = 0; w < 32; ++w) { a += s_cnt[0][w]; b2 += s_cnt[1][w]; }
        qstats[kQMin] = (double)mn;
        qstats[kQMax] = (double)mx;
        qstats[kQNMin] = (double)a;
        qstats[kQNMax] = (double)b2;
        for (int k = kQSumLogScale; k < kQTicket; ++k) qstats[k] = 0.0;   // the backward's accumulators
    }
}

struct LogQ {
    float L, code, r, y, scale, beta, mx;
    bool inside;
};

// LogQuantizer.forward, learned=False (quantize.py:223-235)
__device__ __forceinline__ LogQ logq(float x, float beta, float mx, float qmax) {
    LogQ o;
    o.beta = beta;
    o.mx = mx;
    o.scale = __fmul_rn(__fsub_rn(mx, beta), __fdiv_rn(1.f, qmax));   // (torch divides by a host scalar as x * (1 / s))
    o.L = logf(fabsf(x) + 1e-6f);
    o.code = __fdiv_rn(__fsub_rn(o.L, beta), o.scale);
    o.inside = o.code >= 0.f && o.code <= qmax;
    o.r = rintf(fminf(fmaxf(o.code, 0.f), qmax));
    o.y = expf(__fadd_rn(__fmul_rn(o.r, o.scale), beta));
    return o;
}

__global__ void __launch_bounds__(256)
quant_forward_kernel(gi2d_quant_params p, gi2d_quant_buffers b) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= p.num_points) return;
    const float *q = b.qparams;
    const float beta = (float)b.qstats[kQMin], mx = (float)b.qstats[kQMax];
    const Lsq x0 = lsq(b.xyz[2 * g], q[0], q[2], (float)p.xy_qmax);
    const Lsq x1 = lsq(b.xyz[2 * g + 1], q[1], q[3], (float)p.xy_qmax);
    b.out_xyz[2 * g] = x0.y;
    b.out_xyz[2 * g + 1] = x1.y;
    const float e0 = __fadd_rn(b.cov[3 * g], b.bound[3 * g]), e1 = __fadd_rn(b.cov[3 * g + 1], b.bound[3 * g + 1]);
    const float e2 = __fadd_rn(b.cov[3 * g + 2], b.bound[3 * g + 2]);
    b.out_cov[3 * g] = logq(e0, beta, mx, (float)p.cov_qmax).y;
    b.out_cov[3 * g + 1] = lsq(e1, q[4], q[5], (float)p.cov_qmax).y;
    b.out_cov[3 * g + 2] = logq(e2, beta, mx, (float)p.cov_qmax).y;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float f = b.rgb[3 * g + k];
        if (p.color_sigmoid) f = sigmoid_(f);
        b.out_rgb[3 * g + k] = lsq(f, q[6 + k], q[9 + k], (float)p.color_qmax).y;
    }
}

__device__ __forceinline__ double block_sum(double v, double *s_red) {
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 8) t = s_red[threadIdx.x];
    for (int d = 4; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    return t;   // valid in thread 0
}

// the 14 global sums of the backward (torch: `.sum(dim=0)` of the LSQ functions, `.sum()` of the log function)
__global__ void __launch_bounds__(256)
quant_reduce_kernel(gi2d_quant_params p, gi2d_quant_buffers b) {
    __shared__ double s_red[8];
    const int g = blockIdx.x * 256 + threadIdx.x;
    const bool mine = g < p.num_points;
    const float *q = b.qparams;
    const float beta = (float)b.qstats[kQMin], mx = (float)b.qstats[kQMax];
    double v[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) v[k] = 0.0;
    if (mine) {
        const float4 g0 = reinterpret_cast<const float4 *>(b.in_grads)[2 * g];
        const float4 g1 = reinterpret_cast<const float4 *>(b.in_grads)[2 * g + 1];
        const float gin[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        auto lsq_sums = [&](int ch, float x, float s, float bt, float qmax, float gy) {
            const Lsq o = lsq(x, s, bt, qmax);
            const float m = o.inside ? 1.f : 0.f;
            v[2 + ch] = (double)(gy * (o.r - m * o.code));   // d y / d s
            v[8 + ch] = (double)(gy * (1.f - m));            // d y / d beta
        };
        lsq_sums(0, b.xyz[2 * g], q[0], q[2], (float)p.xy_qmax, gin[0]);
        lsq_sums(1, b.xyz[2 * g + 1], q[1], q[3], (float)p.xy_qmax, gin[1]);
        lsq_sums(2, __fadd_rn(b.cov[3 * g + 1], b.bound[3 * g + 1]), q[4], q[5], (float)p.cov_qmax, gin[3]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float f = b.rgb[3 * g + k];
            if (p.color_sigmoid) f = sigmoid_(f);
            lsq_sums(3 + k, f, q[6 + k], q[9 + k], (float)p.color_qmax, gin[5 + k]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const LogQ o = logq(__fadd_rn(b.cov[3 * g + 2 * k], b.bound[3 * g + 2 * k]), beta, mx, (float)p.cov_qmax);
            const float gq = gin[2 + 2 * k] * o.y;           // through exp
            const float m = o.inside ? 1.f : 0.f;
            v[0] += (double)(gq * (o.r - m * o.code));
            v[1] += (double)(gq * (1.f - m));
        }
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        const double t = block_sum(v[k], s_red);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(b.qstats + kQSumLogScale + k, t);
    }
}

// torch.optim.Adam (_single_tensor_adam), float tensors, double scalars; returns the new parameter
__device__ __forceinline__ float adam1(float param, float &m, float &v, float grad, float beta1, float beta2,
                                       float step_size, float bc2_sqrt, float eps) {
    m = m + (grad - m) * (float)(1.0 - (double)beta1);
    v = v * beta2 + (float)(1.0 - (double)beta2) * grad * grad;
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
    return param - step_size * __fdiv_rn(m, denom);
}

__global__ void __launch_bounds__(256)
quant_update_kernel(gi2d_quant_params p, gi2d_quant_buffers b) {
    __shared__ float s_sc[4];   // step sizes (attributes, quantisers), sqrt(1 - beta2^t)
    const int g = blockIdx.x * 256 + threadIdx.x;
    double *qs = b.qstats;
    if (threadIdx.x == 0) {
        const double t = qs[kQStep] + 1.0;   // this iteration's step number (1-based)
        const double bc1 = 1.0 - pow((double)p.beta1, t), bc2 = 1.0 - pow((double)p.beta2, t);
        const double lr = (double)p.lr0 * pow((double)p.lr_gamma, p.lr_step > 0 ? floor((t - 1.0) / p.lr_step) : 0.0);
        const double lrq = (double)p.lr_q0 * pow((double)p.lr_gamma, p.lr_q_step > 0 ? floor((t - 1.0) / p.lr_q_step) : 0.0);
        s_sc[0] = (float)(lr / bc1);
        s_sc[1] = (float)(lrq / bc1);
        s_sc[2] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_a = s_sc[0], step_q = s_sc[1], bc2s = s_sc[2];
    const float *q = b.qparams;
    const float beta = (float)qs[kQMin], mx = (float)qs[kQMax];
    const float span = (float)p.cov_qmax;
    // gradients of the log range (quantize.py:225-229 through _LogQuantFn of the torch mirror)
    const float g_scale = (float)qs[kQSumLogScale];
    const float g_beta = (float)qs[kQSumLogBeta] - g_scale / span;
    const float g_max = g_scale / span;
    const float n_min = (float)qs[kQNMin], n_max = (float)qs[kQNMax];
    if (g < p.num_points) {
        const float4 g0 = reinterpret_cast<const float4 *>(b.in_grads)[2 * g];
        const float4 g1 = reinterpret_cast<const float4 *>(b.in_grads)[2 * g + 1];
        const float gin[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        // positions
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float x = b.xyz[2 * g + k];
            const Lsq o = lsq(x, q[k], q[2 + k], (float)p.xy_qmax);
            const float gx = o.inside ? gin[k] : 0.f;
            if (b.dbg_grads) b.dbg_grads[8 * g + k] = gx;
            b.xyz[2 * g + k] = adam1(x, b.m_xyz[2 * g + k], b.v_xyz[2 * g + k], gx, p.beta1, p.beta2, step_a, bc2s, p.eps);
        }
        // covariance parameters (the bound is a constant: d(cov + bound)/d cov = 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float c = b.cov[3 * g + k], e = __fadd_rn(c, b.bound[3 * g + k]);
            float gx;
            if (k == 1) {
                const Lsq o = lsq(e, q[4], q[5], (float)p.cov_qmax);
                gx = o.inside ? gin[3] : 0.f;
            } else {
                const LogQ o = logq(e, beta, mx, (float)p.cov_qmax);
                const float gq = gin[2 + k] * o.y;
                float gL = o.inside ? gq : 0.f;
                if (o.L == beta) gL += g_beta / n_min;
                if (o.L == mx) gL += g_max / n_max;
                const float sgn = (e > 0.f) - (e < 0.f);
                gx = gL * sgn / (fabsf(e) + 1e-6f);
            }
            if (b.dbg_grads) b.dbg_grads[8 * g + 2 + k] = gx;
            b.cov[3 * g + k] = adam1(c, b.m_cov[3 * g + k], b.v_cov[3 * g + k], gx, p.beta1, p.beta2, step_a, bc2s, p.eps);
        }
        // colours
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float raw = b.rgb[3 * g + k];
            float f = raw, df = 1.f;
            if (p.color_sigmoid) { f = sigmoid_(raw); df = f * (1.f - f); }
            const Lsq o = lsq(f, q[6 + k], q[9 + k], (float)p.color_qmax);
            const float gx = o.inside ? gin[5 + k] * df : 0.f;
            if (b.dbg_grads) b.dbg_grads[8 * g + 5 + k] = gx;
            b.rgb[3 * g + k] = adam1(raw, b.m_rgb[3 * g + k], b.v_rgb[3 * g + k], gx, p.beta1, p.beta2, step_a, bc2s, p.eps);
        }
    }
    // the last block to finish steps the 12 quantiser parameters (everybody has read them) and the counter
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd((unsigned *)(qs + kQTicket), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        // layout of qparams: s_xyz[2] b_xyz[2] s_cov b_cov s_rgb[3] b_rgb[3]; of the sums: scale sums then beta sums
        // per LSQ channel (xyz 0,1 | cov 2 | rgb 3,4,5)
        int ch;
        bool is_scale;
        if (k < 2) { ch = k; is_scale = true; }
        else if (k < 4) { ch = k - 2; is_scale = false; }
        else if (k == 4) { ch = 2; is_scale = true; }
        else if (k == 5) { ch = 2; is_scale = false; }
        else if (k < 9) { ch = 3 + (k - 6); is_scale = true; }
        else { ch = 3 + (k - 9); is_scale = false; }
        const float grad = (float)qs[(is_scale ? kQSumS : kQSumB) + ch];
        const float eps = k < 4 ? p.eps_xyz_q : p.eps;   // (the xyz quantiser's Adam keeps torch's default eps)
        b.qparams[k] = adam1(b.qparams[k], b.qm[k], b.qv[k], grad, p.beta1, p.beta2, step_q, bc2s, eps);
    }
    if (threadIdx.x == 0) {
        *(unsigned *)(qs + kQTicket) = 0u;
        qs[kQStep] = qs[kQStep] + 1.0;
    }
}

// UniformQuantizer._init_data / HybirdQuant._init_data (quantize.py:72-80, 352-354): scale and beta of the learned
// quantisers from the per-channel min / max of their first input.  ONE CTA.
__global__ void __launch_bounds__(1024)
quant_init_kernel(gi2d_quant_params p, gi2d_quant_buffers b) {
    __shared__ float s_min[6][32], s_max[6][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mn[6], mx[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { mn[c] = __int_as_float(0x7f800000); mx[c] = -mn[c]; }
    for (int g = tid; g < p.num_points; g += 1024) {
        float v[6] = {b.xyz[2 * g], b.xyz[2 * g + 1], __fadd_rn(b.cov[3 * g + 1], b.bound[3 * g + 1]),
                      b.rgb[3 * g], b.rgb[3 * g + 1], b.rgb[3 * g + 2]};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            if (c >= 3 && p.color_sigmoid) v[c] = sigmoid_(v[c]);
            mn[c] = fminf(mn[c], v[c]);
            mx[c] = fmaxf(mx[c], v[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        for (int d = 16; d >= 1; d >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
        }
        if (lane == 0) { s_min[c][warp] = mn[c]; s_max[c][warp] = mx[c]; }
    }
    __syncthreads();
    if (tid < 6) {
        float a = s_min[tid][0], z = s_max[tid][0];
        for (int w = 1; w < 32; ++w) { a = fminf(a, s_min[tid][w]); z = fmaxf(z, s_max[tid][w]); }
        const float qmax = tid < 2 ? (float)p.xy_qmax : (tid == 2 ? (float)p.cov_qmax : (float)p.color_qmax);
        // (t_max - t_min) / (qmax - qmin), qmin = 0; torch divides by a host scalar as x * (1 / s)
        const float scale = __fmul_rn(__fsub_rn(z, a), __fdiv_rn(1.f, qmax));
        const int si = tid < 2 ? tid : (tid == 2 ? 4 : 6 + (tid - 3));
        const int bi = tid < 2 ? 2 + tid : (tid == 2 ? 5 : 9 + (tid - 3));
        b.qparams[si] = scale;
        b.qparams[bi] = a;                                              // t_min - qmin * scale
    }
    if (tid < 12) { b.qm[tid] = 0.f; b.qv[tid] = 0.f; }
    if (tid == 0) {
        for (int k = 0; k < 32; ++k) b.qstats[k] = 0.0;
    }
}

int validate_q(const gi2d_quant_params *p, const gi2d_quant_buffers *b) {
    GI2D_REQUIRE(p && b, "null argument");
    GI2D_REQUIRE(p->num_points >= 0, "bad size");
    GI2D_REQUIRE(p->xy_qmax > 0 && p->cov_qmax > 0 && p->color_qmax > 0, "bad quantiser ranges");
    GI2D_REQUIRE(b->xyz && b->cov && b->rgb && b->bound && b->qparams && b->qm && b->qv && b->qstats, "null buffer");
    return GI2D_OK;
}

}  // namespace
}  // namespace gi2d

using namespace gi2d;

extern "C" int gi2d_quant_init(const gi2d_quant_params *p, const gi2d_quant_buffers *b, gi2d_stream_t stream) {
    const int rc = validate_q(p, b);
    if (rc != GI2D_OK) return rc;
    quant_init_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*p, *b);
    return check_launch(__func__);
}

extern "C" int gi2d_quant_forward(const gi2d_quant_params *p, const gi2d_quant_buffers *b, gi2d_stream_t stream) {
    const int rc = validate_q(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(b->out_xyz && b->out_cov && b->out_rgb, "null output");
    if (p->num_points == 0) return GI2D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    quant_range_kernel<<<1, 1024, 0, st>>>(*p, b->cov, b->bound, b->qstats);
    quant_forward_kernel<<<cdiv(p->num_points, 256), 256, 0, st>>>(*p, *b);
    return check_launch(__func__);
}

extern "C" int gi2d_quant_backward_step(const gi2d_quant_params *p, const gi2d_quant_buffers *b,
                                        gi2d_stream_t stream) {
    const int rc = validate_q(p, b);
    if (rc != GI2D_OK) return rc;
    GI2D_REQUIRE(b->in_grads && b->m_xyz && b->v_xyz && b->m_cov && b->v_cov && b->m_rgb && b->v_rgb, "null buffer");
    if (p->num_points == 0) return GI2D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = cdiv(p->num_points, 256);
    quant_reduce_kernel<<<grid, 256, 0, st>>>(*p, *b);
    quant_update_kernel<<<grid, 256, 0, st>>>(*p, *b);
    return check_launch(__func__);
}
