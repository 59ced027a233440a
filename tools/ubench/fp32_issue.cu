// fp32_issue.cu -- issue-rate microbenchmarks behind the rasterizer's roofline (B200, sm_100a).
//   ffma     : 3-register FFMA, 16 independent chains, 512 FFMAs per loop trip  -> the FP32 peak
//   ffma2    : packed fma.rn.f32x2 (SASS FFMA2), same chains                    -> FLOP/s and issue slots per FLOP
//   mix_*    : FFMA(2) interleaved with integer ALU work (LOP3/IADD3)           -> do the pipes co-issue?
//   mufu     : ex2.approx throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int kChains = 16, kUnroll = 32;  // 512 FFMA per trip

__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float m, float c) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, float m, float c) {
    u64 a[kChains];
    const u64 mm = pack(m, m), cc = pack(c, c);
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = pack(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fma2(a[i], mm, cc);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { float x, y; unpack(a[i], x, y); s += x + y; }
    if (s == 12345.678f) out[0] = s;
}

// per trip: 256 FFMA + 512 integer ALU ops (LOP3 + SHF per chain step)
__global__ void __launch_bounds__(256) k_mix(float *out, int iters, float m, float c, unsigned k) {
    float a[kChains];
    unsigned b[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = threadIdx.x + i; b[i] = threadIdx.x * 17u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll / 2; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) { a[i] = fmaf(a[i], m, c); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(k)); asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(b[i])); }
    }
    float s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { s += a[i]; t ^= b[i]; }
    if (s == 12345.678f || t == 0xdeadbeefu) out[0] = s + t;
}

__global__ void __launch_bounds__(256) k_mix2(float *out, int iters, float m, float c, unsigned k) {
    u64 a[kChains];
    unsigned b[kChains];
    const u64 mm = pack(m, m), cc = pack(c, c);
#pragma unroll
    for (int i = 0; i < kChains; ++i) { a[i] = pack(threadIdx.x + i, threadIdx.x - i); b[i] = threadIdx.x * 17u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll / 2; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) { a[i] = fma2(a[i], mm, cc); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(k)); asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(b[i])); }
    }
    float s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { float x, y; unpack(a[i], x, y); s += x + y; t ^= b[i]; }
    if (s == 12345.678f || t == 0xdeadbeefu) out[0] = s + t;
}

__global__ void __launch_bounds__(256) k_mufu(float *out, int iters) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = -(float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float *d; cudaMalloc(&d, 4);
    const int iters = 2048;
    const double nominal = sms * 128.0 * 2.0 * clk * 1e3 / 1e12;
    for (int occ : {2, 4, 8}) {
        const int grid = sms * occ;
        const double thr = (double)grid * 256.0 * iters;
        float ms;
        ms = time_ms([&] { k_ffma<<<grid, 256>>>(d, iters, 1.0000001f, 1e-7f); });
        printf("{\"bench\":\"ffma\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"frac_nominal\":%.3f,\"warp_inst_per_clk_per_smsp\":%.3f}\n", occ,
               thr * 512 * 2 / (ms * 1e-3) / 1e12, thr * 512 * 2 / (ms * 1e-3) / 1e12 / nominal,
               thr / 32 * 512 / (ms * 1e-3) / (sms * 4.0 * clk * 1e3));
        ms = time_ms([&] { k_ffma2<<<grid, 256>>>(d, iters, 1.0000001f, 1e-7f); });
        printf("{\"bench\":\"ffma2\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"frac_nominal\":%.3f,\"warp_inst_per_clk_per_smsp\":%.3f}\n", occ,
               thr * 512 * 4 / (ms * 1e-3) / 1e12, thr * 512 * 4 / (ms * 1e-3) / 1e12 / nominal,
               thr / 32 * 512 / (ms * 1e-3) / (sms * 4.0 * clk * 1e3));
        ms = time_ms([&] { k_mix<<<grid, 256>>>(d, iters, 1.0000001f, 1e-7f, 0x1234u); });
        printf("{\"bench\":\"mix_ffma_alu\",\"ctas_per_sm\":%d,\"fp_tflops\":%.2f,\"warp_inst_per_clk_per_smsp\":%.3f}\n", occ,
               thr * 256 * 2 / (ms * 1e-3) / 1e12, thr / 32 * (256 + 256 * 2) / (ms * 1e-3) / (sms * 4.0 * clk * 1e3));
        ms = time_ms([&] { k_mix2<<<grid, 256>>>(d, iters, 1.0000001f, 1e-7f, 0x1234u); });
        printf("{\"bench\":\"mix_ffma2_alu\",\"ctas_per_sm\":%d,\"fp_tflops\":%.2f,\"warp_inst_per_clk_per_smsp\":%.3f}\n", occ,
               thr * 256 * 4 / (ms * 1e-3) / 1e12, thr / 32 * (256 + 256 * 2) / (ms * 1e-3) / (sms * 4.0 * clk * 1e3));
        ms = time_ms([&] { k_mufu<<<grid, 256>>>(d, iters); });
        printf("{\"bench\":\"mufu_ex2\",\"ctas_per_sm\":%d,\"gops\":%.1f,\"per_clk_per_sm\":%.2f}\n", occ,
               thr * 512 / (ms * 1e-3) / 1e9, thr * 512 / (ms * 1e-3) / (sms * (double)clk * 1e3));
    }
    printf("{\"sms\":%d,\"clock_khz\":%d,\"nominal_tflops\":%.2f}\n", sms, clk, nominal);
    return cudaDeviceSynchronize() != cudaSuccess;
}
