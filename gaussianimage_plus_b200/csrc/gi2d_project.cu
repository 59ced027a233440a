// gi2d_project.cu -- projection forward/backward kernels.
//
// One thread per Gaussian, 256 threads per CTA.  The [N,3] arrays (covariance parameters,
// conics, gradients) are AoS with a 12-byte stride; a CTA's slice of them is 3 KiB of
// contiguous memory, so it is moved with 128-bit vector loads/stores through shared memory
// (192 LDG.128 per CTA instead of 768 strided LDG.32) and read back per thread with a stride
// of 3 words, which is bank-conflict free.  These kernels are HBM-bound at large N
// (52 B/Gaussian forward, 68 B/Gaussian backward) and launch-bound at small N.
#include "gi2d_project_core.cuh"

namespace gi2d {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Cooperative copy of `count` floats global -> shared (vectorised when aligned).
__device__ __forceinline__ void block_load(const float *__restrict__ g, float *s, int count) {
    const int tid = threadIdx.x;
    if (aligned16(g)) {
        const int nvec = count >> 2;
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        float4 *s4 = reinterpret_cast<float4 *>(s);
        for (int i = tid; i < nvec; i += kThreads) s4[i] = __ldg(g4 + i);
        for (int i = (nvec << 2) + tid; i < count; i += kThreads) s[i] = __ldg(g + i);
    } else {
        for (int i = tid; i < count; i += kThreads) s[i] = __ldg(g + i);
    }
}

__device__ __forceinline__ void block_store(float *__restrict__ g, const float *s, int count) {
    const int tid = threadIdx.x;
    if (aligned16(g)) {
        const int nvec = count >> 2;
        float4 *g4 = reinterpret_cast<float4 *>(g);
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        for (int i = tid; i < nvec; i += kThreads) g4[i] = s4[i];
        for (int i = (nvec << 2) + tid; i < count; i += kThreads) g[i] = s[i];
    } else {
        for (int i = tid; i < count; i += kThreads) g[i] = s[i];
    }
}

enum class Param { Cov, Chol, ScaleRot };

template <Param P>
__global__ void __launch_bounds__(kThreads)
project_fwd_kernel(int n, const float *__restrict__ means2d, const float *__restrict__ p3,
                   const float *__restrict__ rot, int img_w, int img_h, int tiles_x, int tiles_y,
                   float clip_coe, float radius_clip, float *__restrict__ xys,
                   float *__restrict__ depths, int32_t *__restrict__ radii,
                   float *__restrict__ conics, int32_t *__restrict__ num_tiles_hit) {
    __shared__ __align__(16) float s3[kThreads * 3];
    const int base = blockIdx.x * kThreads;
    const int cnt = min(kThreads, n - base);
    const int tid = threadIdx.x;
    const int idx = base + tid;
    constexpr int kW = (P == Param::ScaleRot) ? 2 : 3;  // width of the second parameter array
    block_load(p3 + (size_t)base * kW, s3, cnt * kW);
    __syncthreads();
    Projected pr;
    if (tid < cnt) {
        const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + idx);
        if (P == Param::Cov) {
            pr = project_cov(m.x, m.y, s3[3 * tid], s3[3 * tid + 1], s3[3 * tid + 2], clip_coe,
                             radius_clip, tiles_x, tiles_y);
        } else if (P == Param::Chol) {
            pr = project_chol(m.x, m.y, s3[3 * tid], s3[3 * tid + 1], s3[3 * tid + 2], img_w, img_h,
                              clip_coe, radius_clip, tiles_x, tiles_y);
        } else {
            pr = project_rs(m.x, m.y, s3[2 * tid], s3[2 * tid + 1], __ldg(rot + idx), clip_coe,
                            radius_clip, tiles_x, tiles_y);
        }
    }
    __syncthreads();  // everyone is done reading s3 -> reuse it for the conics
    if (tid < cnt) {
        s3[3 * tid] = pr.a;
        s3[3 * tid + 1] = pr.b;
        s3[3 * tid + 2] = pr.c;
        reinterpret_cast<float2 *>(xys)[idx] = make_float2(pr.x, pr.y);
        depths[idx] = 0.f;
        radii[idx] = pr.radius;
        num_tiles_hit[idx] = pr.ntiles;
    }
    __syncthreads();
    block_store(conics + (size_t)base * 3, s3, cnt * 3);
}

__global__ void __launch_bounds__(kThreads)
cov2d_bounds_kernel(int n, float clip_coe, const float *__restrict__ cov2d,
                    float *__restrict__ conics, float *__restrict__ radii) {
    // reference: csrc/bindings.cu:21-39 (outputs stay zero when det == 0)
    const int idx = blockIdx.x * kThreads + threadIdx.x;
    if (idx >= n) return;
    const Cov2dBounds cb = cov2d_bounds(cov2d[3 * idx], cov2d[3 * idx + 1], cov2d[3 * idx + 2], clip_coe);
    conics[3 * idx] = cb.a;
    conics[3 * idx + 1] = cb.b;
    conics[3 * idx + 2] = cb.c;
    radii[idx] = cb.ok ? cb.rx : 0.f;
}

template <Param P>
__global__ void __launch_bounds__(kThreads)
project_bwd_kernel(int n, const float *__restrict__ p3, const float *__restrict__ rot, int img_w,
                   int img_h, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ v_xy, const float *__restrict__ v_conic,
                   float *__restrict__ v_cov2d, float *__restrict__ v_mean2d,
                   float *__restrict__ v_p3, float *__restrict__ v_rot) {
    __shared__ __align__(16) float s_conic[kThreads * 3];
    __shared__ __align__(16) float s_vconic[kThreads * 3];
    __shared__ __align__(16) float s_par[kThreads * 3];
    const int base = blockIdx.x * kThreads;
    const int cnt = min(kThreads, n - base);
    const int tid = threadIdx.x;
    const int idx = base + tid;
    constexpr int kW = (P == Param::ScaleRot) ? 2 : 3;
    block_load(conics + (size_t)base * 3, s_conic, cnt * 3);
    block_load(v_conic + (size_t)base * 3, s_vconic, cnt * 3);
    if (P != Param::Cov) block_load(p3 + (size_t)base * kW, s_par, cnt * kW);
    __syncthreads();
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;      // v_cov2d
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;      // v_L / v_scale
    float vr = 0.f, mx = 0.f, my = 0.f;
    if (tid < cnt && radii[idx] > 0) {
        conic_vjp(s_conic[3 * tid], s_conic[3 * tid + 1], s_conic[3 * tid + 2], s_vconic[3 * tid],
                  s_vconic[3 * tid + 1], s_vconic[3 * tid + 2], c0, c1, c2);
        const float2 vxy = __ldg(reinterpret_cast<const float2 *>(v_xy) + idx);
        if (P == Param::Cov) {
            // backward2d.cu:196-205 : the covariance entries are the parameters themselves
            q0 = c0; q1 = c1; q2 = c2;
            mx = vxy.x; my = vxy.y;
        } else if (P == Param::Chol) {
            // backward2d.cu:39-49 (G_12 is the already-summed off-diagonal: SURVEY Q5)
            const float l11 = s_par[3 * tid], l21 = s_par[3 * tid + 1], l22 = s_par[3 * tid + 2];
            q0 = 2.f * l11 * c0 + 2.f * c1 * l21;
            q1 = 2.f * l11 * c1 + 2.f * l21 * c2;
            q2 = 2.f * l22 * c2;
            mx = vxy.x * (0.5f * (float)(unsigned)img_w);
            my = vxy.y * (0.5f * (float)(unsigned)img_h);
        } else {
            // backward2d.cu:76-99 with the glm algebra expanded.  R = [[cs, sn],[-sn, cs]]
            // (row,col), S = diag(sx,sy), M = R S, Sigma = M M^T.
            const float sx = s_par[2 * tid], sy = s_par[2 * tid + 1];
            const float th = __ldg(rot + idx);
            const float cs = cosf(th), sn = sinf(th);
            // dSigma/dsx = 2 sx * r0 r0^T with r0 = first column of R = (cs, -sn)
            // dSigma/dsy = 2 sy * r1 r1^T with r1 = second column of R = (sn, cs)
            const float ax00 = 2.f * sx * cs * cs, ax01 = -2.f * sx * cs * sn, ax11 = 2.f * sx * sn * sn;
            const float ay00 = 2.f * sy * sn * sn, ay01 = 2.f * sy * sn * cs, ay11 = 2.f * sy * cs * cs;
            q0 = c0 * ax00 + 2.f * c1 * ax01 + c2 * ax11;
            q1 = c0 * ay00 + 2.f * c1 * ay01 + c2 * ay11;
            // dSigma/dtheta = R' S^2 R^T + R S^2 R'^T, R' = [[-sn, cs],[-cs, -sn]]
            const float s2x = sx * sx, s2y = sy * sy;
            const float t00 = 2.f * (s2y - s2x) * sn * cs;            // d(cs^2 s2x + sn^2 s2y)
            const float t01 = (s2y - s2x) * (cs * cs - sn * sn);      // d((s2y - s2x) sn cs)
            const float t11 = -t00;
            vr = c0 * t00 + 2.f * c1 * t01 + c2 * t11;
            mx = vxy.x; my = vxy.y;
        }
    }
    __syncthreads();
    if (tid < cnt) {
        s_conic[3 * tid] = c0; s_conic[3 * tid + 1] = c1; s_conic[3 * tid + 2] = c2;
        if (P == Param::ScaleRot) {
            reinterpret_cast<float2 *>(v_p3)[idx] = make_float2(q0, q1);
            v_rot[idx] = vr;
        } else {
            s_vconic[3 * tid] = q0; s_vconic[3 * tid + 1] = q1; s_vconic[3 * tid + 2] = q2;
        }
        reinterpret_cast<float2 *>(v_mean2d)[idx] = make_float2(mx, my);
    }
    __syncthreads();
    block_store(v_cov2d + (size_t)base * 3, s_conic, cnt * 3);
    if (P != Param::ScaleRot) block_store(v_p3 + (size_t)base * 3, s_vconic, cnt * 3);
}

}  // namespace
}  // namespace gi2d

using namespace gi2d;

#define PROJECT_FWD_BODY(PARAM, P3, ROT)                                                          \
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");                                             \
    GI2D_REQUIRE(num_points == 0 || (means2d && P3 && xys && depths && radii && conics &&        \
                                     num_tiles_hit),                                             \
                 "null pointer");                                                                \
    GI2D_REQUIRE(tiles_x >= 0 && tiles_y >= 0 && img_width >= 0 && img_height >= 0, "bad size"); \
    if (num_points == 0) return GI2D_OK;                                                         \
    project_fwd_kernel<PARAM><<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>( \
        num_points, means2d, P3, ROT, img_width, img_height, tiles_x, tiles_y, clip_coe,         \
        radius_clip, xys, depths, radii, conics, num_tiles_hit);                                 \
    return check_launch(__func__);

extern "C" int gi2d_project_cov_fwd(int num_points, const float *means2d, const float *cov2d,
                                    int img_width, int img_height, int tiles_x, int tiles_y,
                                    float clip_coe, float radius_clip, float *xys, float *depths,
                                    int32_t *radii, float *conics, int32_t *num_tiles_hit,
                                    gi2d_stream_t stream) {
    PROJECT_FWD_BODY(Param::Cov, cov2d, nullptr)
}

extern "C" int gi2d_project_chol_fwd(int num_points, const float *means2d, const float *L_elements,
                                     int img_width, int img_height, int tiles_x, int tiles_y,
                                     float clip_coe, float radius_clip, float *xys, float *depths,
                                     int32_t *radii, float *conics, int32_t *num_tiles_hit,
                                     gi2d_stream_t stream) {
    PROJECT_FWD_BODY(Param::Chol, L_elements, nullptr)
}

extern "C" int gi2d_project_rs_fwd(int num_points, const float *means2d, const float *scales2d,
                                   const float *rotation, int img_width, int img_height,
                                   int tiles_x, int tiles_y, float clip_coe, float radius_clip,
                                   float *xys, float *depths, int32_t *radii, float *conics,
                                   int32_t *num_tiles_hit, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points == 0 || rotation, "null rotation");
    PROJECT_FWD_BODY(Param::ScaleRot, scales2d, rotation)
}

extern "C" int gi2d_compute_cov2d_bounds(int num_points, float clip_coe, const float *cov2d,
                                         float *conics, float *radii_f32, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(cov2d && conics && radii_f32, "null pointer");
    cov2d_bounds_kernel<<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        num_points, clip_coe, cov2d, conics, radii_f32);
    return check_launch(__func__);
}

extern "C" int gi2d_project_cov_bwd(int num_points, const int32_t *radii, const float *conics,
                                    const float *v_xy, const float *v_conic, float *v_cov2d,
                                    float *v_mean2d, float *v_L, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(radii && conics && v_xy && v_conic && v_cov2d && v_mean2d && v_L, "null pointer");
    project_bwd_kernel<Param::Cov><<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        num_points, nullptr, nullptr, 0, 0, radii, conics, v_xy, v_conic, v_cov2d, v_mean2d, v_L,
        nullptr);
    return check_launch(__func__);
}

extern "C" int gi2d_project_chol_bwd(int num_points, const float *L_elements, int img_width,
                                     int img_height, const int32_t *radii, const float *conics,
                                     const float *v_xy, const float *v_conic, float *v_cov2d,
                                     float *v_mean2d, float *v_L, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(L_elements && radii && conics && v_xy && v_conic && v_cov2d && v_mean2d && v_L,
                 "null pointer");
    project_bwd_kernel<Param::Chol><<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        num_points, L_elements, nullptr, img_width, img_height, radii, conics, v_xy, v_conic,
        v_cov2d, v_mean2d, v_L, nullptr);
    return check_launch(__func__);
}

extern "C" int gi2d_project_rs_bwd(int num_points, const float *scales2d, const float *rotation,
                                   const int32_t *radii, const float *conics, const float *v_xy,
                                   const float *v_conic, float *v_cov2d, float *v_mean2d,
                                   float *v_scale, float *v_rot, gi2d_stream_t stream) {
    GI2D_REQUIRE(num_points >= 0, "num_points < 0");
    if (num_points == 0) return GI2D_OK;
    GI2D_REQUIRE(scales2d && rotation && radii && conics && v_xy && v_conic && v_cov2d &&
                     v_mean2d && v_scale && v_rot,
                 "null pointer");
    project_bwd_kernel<Param::ScaleRot><<<cdiv(num_points, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        num_points, scales2d, rotation, 0, 0, radii, conics, v_xy, v_conic, v_cov2d, v_mean2d,
        v_scale, v_rot);
    return check_launch(__func__);
}
