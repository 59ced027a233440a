// gi2d_project_core.cuh -- per-Gaussian projection math shared by the standalone
// projection kernels (gi2d_project.cu) and the fused fit step (gi2d_fit.cu).
#pragma once
#include "gi2d_common.cuh"

namespace gi2d {

// Result of projecting one Gaussian (reference: csrc/foward2d.cu:12-69,130-187,192-288).
struct Projected {
    float x, y;        // xys   (0 when culled before the write point)
    float a, b, c;     // conic (0 when culled before the write point)
    int radius;        // radii (0 when culled)
    int ntiles;        // num_tiles_hit
    TileBox box;       // valid when ntiles > 0
};

// Shared tail of the three forward kernels: bounds -> cull -> write point -> tile area.
// `bbox_uses_int_radius` reproduces foward2d.cu:177 (scale-rot passes the int radii[idx]).
__device__ __forceinline__ Projected finish_projection(float cx, float cy, float sx, float sxy,
                                                       float sy, float clip_coe, float radius_clip,
                                                       int tiles_x, int tiles_y,
                                                       bool bbox_uses_int_radius) {
    Projected p;
    p.x = p.y = p.a = p.b = p.c = 0.f;
    p.radius = 0;
    p.ntiles = 0;
    p.box = TileBox{0, 0, 0, 0};
    const Cov2dBounds cb = cov2d_bounds(sx, sxy, sy, clip_coe);
    if (!cb.ok) return p;                       // zero determinant
    if (cb.ry < radius_clip) return p;          // false for NaN, as in the reference
    p.a = cb.a; p.b = cb.b; p.c = cb.c;
    p.x = cx; p.y = cy;
    p.radius = f2i_rz(cb.rx);                   // (int)radius.x : saturating, NaN -> 0
    const float r = bbox_uses_int_radius ? (float)p.radius : cb.rx;
    p.box = tile_bbox(cx, cy, r, tiles_x, tiles_y);
    const int area = (p.box.x1 - p.box.x0) * (p.box.y1 - p.box.y0);
    p.ntiles = area > 0 ? area : 0;
    return p;
}

// R1: covariance parameterisation, centre in pixels.
__device__ __forceinline__ Projected project_cov(float mx, float my, float sx, float sxy, float sy,
                                                 float clip_coe, float radius_clip, int tiles_x,
                                                 int tiles_y) {
    return finish_projection(mx, my, sx, sxy, sy, clip_coe, radius_clip, tiles_x, tiles_y, false);
}

// R2: Cholesky parameterisation; centre = fma(0.5*W, x, 0.5*W); cov = (l11^2, l11*l21,
// fma(l21,l21,l22^2)) exactly as the -O3 SASS contracts it.
__device__ __forceinline__ Projected project_chol(float mx, float my, float l11, float l21,
                                                  float l22, int img_w, int img_h, float clip_coe,
                                                  float radius_clip, int tiles_x, int tiles_y) {
    const float hw = __fmul_rn((float)(unsigned)img_w, 0.5f);
    const float hh = __fmul_rn((float)(unsigned)img_h, 0.5f);
    const float cx = __fmaf_rn(hw, mx, hw);
    const float cy = __fmaf_rn(my, hh, hh);
    const float sxy = __fmul_rn(l11, l21);
    const float sx = __fmul_rn(l11, l11);
    const float sy = __fmaf_rn(l21, l21, __fmul_rn(l22, l22));
    return finish_projection(cx, cy, sx, sxy, sy, clip_coe, radius_clip, tiles_x, tiles_y, false);
}

// R3: scale/rotation.  glm column-major algebra of helpers.cuh:579-598 hand-expanded,
// including the 0*x products glm's full mat2*mat2 leaves in (they matter for inf/NaN).
__device__ __forceinline__ void rs_to_cov(float sclx, float scly, float theta, float &sx, float &sxy,
                                          float &sy) {
    const float cs = cosf(theta), sn = sinf(theta);
    const float zc = __fmul_rn(0.f, cs), zs = __fmul_rn(0.f, sn);
    const float m10 = __fmaf_rn(sn, scly, zc);
    const float m11 = __fmaf_rn(cs, scly, -zs);
    const float m01 = __fmaf_rn(sn, -sclx, zc);
    const float m00 = __fmaf_rn(cs, sclx, zs);
    sxy = __fmaf_rn(m00, m01, __fmul_rn(m10, m11));
    sx = __fmaf_rn(m00, m00, __fmul_rn(m10, m10));
    sy = __fmaf_rn(m01, m01, __fmul_rn(m11, m11));
}

__device__ __forceinline__ Projected project_rs(float mx, float my, float sclx, float scly,
                                                float theta, float clip_coe, float radius_clip,
                                                int tiles_x, int tiles_y) {
    float sx, sxy, sy;
    rs_to_cov(sclx, scly, theta, sx, sxy, sy);
    return finish_projection(mx, my, sx, sxy, sy, clip_coe, radius_clip, tiles_x, tiles_y, true);
}

// helpers.cuh:384-395 cov2d_to_conic_vjp:  V = -X G X,  v_cov = (V00, V01+V10, V11)
// with X = [[a,b],[b,c]] and G = [[g0,g1],[g1,g2]].
__device__ __forceinline__ void conic_vjp(float a, float b, float c, float g0, float g1, float g2,
                                          float &v0, float &v1, float &v2) {
    // T = X*G
    const float t00 = a * g0 + b * g1, t01 = a * g1 + b * g2;
    const float t10 = b * g0 + c * g1, t11 = b * g1 + c * g2;
    // V = -(T*X)
    const float w00 = t00 * a + t01 * b, w01 = t00 * b + t01 * c;
    const float w10 = t10 * a + t11 * b, w11 = t10 * b + t11 * c;
    v0 = -w00;
    v1 = -(w01 + w10);
    v2 = -w11;
}

}  // namespace gi2d
