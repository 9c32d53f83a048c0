// MUFU approximations and a branch-free atan2 for the sparse stage (included by detect.cu before the
// orientation and descriptor kernels).  All are accurate to a few float32 ulp; where they are used
// the consumers are continuous in them (trilinear / Gaussian weights) or their error is far below
// the float32 spacing of the quantity they feed (angles in degrees near 360).
#pragma once

__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// rad2deg(atan2(y, x)) mod 360 (sift_impl.py:270-271, :416-417) without branches: atan(t) on [0, 1] as
// t * P(t^2) (minimax in degrees, max error 2.1e-6 deg, 7e-6 deg evaluated in float32 -- below the
// 3e-5 deg float32 spacing of an angle near 360), then the octant folds.
__device__ __forceinline__ float atan2_deg_fast(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mn * fast_rcp(fmaxf(mx, 1e-30f));   // 0 when both are 0, like atan2(0, 0)
    const float s = t * t;
    float p = -2.323093867e-01f;
    p = __fmaf_rn(p, s, 1.252654464e+00f);
    p = __fmaf_rn(p, s, -3.203539237e+00f);
    p = __fmaf_rn(p, s, 5.524571288e+00f);
    p = __fmaf_rn(p, s, -7.969057386e+00f);
    p = __fmaf_rn(p, s, 1.142854021e+01f);
    p = __fmaf_rn(p, s, -1.909660354e+01f);
    p = __fmaf_rn(p, s, 5.729574144e+01f);
    float r = p * t;
    r = ay > ax ? 90.f - r : r;
    r = x < 0.f ? 180.f - r : r;
    r = y < 0.f ? 360.f - r : r;
    return r;
}
