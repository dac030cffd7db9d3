// Per-pixel math of the Farneback displacement solve, shared by every kernel variant.
// Follows OpenCV 4.x optflowgf.cpp (FarnebackUpdateMatrices / FarnebackUpdateFlow_Blur) as
// restated in oracle/farneback_np.py.
#pragma once
#include "common.cuh"

#define FB_MAX_POLY_N 10
#define FB_MAX_WIN_RADIUS 16

struct PolyCoef {
    int n;
    float g[FB_MAX_POLY_N + 2], xg[FB_MAX_POLY_N + 2], xxg[FB_MAX_POLY_N + 2];  // taps k = 0..n (n + 1 when folded)
    float ig11, ig03, ig33, ig55;
};

// R (polynomial expansion) storage layout "4+1" for an image of n = w*h pixels and element type RT:
//   elements [0, 4n)  : (d/dy, d/dx, yy, xx) interleaved per pixel -> one 128-bit (fp32) load per tap
//   elements [4n, 5n) : xy plane
__device__ __forceinline__ float4 load_quad(const float* R, size_t px) {
    return __ldg(reinterpret_cast<const float4*>(R) + px);
}
__device__ __forceinline__ float4 load_quad(const __half* R, size_t px) {
    uint2 raw = __ldg(reinterpret_cast<const uint2*>(R) + px);
    float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float load_c4(const float* R, size_t plane, size_t px) { return __ldg(R + 4 * plane + px); }
__device__ __forceinline__ float load_c4(const __half* R, size_t plane, size_t px) {
    return __half2float(__ldg(R + 4 * plane + px));
}
__device__ __forceinline__ void store_quad(float* R, size_t px, float4 v) { reinterpret_cast<float4*>(R)[px] = v; }
__device__ __forceinline__ void store_quad(__half* R, size_t px, float4 v) {
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<unsigned*>(&lo);
    raw.y = *reinterpret_cast<unsigned*>(&hi);
    reinterpret_cast<uint2*>(R)[px] = raw;
}
__device__ __forceinline__ void store_c4(float* R, size_t plane, size_t px, float v) { R[4 * plane + px] = v; }
__device__ __forceinline__ void store_c4(__half* R, size_t plane, size_t px, float v) {
    R[4 * plane + px] = __float2half_rn(v);
}

__device__ __forceinline__ float fb_border(int d) { return d < 2 ? 0.14f : 0.4472f; }

// FarnebackUpdateMatrices for one pixel: R0 at (x, y) (already loaded into a[5]), R1 sampled
// bilinearly at (x+dx, y+dy) (falls back to R0 only when the sample leaves the image), 5-px border
// attenuation.  Coefficients are ordered (d/dy, d/dx, yy, xx, xy), stored "4+1" (see above).
template <typename RT>
__device__ __forceinline__ void fb_load_r0(const RT* __restrict__ R0, size_t plane, size_t at, float* a) {
    float4 q = load_quad(R0, at);
    a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
    a[4] = load_c4(R0, plane, at);
}

template <typename RT>
__device__ __forceinline__ void fb_update_matrix_pre(const float* a, const RT* __restrict__ R1, size_t plane, int w,
                                                     int h, int x, int y, float2 f, float* m) {
    float dx = f.x, dy = f.y;
    float fx = (float)x + dx, fy = (float)y + dy;
    // cvFloor; the float->int conversion saturates, so absurd displacements land outside the image
    int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    fx -= (float)x1;
    fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const size_t q = (size_t)y1 * w + x1;
        float4 q00 = load_quad(R1, q), q01 = load_quad(R1, q + 1), q10 = load_quad(R1, q + w), q11 = load_quad(R1, q + w + 1);
        float e00 = load_c4(R1, plane, q), e01 = load_c4(R1, plane, q + 1), e10 = load_c4(R1, plane, q + w),
              e11 = load_c4(R1, plane, q + w + 1);
        r2 = a00 * q00.x + a01 * q01.x + a10 * q10.x + a11 * q11.x;
        r3 = a00 * q00.y + a01 * q01.y + a10 * q10.y + a11 * q11.y;
        r4 = a00 * q00.z + a01 * q01.z + a10 * q10.z + a11 * q11.z;
        r5 = a00 * q00.w + a01 * q01.w + a10 * q10.w + a11 * q11.w;
        r6 = a00 * e00 + a01 * e01 + a10 * e10 + a11 * e11;
        r4 = (a[2] + r4) * 0.5f;
        r5 = (a[3] + r5) * 0.5f;
        r6 = (a[4] + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = a[2];
        r5 = a[3];
        r6 = a[4] * 0.5f;
    }
    r2 = (a[0] - r2) * 0.5f;
    r3 = (a[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        float s = (x < 5 ? fb_border(x) : 1.f) * (x >= w - 5 ? fb_border(w - x - 1) : 1.f) *
                  (y < 5 ? fb_border(y) : 1.f) * (y >= h - 5 ? fb_border(h - y - 1) : 1.f);
        r2 *= s; r3 *= s; r4 *= s; r5 *= s; r6 *= s;
    }
    m[0] = r4 * r4 + r6 * r6;  // G(1,1)
    m[1] = (r4 + r5) * r6;     // G(1,2)
    m[2] = r5 * r5 + r6 * r6;  // G(2,2)
    m[3] = r4 * r2 + r6 * r3;  // h(1)
    m[4] = r6 * r2 + r5 * r3;  // h(2)
}

template <typename RT>
__device__ __forceinline__ void fb_update_matrix(const RT* __restrict__ R0, const RT* __restrict__ R1, size_t plane,
                                                 int w, int h, int x, int y, float2 f, float* m) {
    float a[5];
    fb_load_r0<RT>(R0, plane, (size_t)y * w + x, a);
    fb_update_matrix_pre<RT>(a, R1, plane, w, h, x, y, f, m);
}

// 2x2 solve of FarnebackUpdateFlow_Blur, in double like cv2 (the determinant cancels).
__device__ __forceinline__ float2 fb_solve(const double* s, double scale) {
    double g11 = s[0] * scale, g12 = s[1] * scale, g22 = s[2] * scale, h1 = s[3] * scale, h2 = s[4] * scale;
    double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
    return make_float2((float)((g11 * h2 - g12 * h1) * idet), (float)((g22 * h1 - g12 * h2) * idet));
}
