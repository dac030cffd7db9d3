// Per-pixel math of the Farneback displacement solve, shared by every kernel variant.
// Follows OpenCV 4.x optflowgf.cpp (FarnebackUpdateMatrices / FarnebackUpdateFlow_Blur) as
// restated in oracle/farneback_np.py.
#pragma once
#include "common.cuh"

#define FB_MAX_POLY_N 10
#define FB_MAX_WIN_RADIUS 16

struct PolyCoef {
    int n;
    float g[FB_MAX_POLY_N + 1], xg[FB_MAX_POLY_N + 1], xxg[FB_MAX_POLY_N + 1];  // taps k = 0..n
    float ig11, ig03, ig33, ig55;
};

__device__ __forceinline__ void store_r(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_r(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float load_r(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_r(const __half* p) { return __half2float(__ldg(p)); }

__device__ __forceinline__ float fb_border(int d) { return d < 2 ? 0.14f : 0.4472f; }

// FarnebackUpdateMatrices for one pixel: R0 at (x, y) (already loaded into a[5]), R1 sampled
// bilinearly at (x+dx, y+dy) (falls back to R0 only when the sample leaves the image), 5-px border
// attenuation.  R planes are ordered (d/dy, d/dx, yy, xx, xy), each w*h elements.
template <typename RT>
__device__ __forceinline__ void fb_load_r0(const RT* __restrict__ R0, size_t plane, size_t at, float* a) {
#pragma unroll
    for (int c = 0; c < 5; c++) a[c] = load_r(R0 + c * plane + at);
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
        const RT* p = R1 + ((size_t)y1 * w + x1);
        r2 = a00 * load_r(p) + a01 * load_r(p + 1) + a10 * load_r(p + w) + a11 * load_r(p + w + 1);
        p += plane;
        r3 = a00 * load_r(p) + a01 * load_r(p + 1) + a10 * load_r(p + w) + a11 * load_r(p + w + 1);
        p += plane;
        r4 = a00 * load_r(p) + a01 * load_r(p + 1) + a10 * load_r(p + w) + a11 * load_r(p + w + 1);
        p += plane;
        r5 = a00 * load_r(p) + a01 * load_r(p + 1) + a10 * load_r(p + w) + a11 * load_r(p + w + 1);
        p += plane;
        r6 = a00 * load_r(p) + a01 * load_r(p + 1) + a10 * load_r(p + w) + a11 * load_r(p + w + 1);
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
