// Pyramidal Lucas-Kanade on a dense grid of every `step`-th pixel, replacing
// cv2.calcOpticalFlowPyrLK as called at transflow/flow/methods/lukas_kanade.py:26-32
// (winSize (w, w), maxLevel, criteria COUNT+EPS 30 / 0.01, flags 0, minEigThreshold 1e-4;
// status is never read by the reference).  Algorithm: OpenCV 4.x lkpyramid.cpp, restated in
// oracle/lk_np.py: integer pyrDown, int16 Scharr derivatives, 14-bit bilinear weights, 5
// fractional bits on intensities, float32 2x2 solve.
//
// This path is bounded by integer ALU throughput, not HBM: one thread tracks one
// point; window rows slide through registers so every tap is loaded once per pass.
#include "common.cuh"

#include <stdlib.h>
#include <vector>

using namespace tf;

#define LK_W_BITS 14

struct LkLevel {
    int w, h;
    uint8_t* img[2];  // [0] = left/prev, [1] = right/next; level 0 aliases the caller's frames
    short2* deriv;    // Scharr (Ix, Iy) of the left image
};

struct tf_lucas_kanade {
    int H, W, win, max_level, step;
    int variant;  // 0 = warp-per-point, column mapping (default for win <= 16); 2 = warp-per-point, linear mapping;
                  // 1 = thread-per-point reference kernel
    int gw, gh;  // grid of tracked points
    std::vector<LkLevel> lv;
    float2* next_pts;
};

// cv::pyrDown (uint8): [1 4 6 4 1] x [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8
__global__ void __launch_bounds__(256) k_lk_pyrdown(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sw,
                                                    int sh, int dw, int dh) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int k[5] = {1, 4, 6, 4, 1};
    int cx[5];
#pragma unroll
    for (int j = 0; j < 5; j++) cx[j] = reflect101(2 * x + j - 2, sw);
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint8_t* row = src + (size_t)reflect101(2 * y + i - 2, sh) * sw;
        int r = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) r += k[j] * (int)row[cx[j]];
        acc += k[i] * r;
    }
    dst[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
}

// ScharrDerivInvoker: Ix = t0(x+1) - t0(x-1), t0 = 3*(I(y-1) + I(y+1)) + 10*I(y);
//                     Iy = 3*(t1(x-1) + t1(x+1)) + 10*t1(x), t1 = I(y+1) - I(y-1)
__global__ void __launch_bounds__(256) k_lk_scharr(const uint8_t* __restrict__ img, short2* __restrict__ deriv, int w,
                                                   int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
    const uint8_t* ru = img + (size_t)reflect101(y - 1, h) * w;
    const uint8_t* rc = img + (size_t)y * w;
    const uint8_t* rd = img + (size_t)reflect101(y + 1, h) * w;
    int t0l = ((int)ru[xl] + (int)rd[xl]) * 3 + (int)rc[xl] * 10;
    int t0r = ((int)ru[xr] + (int)rd[xr]) * 3 + (int)rc[xr] * 10;
    int t1l = (int)rd[xl] - (int)ru[xl], t1c = (int)rd[x] - (int)ru[x], t1r = (int)rd[xr] - (int)ru[xr];
    deriv[(size_t)y * w + x] = make_short2((short)(t0r - t0l), (short)((t1r + t1l) * 3 + t1c * 10));
}

struct LkWeights {
    int w00, w01, w10, w11;
};

__device__ __forceinline__ LkWeights lk_weights(float a, float b) {
    LkWeights q;
    const float s = (float)(1 << LK_W_BITS);
    q.w00 = __float2int_rn((1.f - a) * (1.f - b) * s);
    q.w01 = __float2int_rn(a * (1.f - b) * s);
    q.w10 = __float2int_rn((1.f - a) * b * s);
    q.w11 = (1 << LK_W_BITS) - q.w00 - q.w01 - q.w10;
    return q;
}

__device__ __forceinline__ int lk_descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// image tap with the REFLECT_101 border cv2 pads the pyramid with; derivative tap with zeros
__device__ __forceinline__ int lk_img(const uint8_t* __restrict__ img, int x, int y, int w, int h, bool inside) {
    if (!inside) {
        x = reflect101(x, w);
        y = reflect101(y, h);
    }
    return (int)__ldg(img + (size_t)y * w + x);
}
__device__ __forceinline__ short2 lk_der(const short2* __restrict__ d, int x, int y, int w, int h, bool inside) {
    if (!inside && ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h)) return make_short2(0, 0);
    return __ldg(d + (size_t)y * w + x);
}

// Interpolated template values of the window pixel (x, y) given the four taps.
struct LkTemplate {
    int ival, ix, iy;
};

// One pyramid level of LKTrackerInvoker for every grid point.
__global__ void __launch_bounds__(128) k_lk_track(const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
                                                  const short2* __restrict__ D, float2* __restrict__ next_pts, int w,
                                                  int h, int gw, int gh, int step, int win, int level, int is_top) {
    int gx = blockIdx.x * 32 + (threadIdx.x & 31), gy = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (gx >= gw || gy >= gh) return;
    size_t pid = (size_t)gy * gw + gx;
    const float lscale = 1.f / (float)(1 << level);
    const float half = (float)(win - 1) * 0.5f;
    float px = (float)(gx * step) * lscale, py = (float)(gy * step) * lscale;
    float2 np;
    if (is_top) {
        np = make_float2(px, py);
    } else {
        np = next_pts[pid];
        np.x *= 2.f;
        np.y *= 2.f;
    }
    next_pts[pid] = np;  // stored before any test (points that fail keep this value)
    px -= half;
    py -= half;
    int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) return;
    LkWeights q = lk_weights(px - (float)ipx, py - (float)ipy);
    const bool insideI = ipx >= 0 && ipy >= 0 && ipx + win < w && ipy + win < h;
    const float FLT_SCALE = 1.f / (float)(1 << 20);

    // pass 1: covariance of the interpolated derivatives
    float A11 = 0.f, A12 = 0.f, A22 = 0.f;
    for (int y = 0; y < win; y++) {
        short2 d0 = lk_der(D, ipx, ipy + y, w, h, insideI), d1 = lk_der(D, ipx, ipy + y + 1, w, h, insideI);
        for (int x = 0; x < win; x++) {
            short2 e0 = lk_der(D, ipx + x + 1, ipy + y, w, h, insideI), e1 = lk_der(D, ipx + x + 1, ipy + y + 1, w, h, insideI);
            int ixv = lk_descale(d0.x * q.w00 + e0.x * q.w01 + d1.x * q.w10 + e1.x * q.w11, LK_W_BITS);
            int iyv = lk_descale(d0.y * q.w00 + e0.y * q.w01 + d1.y * q.w10 + e1.y * q.w11, LK_W_BITS);
            A11 += (float)(ixv * ixv);
            A12 += (float)(ixv * iyv);
            A22 += (float)(iyv * iyv);
            d0 = e0;
            d1 = e1;
        }
    }
    A11 *= FLT_SCALE;
    A12 *= FLT_SCALE;
    A22 *= FLT_SCALE;
    float Dd = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (minEig < 1e-4f || Dd < 1.1920929e-07f) return;
    Dd = 1.f / Dd;

    float nx = np.x - half, ny = np.y - half;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < 30; j++) {
        int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= w || iny < -win || iny >= h) break;
        LkWeights r = lk_weights(nx - (float)inx, ny - (float)iny);
        const bool insideJ = inx >= 0 && iny >= 0 && inx + win < w && iny + win < h;
        float b1 = 0.f, b2 = 0.f;
        for (int y = 0; y < win; y++) {
            short2 d0 = lk_der(D, ipx, ipy + y, w, h, insideI), d1 = lk_der(D, ipx, ipy + y + 1, w, h, insideI);
            int i0 = lk_img(I, ipx, ipy + y, w, h, insideI), i1 = lk_img(I, ipx, ipy + y + 1, w, h, insideI);
            int j0 = lk_img(J, inx, iny + y, w, h, insideJ), j1 = lk_img(J, inx, iny + y + 1, w, h, insideJ);
            for (int x = 0; x < win; x++) {
                short2 e0 = lk_der(D, ipx + x + 1, ipy + y, w, h, insideI),
                       e1 = lk_der(D, ipx + x + 1, ipy + y + 1, w, h, insideI);
                int k0 = lk_img(I, ipx + x + 1, ipy + y, w, h, insideI), k1 = lk_img(I, ipx + x + 1, ipy + y + 1, w, h, insideI);
                int l0 = lk_img(J, inx + x + 1, iny + y, w, h, insideJ), l1 = lk_img(J, inx + x + 1, iny + y + 1, w, h, insideJ);
                int ival = lk_descale(i0 * q.w00 + k0 * q.w01 + i1 * q.w10 + k1 * q.w11, LK_W_BITS - 5);
                int ixv = lk_descale(d0.x * q.w00 + e0.x * q.w01 + d1.x * q.w10 + e1.x * q.w11, LK_W_BITS);
                int iyv = lk_descale(d0.y * q.w00 + e0.y * q.w01 + d1.y * q.w10 + e1.y * q.w11, LK_W_BITS);
                int diff = lk_descale(j0 * r.w00 + l0 * r.w01 + j1 * r.w10 + l1 * r.w11, LK_W_BITS - 5) - ival;
                b1 += (float)(diff * ixv);
                b2 += (float)(diff * iyv);
                d0 = e0; d1 = e1; i0 = k0; i1 = k1; j0 = l0; j1 = l1;
            }
        }
        b1 *= FLT_SCALE;
        b2 *= FLT_SCALE;
        float dx = (A12 * b2 - A22 * b1) * Dd, dy = (A12 * b1 - A11 * b2) * Dd;
        nx += dx;
        ny += dy;
        float2 o = make_float2(nx + half, ny + half);
        if ((double)dx * dx + (double)dy * dy <= 1e-4) {  // epsilon^2, epsilon = 0.01
            next_pts[pid] = o;
            break;
        }
        if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
            o.x -= dx * 0.5f;
            o.y -= dy * 0.5f;
            next_pts[pid] = o;
            break;
        }
        next_pts[pid] = o;
        pdx = dx;
        pdy = dy;
    }
}

// ---- warp-per-point tracker ---------------------------------------------------------------------
// One warp tracks one point: the win x win window is spread over the 32 lanes (pixel i -> lane i % 32),
// each lane keeps the interpolated template (I, dIx, dIy) of its pixels in registers for the whole
// level, the 2x2 sums are reduced with shuffles, and every lane takes the same convergence branch --
// so a straggler that needs 30 iterations no longer holds 31 other points hostage, and the template
// is interpolated once per level instead of once per iteration.  Sums are exact 64-bit integers.
__device__ __forceinline__ long long lk_warp_sum(long long v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

template <int KMAX>
__global__ void __launch_bounds__(256, 4) k_lk_track_warp(const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
                                                       const short2* __restrict__ D, float2* __restrict__ next_pts,
                                                       int w, int h, int gw, int gh, int step, int win, int level,
                                                       int is_top) {
    const int lane = threadIdx.x & 31;
    const long long pid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pid >= (long long)gw * gh) return;
    const int gy = (int)(pid / gw), gx = (int)(pid - (long long)gy * gw);
    const float lscale = 1.f / (float)(1 << level);
    const float half = (float)(win - 1) * 0.5f;
    float px = (float)(gx * step) * lscale, py = (float)(gy * step) * lscale;
    float2 np;
    if (is_top) {
        np = make_float2(px, py);
    } else {
        np = next_pts[pid];
        np.x *= 2.f;
        np.y *= 2.f;
    }
    if (lane == 0) next_pts[pid] = np;  // stored before any test (points that fail keep this value)
    px -= half;
    py -= half;
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -win || ipx >= w || ipy < -win || ipy >= h) return;
    const LkWeights q = lk_weights(px - (float)ipx, py - (float)ipy);
    const bool insideI = ipx >= 0 && ipy >= 0 && ipx + win < w && ipy + win < h;
    const int npx = win * win;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const float inv_win = 1.f / (float)win;
    const bool unit = (q.w01 | q.w10 | q.w11) == 0;  // then w00 == 1 << LK_W_BITS: the template is the raw tap

    // template of this lane's pixels + covariance of the interpolated derivatives
    int ival[KMAX];
    short2 dval[KMAX];
    int woff[KMAX];  // window pixel as an offset y * w + x (x, y are re-derived on the rare border path)
    long long a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
        int i = lane + 32 * k;
        ival[k] = 0;
        dval[k] = make_short2(0, 0);
        woff[k] = 0;
        if (i < npx) {
            // (i + 0.5) / win is at least 0.5 / win away from an integer: the float quotient truncates exactly
            int y = __float2int_rz(((float)i + 0.5f) * inv_win), x = i - y * win;
            woff[k] = y * w + x;
            int X = ipx + x, Y = ipy + y;
            int ixv, iyv;
            if (insideI) {
                // window inside the image (almost every point): no border logic; points on the integer grid
                // (level 0 of a dense or strided grid) have weights (1, 0, 0, 0) and need one tap, not four
                const size_t at = (size_t)Y * w + X;
                if (unit) {
                    ival[k] = (int)__ldg(I + at) << 5;
                    short2 d = __ldg(D + at);
                    ixv = d.x;
                    iyv = d.y;
                } else {
                    int i00 = __ldg(I + at), i01 = __ldg(I + at + 1), i10 = __ldg(I + at + w), i11 = __ldg(I + at + w + 1);
                    short2 d00 = __ldg(D + at), d01 = __ldg(D + at + 1), d10 = __ldg(D + at + w), d11 = __ldg(D + at + w + 1);
                    ival[k] = lk_descale(i00 * q.w00 + i01 * q.w01 + i10 * q.w10 + i11 * q.w11, LK_W_BITS - 5);
                    ixv = lk_descale(d00.x * q.w00 + d01.x * q.w01 + d10.x * q.w10 + d11.x * q.w11, LK_W_BITS);
                    iyv = lk_descale(d00.y * q.w00 + d01.y * q.w01 + d10.y * q.w10 + d11.y * q.w11, LK_W_BITS);
                }
            } else {
                int i00 = lk_img(I, X, Y, w, h, false), i01 = lk_img(I, X + 1, Y, w, h, false);
                int i10 = lk_img(I, X, Y + 1, w, h, false), i11 = lk_img(I, X + 1, Y + 1, w, h, false);
                short2 d00 = lk_der(D, X, Y, w, h, false), d01 = lk_der(D, X + 1, Y, w, h, false);
                short2 d10 = lk_der(D, X, Y + 1, w, h, false), d11 = lk_der(D, X + 1, Y + 1, w, h, false);
                ival[k] = lk_descale(i00 * q.w00 + i01 * q.w01 + i10 * q.w10 + i11 * q.w11, LK_W_BITS - 5);
                ixv = lk_descale(d00.x * q.w00 + d01.x * q.w01 + d10.x * q.w10 + d11.x * q.w11, LK_W_BITS);
                iyv = lk_descale(d00.y * q.w00 + d01.y * q.w01 + d10.y * q.w10 + d11.y * q.w11, LK_W_BITS);
            }
            dval[k] = make_short2((short)ixv, (short)iyv);
            a11 += (long long)(ixv * ixv);
            a12 += (long long)(ixv * iyv);
            a22 += (long long)(iyv * iyv);
        }
    }
    float A11 = (float)lk_warp_sum(a11) * FLT_SCALE, A12 = (float)lk_warp_sum(a12) * FLT_SCALE,
          A22 = (float)lk_warp_sum(a22) * FLT_SCALE;
    float Dd = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
    if (minEig < 1e-4f || Dd < 1.1920929e-07f) return;
    Dd = 1.f / Dd;

    float nx = np.x - half, ny = np.y - half;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < 30; j++) {
        int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= w || iny < -win || iny >= h) break;
        LkWeights r = lk_weights(nx - (float)inx, ny - (float)iny);
        const bool insideJ = inx >= 0 && iny >= 0 && inx + win < w && iny + win < h;
        long long b1 = 0, b2 = 0;
#pragma unroll
        for (int k = 0; k < KMAX; k++) {
            if (lane + 32 * k < npx) {
                int j00, j01, j10, j11;
                if (insideJ) {
                    const uint8_t* p = J + ((size_t)iny * w + inx) + woff[k];
                    j00 = __ldg(p); j01 = __ldg(p + 1); j10 = __ldg(p + w); j11 = __ldg(p + w + 1);
                } else {
                    const int i = lane + 32 * k, wy = __float2int_rz(((float)i + 0.5f) * inv_win), wx = i - wy * win;
                    const int X = inx + wx, Y = iny + wy;
                    j00 = lk_img(J, X, Y, w, h, false); j01 = lk_img(J, X + 1, Y, w, h, false);
                    j10 = lk_img(J, X, Y + 1, w, h, false); j11 = lk_img(J, X + 1, Y + 1, w, h, false);
                }
                int diff = lk_descale(j00 * r.w00 + j01 * r.w01 + j10 * r.w10 + j11 * r.w11, LK_W_BITS - 5) - ival[k];
                b1 += (long long)(diff * (int)dval[k].x);
                b2 += (long long)(diff * (int)dval[k].y);
            }
        }
        float fb1 = (float)lk_warp_sum(b1) * FLT_SCALE, fb2 = (float)lk_warp_sum(b2) * FLT_SCALE;
        float dx = (A12 * fb2 - A22 * fb1) * Dd, dy = (A12 * fb1 - A11 * fb2) * Dd;
        nx += dx;
        ny += dy;
        float2 o = make_float2(nx + half, ny + half);
        bool stop = (double)dx * dx + (double)dy * dy <= 1e-4;  // epsilon^2, epsilon = 0.01
        if (!stop && j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
            o.x -= dx * 0.5f;
            o.y -= dy * 0.5f;
            stop = true;
        }
        if (lane == 0) next_pts[pid] = o;
        if (stop) break;
        pdx = dx;
        pdy = dy;
    }
}

// ---- warp-per-point tracker, column mapping (default for windows up to 16 x 16) -------------------------------
// Same algorithm and the same exact integer sums as k_lk_track_warp, reorganised around what bounds the tracker
// (instruction issue and L1 tag lookups):
//   * lane = (window column, upper / lower half of the rows): lanes 0..WIN-1 own rows 0..R0-1 of their column, lanes
//     16..16+WIN-1 rows R0..WIN-1.  A load instruction then reads ONE image row per half-warp (15 consecutive bytes:
//     2 cache lines per instruction instead of one line per window row), and a lane walks down its column: the
//     bottom taps of a window pixel are the top taps of the next one, so a pixel costs 2 loads instead of 4 (18 per
//     lane and iteration instead of 32);
//   * the per-lane partial sums fit 32 bits (<= 8 products of 13 x 12 bits), so the inner loop is 32-bit IMADs and the
//     warp total is formed by four REDUX instructions (low / high 16 bits separately, recombined exactly in 64 bits)
//     instead of two 64-bit shuffle trees.
// The value of every sum is identical to k_lk_track_warp's, so the tracks are bit-identical to that kernel.
__device__ __forceinline__ long long lk_redux_exact(int v) {
    const int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return (long long)hi * 65536ll + (long long)lo;
}

template <int WIN, bool SHFL>
__global__ void __launch_bounds__(256, 4) k_lk_track_cols(const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
                                                         const short2* __restrict__ D, float2* __restrict__ next_pts,
                                                         int w, int h, int gw, int gh, int step, int level, int is_top) {
    static_assert(WIN >= 3 && WIN <= 16, "one half-warp per window row group");
    constexpr int R0 = (WIN + 1) / 2;  // rows of the upper half
    const int lane = threadIdx.x & 31;
    const int col = lane & 15, half = lane >> 4;
    const bool active = col < WIN;
    const int r0 = half ? R0 : 0;             // first window row of this lane
    const int nr = half ? WIN - R0 : R0;      // its number of rows
    const long long pid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pid >= (long long)gw * gh) return;
    const int gy = (int)(pid / gw), gx = (int)(pid - (long long)gy * gw);
    const float lscale = 1.f / (float)(1 << level);
    const float halfw = (float)(WIN - 1) * 0.5f;
    float px = (float)(gx * step) * lscale, py = (float)(gy * step) * lscale;
    float2 np;
    if (is_top) {
        np = make_float2(px, py);
    } else {
        np = next_pts[pid];
        np.x *= 2.f;
        np.y *= 2.f;
    }
    if (lane == 0) next_pts[pid] = np;  // stored before any test (points that fail keep this value)
    px -= halfw;
    py -= halfw;
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) return;
    const LkWeights q = lk_weights(px - (float)ipx, py - (float)ipy);
    const bool insideI = ipx >= 0 && ipy >= 0 && ipx + WIN < w && ipy + WIN < h;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    const bool unit = (q.w01 | q.w10 | q.w11) == 0;  // then w00 == 1 << LK_W_BITS: the template is the raw tap

    // template of this lane's column segment + covariance of the interpolated derivatives
    int ival[R0], dxv[R0], dyv[R0];
    int a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
    for (int k = 0; k < R0; k++) ival[k] = dxv[k] = dyv[k] = 0;
    if (active) {
        if (insideI) {
            const size_t at0 = (size_t)(ipy + r0) * w + ipx + col;
            if (unit) {
#pragma unroll
                for (int k = 0; k < R0; k++) {
                    if (k < nr) {
                        const size_t at = at0 + (size_t)k * w;
                        ival[k] = (int)__ldg(I + at) << 5;
                        const short2 d = __ldg(D + at);
                        dxv[k] = d.x;
                        dyv[k] = d.y;
                    }
                }
            } else {
                int ia[R0 + 1], ib[R0 + 1];
                short2 da[R0 + 1], db[R0 + 1];
#pragma unroll
                for (int k = 0; k <= R0; k++) {
                    ia[k] = ib[k] = 0;
                    da[k] = db[k] = make_short2(0, 0);
                    if (k <= nr) {
                        const size_t at = at0 + (size_t)k * w;
                        ia[k] = __ldg(I + at); ib[k] = __ldg(I + at + 1);
                        da[k] = __ldg(D + at); db[k] = __ldg(D + at + 1);
                    }
                }
#pragma unroll
                for (int k = 0; k < R0; k++) {
                    if (k < nr) {
                        ival[k] = lk_descale(ia[k] * q.w00 + ib[k] * q.w01 + ia[k + 1] * q.w10 + ib[k + 1] * q.w11, LK_W_BITS - 5);
                        dxv[k] = lk_descale(da[k].x * q.w00 + db[k].x * q.w01 + da[k + 1].x * q.w10 + db[k + 1].x * q.w11, LK_W_BITS);
                        dyv[k] = lk_descale(da[k].y * q.w00 + db[k].y * q.w01 + da[k + 1].y * q.w10 + db[k + 1].y * q.w11, LK_W_BITS);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < R0; k++) {
                if (k < nr) {
                    const int X = ipx + col, Y = ipy + r0 + k;
                    int i00 = lk_img(I, X, Y, w, h, false), i01 = lk_img(I, X + 1, Y, w, h, false);
                    int i10 = lk_img(I, X, Y + 1, w, h, false), i11 = lk_img(I, X + 1, Y + 1, w, h, false);
                    short2 d00 = lk_der(D, X, Y, w, h, false), d01 = lk_der(D, X + 1, Y, w, h, false);
                    short2 d10 = lk_der(D, X, Y + 1, w, h, false), d11 = lk_der(D, X + 1, Y + 1, w, h, false);
                    ival[k] = lk_descale(i00 * q.w00 + i01 * q.w01 + i10 * q.w10 + i11 * q.w11, LK_W_BITS - 5);
                    dxv[k] = lk_descale(d00.x * q.w00 + d01.x * q.w01 + d10.x * q.w10 + d11.x * q.w11, LK_W_BITS);
                    dyv[k] = lk_descale(d00.y * q.w00 + d01.y * q.w01 + d10.y * q.w10 + d11.y * q.w11, LK_W_BITS);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R0; k++) {  // rows beyond nr hold zeros
            a11 += dxv[k] * dxv[k];
            a12 += dxv[k] * dyv[k];
            a22 += dyv[k] * dyv[k];
        }
    }
    float A11 = (float)lk_redux_exact(a11) * FLT_SCALE, A12 = (float)lk_redux_exact(a12) * FLT_SCALE,
          A22 = (float)lk_redux_exact(a22) * FLT_SCALE;
    float Dd = A11 * A22 - A12 * A12;
    float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
    if (minEig < 1e-4f || Dd < 1.1920929e-07f) return;
    Dd = 1.f / Dd;

    float nx = np.x - halfw, ny = np.y - halfw;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < 30; j++) {
        int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) break;
        const LkWeights r = lk_weights(nx - (float)inx, ny - (float)iny);
        const bool insideJ = inx >= 0 && iny >= 0 && inx + WIN < w && iny + WIN < h;
        int b1 = 0, b2 = 0;
        if (SHFL && WIN < 16 && insideJ) {
            // every lane of a half-warp loads its column's taps (lane 15: the column right of the window); the right
            // taps come from the neighbouring lane by shuffle instead of a second load
            const uint8_t* p = J + ((size_t)(iny + r0) * w + inx + col);
            int ja[R0 + 1], jb[R0 + 1];
#pragma unroll
            for (int k = 0; k <= R0; k++) {
                ja[k] = 0;
                if (k <= nr) ja[k] = __ldg(p + (size_t)k * w);
            }
#pragma unroll
            for (int k = 0; k <= R0; k++) jb[k] = __shfl_down_sync(0xffffffffu, ja[k], 1);
            if (active) {
#pragma unroll
                for (int k = 0; k < R0; k++) {
                    if (k < nr) {
                        const int diff = lk_descale(ja[k] * r.w00 + jb[k] * r.w01 + ja[k + 1] * r.w10 + jb[k + 1] * r.w11,
                                                    LK_W_BITS - 5) - ival[k];
                        b1 += diff * dxv[k];
                        b2 += diff * dyv[k];
                    }
                }
            }
        } else if (active) {
            if (insideJ) {
                const uint8_t* p = J + ((size_t)(iny + r0) * w + inx + col);
                int ja[R0 + 1], jb[R0 + 1];
#pragma unroll
                for (int k = 0; k <= R0; k++) {
                    ja[k] = jb[k] = 0;
                    if (k <= nr) {
                        ja[k] = __ldg(p + (size_t)k * w);
                        jb[k] = __ldg(p + (size_t)k * w + 1);
                    }
                }
#pragma unroll
                for (int k = 0; k < R0; k++) {
                    if (k < nr) {
                        const int diff = lk_descale(ja[k] * r.w00 + jb[k] * r.w01 + ja[k + 1] * r.w10 + jb[k + 1] * r.w11,
                                                    LK_W_BITS - 5) - ival[k];
                        b1 += diff * dxv[k];
                        b2 += diff * dyv[k];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < R0; k++) {
                    if (k < nr) {
                        const int X = inx + col, Y = iny + r0 + k;
                        const int j00 = lk_img(J, X, Y, w, h, false), j01 = lk_img(J, X + 1, Y, w, h, false);
                        const int j10 = lk_img(J, X, Y + 1, w, h, false), j11 = lk_img(J, X + 1, Y + 1, w, h, false);
                        const int diff = lk_descale(j00 * r.w00 + j01 * r.w01 + j10 * r.w10 + j11 * r.w11, LK_W_BITS - 5) - ival[k];
                        b1 += diff * dxv[k];
                        b2 += diff * dyv[k];
                    }
                }
            }
        }
        float fb1 = (float)lk_redux_exact(b1) * FLT_SCALE, fb2 = (float)lk_redux_exact(b2) * FLT_SCALE;
        float dx = (A12 * fb2 - A22 * fb1) * Dd, dy = (A12 * fb1 - A11 * fb2) * Dd;
        nx += dx;
        ny += dy;
        float2 o = make_float2(nx + halfw, ny + halfw);
        bool stop = (double)dx * dx + (double)dy * dy <= 1e-4;  // epsilon^2, epsilon = 0.01
        if (!stop && j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
            o.x -= dx * 0.5f;
            o.y -= dy * 0.5f;
            stop = true;
        }
        if (lane == 0) next_pts[pid] = o;
        if (stop) break;
        pdx = dx;
        pdy = dy;
    }
}

// flow = p1 - p0, block-replicated (numpy.kron) back to (H, W), optional final clip
__global__ void __launch_bounds__(256) k_lk_flow_out(const float2* __restrict__ next_pts, float2* __restrict__ flow,
                                                     int H, int W, int gw, int step, int clip) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int gx = x / step, gy = y / step;
    float2 n = __ldg(next_pts + (size_t)gy * gw + gx);
    float2 f = make_float2(n.x - (float)(gx * step), n.y - (float)(gy * step));
    if (clip) {
        f.x = fminf(fmaxf(f.x, (float)(-x)), (float)(W - 1 - x));
        f.y = fminf(fmaxf(f.y, (float)(-y)), (float)(H - 1 - y));
    }
    flow[(size_t)y * W + x] = f;
}

extern "C" int tf_lk_destroy(tf_lucas_kanade* h) {
    if (!h) return TF_OK;
    for (size_t i = 0; i < h->lv.size(); i++) {
        if (i > 0) {
            cudaFree(h->lv[i].img[0]);
            cudaFree(h->lv[i].img[1]);
        }
        cudaFree(h->lv[i].deriv);
    }
    cudaFree(h->next_pts);
    delete h;
    return TF_OK;
}

extern "C" int tf_lk_create(tf_lucas_kanade** out, int height, int width, int win_size, int max_level, int step) {
    TF_REQUIRE(out, TF_ERR_INVALID_ARG, "tf_lk_create: null out");
    TF_REQUIRE(height >= 2 && width >= 2 && (size_t)height * width < (1u << 30), TF_ERR_SHAPE,
               "tf_lk_create: bad shape %dx%d", height, width);
    TF_REQUIRE(win_size >= 3 && win_size <= 63, TF_ERR_INVALID_ARG, "lk_window_size must be in [3, 63], got %d", win_size);
    TF_REQUIRE(max_level >= 0 && max_level <= 10, TF_ERR_INVALID_ARG, "lk_max_level must be in [0, 10], got %d", max_level);
    TF_REQUIRE(step >= 1, TF_ERR_INVALID_ARG, "lk_step must be >= 1, got %d", step);
    if (int e = require_sm100()) return e;
    tf_lucas_kanade* h = new (std::nothrow) tf_lucas_kanade();
    TF_REQUIRE(h, TF_ERR_CUDA, "out of host memory");
    h->H = height; h->W = width; h->win = win_size; h->max_level = max_level; h->step = step;
    {
        const char* v = getenv("TFB200_LK_VARIANT");
        h->variant = v ? atoi(v) : 0;
    }
    h->gw = ceil_div(width, step);
    h->gh = ceil_div(height, step);
    h->next_pts = nullptr;
    auto bail = [&](int e) { tf_lk_destroy(h); return e; };
    int w = width, hh = height;
    for (int l = 0; l <= max_level; l++) {
        LkLevel L;
        memset(&L, 0, sizeof(L));
        L.w = w;
        L.h = hh;
        size_t n = (size_t)w * hh;
        if (l > 0 && (cudaMalloc(&L.img[0], n) != cudaSuccess || cudaMalloc(&L.img[1], n) != cudaSuccess))
            return bail(fail(TF_ERR_CUDA, "tf_lk_create: allocation failed"));
        if (cudaMalloc(&L.deriv, n * sizeof(short2)) != cudaSuccess) return bail(fail(TF_ERR_CUDA, "tf_lk_create: allocation failed"));
        h->lv.push_back(L);
        // buildOpticalFlowPyramid stops when the NEXT level would not exceed the window
        w = (w + 1) / 2;
        hh = (hh + 1) / 2;
        if (w <= win_size || hh <= win_size) break;
    }
    if (cudaMalloc(&h->next_pts, (size_t)h->gw * h->gh * sizeof(float2)) != cudaSuccess)
        return bail(fail(TF_ERR_CUDA, "tf_lk_create: allocation failed"));
    *out = h;
    return TF_OK;
}

extern "C" int tf_lk_run(tf_lucas_kanade* h, const uint8_t* left, const uint8_t* right, float* flow, int clip,
                         void* stream) {
    TF_REQUIRE(h && left && right && flow, TF_ERR_INVALID_ARG, "tf_lk_run: null argument");
    TF_REQUIRE(((uintptr_t)flow & 7) == 0, TF_ERR_INVALID_ARG, "tf_lk_run: flow must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    h->lv[0].img[0] = const_cast<uint8_t*>(left);
    h->lv[0].img[1] = const_cast<uint8_t*>(right);
    for (size_t l = 0; l < h->lv.size(); l++) {
        LkLevel& L = h->lv[l];
        if (l > 0) {
            LkLevel& P = h->lv[l - 1];
            dim3 grid(ceil_div(L.w, 256), L.h);
            for (int s = 0; s < 2; s++) {
                k_lk_pyrdown<<<grid, 256, 0, st>>>(P.img[s], L.img[s], P.w, P.h, L.w, L.h);
                TF_LAUNCHED();
            }
        }
        k_lk_scharr<<<dim3(ceil_div(L.w, 256), L.h), 256, 0, st>>>(L.img[0], L.deriv, L.w, L.h);
        TF_LAUNCHED();
    }
    int top = (int)h->lv.size() - 1;
    dim3 tgrid(ceil_div(h->gw, 32), ceil_div(h->gh, 4));
    long long npoints = (long long)h->gw * h->gh;
    unsigned wgrid = (unsigned)((npoints + 7) / 8);
    int kmax = ceil_div(h->win * h->win, 32);
    for (int l = top; l >= 0; l--) {
        LkLevel& L = h->lv[l];
        {
            ScopedKernelTimer timer(l == 0 ? TFK_LK_TRACK_FINEST : -1, st);
#define TF_LKW(K)                                                                                                  \
    k_lk_track_warp<K><<<wgrid, 256, 0, st>>>(L.img[0], L.img[1], L.deriv, h->next_pts, L.w, L.h, h->gw, h->gh, h->step, \
                                              h->win, l, l == top)
            if (h->variant == 0 && h->win == 15)
                k_lk_track_cols<15, true><<<wgrid, 256, 0, st>>>(L.img[0], L.img[1], L.deriv, h->next_pts, L.w, L.h, h->gw, h->gh,
                                                                 h->step, l, l == top);
            else if (h->variant == 3 && h->win == 15)
                k_lk_track_cols<15, false><<<wgrid, 256, 0, st>>>(L.img[0], L.img[1], L.deriv, h->next_pts, L.w, L.h, h->gw, h->gh,
                                                                  h->step, l, l == top);
            else if (h->variant == 0 && h->win == 9)
                k_lk_track_cols<9, true><<<wgrid, 256, 0, st>>>(L.img[0], L.img[1], L.deriv, h->next_pts, L.w, L.h, h->gw, h->gh,
                                                                h->step, l, l == top);
            else if (h->variant == 1 || kmax > 31)   // thread-per-point reference kernel (and very large windows)
                k_lk_track<<<tgrid, 128, 0, st>>>(L.img[0], L.img[1], L.deriv, h->next_pts, L.w, L.h, h->gw, h->gh,
                                                  h->step, h->win, l, l == top);
            else if (kmax <= 8) TF_LKW(8);
            else if (kmax <= 14) TF_LKW(14);
            else TF_LKW(31);
#undef TF_LKW
        }
        TF_LAUNCHED();
    }
    k_lk_flow_out<<<dim3(ceil_div(h->W, 256), h->H), 256, 0, st>>>(h->next_pts, reinterpret_cast<float2*>(flow), h->H, h->W,
                                                                   h->gw, h->step, clip);
    TF_LAUNCHED();
    h->lv[0].img[0] = h->lv[0].img[1] = nullptr;
    return TF_OK;
}
