// Farneback dense optical flow on device (replaces cv2.calcOpticalFlowFarneback as called at
// transflow/flow/sources/cv.py:477-490, flags = 0).  Algorithm: OpenCV 4.x
// modules/video/src/optflowgf.cpp (CPU path), restated in oracle/farneback_np.py.
//
// Stages per pyramid level (coarse -> fine):
//   prepare : Gaussian blur of the FULL-RES frame + bilinear resize (separable, two passes)
//             -> polynomial expansion R (5 coefficient planes, fp32 or fp16 storage)
//   solve   : flow init (zeros / x(1/pyr_scale) bilinear upsample) -> iterations x
//             [update-matrices -> (2m+1)^2 box sums -> 2x2 solve]
//
// Two solve variants share the same per-pixel math:
//   variant 1 ("reference kernels"): M and the vertical sums are materialised in HBM;
//   variant 0 ("fused streaming")  : one kernel per iteration, see fb_iter.cuh.
#include "common.cuh"
#include "fb_math.cuh"

#include <math.h>
#include <vector>

using namespace tf;

// ---------------------------------------------------------------------------------------------
// host-side tables
// ---------------------------------------------------------------------------------------------
static int cv_round(double v) { return (int)nearbyint(v); }  // half-to-even, like cvRound

static std::vector<float> gaussian_kernel(int ksz, double sigma) {
    // cv::getGaussianKernel(ksz, sigma, CV_32F)
    static const float k1[] = {1.f};
    static const float k3[] = {0.25f, 0.5f, 0.25f};
    static const float k5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    static const float k7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
    if (sigma <= 0 && ksz <= 7 && (ksz & 1)) {
        const float* t = ksz == 1 ? k1 : ksz == 3 ? k3 : ksz == 5 ? k5 : k7;
        return std::vector<float>(t, t + ksz);
    }
    double sig = sigma > 0 ? sigma : ((ksz - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> g(ksz);
    double sum = 0;
    for (int i = 0; i < ksz; i++) {
        double x = i - (ksz - 1) * 0.5;
        g[i] = exp(-0.5 / (sig * sig) * x * x);
        sum += g[i];
    }
    std::vector<float> out(ksz);
    for (int i = 0; i < ksz; i++) out[i] = (float)(g[i] / sum);
    return out;
}

// cv::resize INTER_LINEAR source index / weight for one axis
static void linear_table(int dst, int src, std::vector<int>& s, std::vector<float>& t) {
    s.resize(dst);
    t.resize(dst);
    double scale = (double)src / dst;
    for (int d = 0; d < dst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int si = (int)floorf(f);
        float ti = f - (float)si;
        if (si < 0) { si = 0; ti = 0.f; }
        if (si >= src - 1) { si = src - 1; ti = 0.f; }
        s[d] = si;
        t[d] = ti;
    }
}

// FarnebackPrepareGaussian: taps and the four inverse-Gram terms
static void prepare_poly(int n, double sigma, PolyCoef& pc) {
    if (sigma < 1.1920929e-07) sigma = n * 0.3;
    std::vector<float> g(2 * n + 1), xg(2 * n + 1), xxg(2 * n + 1);
    double s = 0;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)exp(-x * x / (2 * sigma * sigma));
        s += g[x + n];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            double gg = (double)g[y + n] * g[x + n];
            G[0][0] += gg;
            G[1][1] += gg * x * x;
            G[3][3] += gg * x * x * x * x;
            G[5][5] += gg * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    // invert the 6x6 by Gauss-Jordan with partial pivoting (double)
    double A[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++) A[i][j] = j < 6 ? G[i][j] : (j - 6 == i ? 1.0 : 0.0);
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        for (int j = 0; j < 12; j++) std::swap(A[c][j], A[piv][j]);
        double d = 1.0 / A[c][c];
        for (int j = 0; j < 12; j++) A[c][j] *= d;
        for (int r = 0; r < 6; r++)
            if (r != c) {
                double f = A[r][c];
                for (int j = 0; j < 12; j++) A[r][j] -= f * A[c][j];
            }
    }
    pc.n = n;
    for (int k = 0; k <= n; k++) {
        pc.g[k] = g[n + k];
        pc.xg[k] = xg[n + k];
        pc.xxg[k] = xxg[n + k];
    }
    pc.ig11 = (float)A[1][6 + 1];
    pc.ig03 = (float)A[0][6 + 3];
    pc.ig33 = (float)A[3][6 + 3];
    pc.ig55 = (float)A[5][6 + 5];
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
// Frame slots and solve lanes.  Slots 0 / 1 and lane 0 exist from tf_farneback_create; slot 2 and lane 1 are
// allocated on first use (tf_farneback_step_lane): with three slots and two lanes, pair (t, t+1) can be solved on
// one stream while pair (t+1, t+2) is solved on another and frame t+2's expansion is being built -- every frame is
// still prepared exactly once.
#define FB_SLOTS 3
#define FB_LANES 2

struct FbLevel {
    int k, w, h, ksz;
    double sigma;
    float* gk;   // device Gaussian taps (ksz)
    int* sx;     // device resize tables
    float* tx;
    int* sy;
    float* ty;
    int br_rows, br_pitch;  // fused blur+resize: largest footprint of a br_tw x br_th tile (0 = use the two passes)
    int br_tw, br_th;       // tile of the fused blur+resize kernel (level pixels), a power of two wide
    size_t br_smem;
    float* img;     // pyramid image of the frame prepared last (h x w)
    void* R[FB_SLOTS];        // polynomial expansion per slot, 5*h*w elements (float or __half) in the "4+1" layout
    float2* flow[FB_LANES];   // per-level flow of a lane (the finest level writes into the caller's buffer)
    float2* flow2[FB_LANES];  // ping-pong partner for the fused iteration kernel
    int* fsx;       // resize tables mapping the next-coarser level's flow onto this level
    float* ftx;
    int* fsy;
    float* fty;
};

struct tf_farneback {
    int H, W;
    double pyr_scale;
    int levels_req, winsize, iterations, poly_n, flags, r_fp16;
    int publish_finest;        // debug: also write the finest level's blurred image (fused into polyexp otherwise)
    double poly_sigma;
    PolyCoef pc;
    PolyCoef pcf;              // pc convolved with the finest level's [1/4, 1/2, 1/4] blur (radius poly_n + 1)
    std::vector<FbLevel> lv;  // coarse -> fine
    float* T;                  // H x max(w) intermediate of the separable blur+resize
    float* M;                  // variant 1: 5 planes at the finest level
    double* VS;                // variant 1: vertical sums, 5 planes (double, like cv2's vsum)
    // tf_farneback_step: the new frame's pyramid + expansion run on an auxiliary stream, level by level
    // (coarse -> fine), while the caller's stream already solves the coarser levels
    cudaStream_t aux;
    cudaEvent_t ev_start;
    std::vector<cudaEvent_t> ev_level[FB_SLOTS];  // per slot and level: that level's R of the slot's frame is built
    cudaEvent_t ev_read[FB_SLOTS][FB_LANES];      // per slot and lane: the lane's last solve that read the slot is done
    bool has_frame[FB_SLOTS];                     // a frame was prepared into the slot
};

// ---------------------------------------------------------------------------------------------
// prepare kernels
// ---------------------------------------------------------------------------------------------
// Horizontal Gaussian (BORDER_REFLECT_101) evaluated only at the two source columns each output
// column interpolates between, then the horizontal lerp: gray u8 (H, W) -> T f32 (H, w).
__global__ void __launch_bounds__(256) k_fb_hpass(const uint8_t* __restrict__ gray, float* __restrict__ T,
                                                  const float* __restrict__ gk, const int* __restrict__ sx,
                                                  const float* __restrict__ tx, int H, int W, int w, int ksz) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (x >= w) return;
    const uint8_t* row = gray + (size_t)r * W;
    int s0 = sx[x];
    float t = tx[x];
    int rad = ksz >> 1;
    float a = 0.f, b = 0.f;
    if (s0 - rad >= 0 && s0 + 1 + rad < W) {
        for (int j = 0; j < ksz; j++) {
            float g = gk[j];
            a = fmaf(g, (float)row[s0 + j - rad], a);
            b = fmaf(g, (float)row[s0 + 1 + j - rad], b);
        }
    } else {
        int s1 = min(s0 + 1, W - 1);
        for (int j = 0; j < ksz; j++) {
            float g = gk[j];
            a = fmaf(g, (float)row[reflect101(s0 + j - rad, W)], a);
            b = fmaf(g, (float)row[reflect101(s1 + j - rad, W)], b);
        }
    }
    T[(size_t)r * w + x] = a * (1.f - t) + b * t;
}

// Vertical Gaussian at the two source rows + vertical lerp: T (H, w) -> img (h, w).
__global__ void __launch_bounds__(256) k_fb_vpass(const float* __restrict__ T, float* __restrict__ img,
                                                  const float* __restrict__ gk, const int* __restrict__ sy,
                                                  const float* __restrict__ ty, int H, int w, int h, int ksz) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int s0 = sy[y];
    int s1 = min(s0 + 1, H - 1);
    float t = ty[y];
    int rad = ksz >> 1;
    float a = 0.f, b = 0.f;
    for (int j = 0; j < ksz; j++) {
        float g = gk[j];
        a = fmaf(g, T[(size_t)reflect101(s0 + j - rad, H) * w + x], a);
        b = fmaf(g, T[(size_t)reflect101(s1 + j - rad, H) * w + x], b);
    }
    img[(size_t)y * w + x] = a * (1.f - t) + b * t;
}

// hpass + vpass fused for one pyramid level: a block produces BR_TX x BR_TY level pixels from the gray
// footprint they depend on (staged in shared memory as bytes, BORDER_REFLECT_101 applied while loading),
// with the SAME operation order as the two separate passes, so the result is bit-identical to them.
#define BR_TX 32
#define BR_TY 8
#define BR_NT 256  // threads per block; tiles are BR_TX x BR_TY, or larger (a level picks the largest that stays small in
                   // shared memory: a tile is a short chain of dependent memory round trips, bigger tiles amortise it)
#define FB_MAX_FUSED_LEVELS 8
struct BrLevel {
    const float* gk;
    const int* sx;
    const float* tx;
    const int* sy;
    const float* ty;
    float* img;
    int w, h, ksz, fr_max, fc_pitch, tiles_x, tile_end;  // tile_end: running total of tiles up to this level
    int tw, th;                                          // tile size in level pixels (tw a power of two <= BR_NT)
};
struct BrArgs {
    BrLevel lv[FB_MAX_FUSED_LEVELS];
    int n;
};
// All pyramid levels that need a blur go in ONE launch (coarsest = widest Gaussian first): every tile is a short
// chain of dependent memory round trips, and the levels' chains overlap instead of running back to back.
__global__ void __launch_bounds__(BR_NT) k_fb_blur_resize(const uint8_t* __restrict__ gray, int H, int W,
                                                                  const __grid_constant__ BrArgs args) {
    extern __shared__ __align__(16) unsigned char br_smem[];
    int li = 0;
    while (li + 1 < args.n && (int)blockIdx.x >= args.lv[li].tile_end) li++;
    const BrLevel& A = args.lv[li];
    const int tile = blockIdx.x - (li ? args.lv[li - 1].tile_end : 0);
    const float* __restrict__ gk = A.gk;
    const int* __restrict__ sx = A.sx;
    const float* __restrict__ tx = A.tx;
    const int* __restrict__ sy = A.sy;
    const float* __restrict__ ty = A.ty;
    float* __restrict__ img = A.img;
    const int w = A.w, h = A.h, ksz = A.ksz, fr_max = A.fr_max, fc_pitch = A.fc_pitch;
    const int tw = A.tw, th = A.th;
    float* sT = reinterpret_cast<float*>(br_smem);                   // [fr_max][tw]
    float* sG = sT + fr_max * tw;                                    // [ksz]
    unsigned char* sS = reinterpret_cast<unsigned char*>(sG + ksz);  // [fr_max][fc_pitch]
    const int tid = threadIdx.x;
    const int x0 = (tile % A.tiles_x) * tw, y0 = (tile / A.tiles_x) * th;
    const int xl = min(x0 + tw - 1, w - 1), yl = min(y0 + th - 1, h - 1);
    const int rad = ksz >> 1;
    const int c0 = sx[x0] - rad, r0 = sy[y0] - rad;
    const int ncols = sx[xl] + 1 + rad - c0 + 1, nrows = sy[yl] + 1 + rad - r0 + 1;
    for (int i = tid; i < ksz; i += BR_NT) sG[i] = gk[i];
    int off = 0;  // column of the footprint's first pixel inside its staged row
    if (c0 >= 0 && ((c0 & ~3) + ((((c0 & 3) + ncols + 3) >> 2) << 2)) <= W && (W & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(gray) & 3) == 0) {
        // tiles inside the frame: aligned 32-bit loads, four independent requests per thread in flight
        off = c0 & 3;
        const int nw = (off + ncols + 3) >> 2, total = nrows * nw;
        constexpr int NTH = BR_NT;
        for (int base = tid; base < total; base += 4 * NTH) {
            uint32_t v[4];
            int fr[4], q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int i = base + k * NTH;
                fr[k] = i / nw;
                q[k] = i - fr[k] * nw;
                if (i < total)
                    v[k] = __ldg(reinterpret_cast<const uint32_t*>(gray + (size_t)reflect101(r0 + fr[k], H) * W + (c0 - off)) + q[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (base + k * NTH < total) reinterpret_cast<uint32_t*>(sS + fr[k] * fc_pitch)[q[k]] = v[k];
        }
    } else {
        // tiles touching a border: one warp per footprint row, byte loads with reflected columns
        const int lane = tid & 31;
        for (int fr = tid >> 5; fr < nrows; fr += BR_NT / 32) {
            const uint8_t* src = gray + (size_t)reflect101(r0 + fr, H) * W;
            unsigned char* dst = sS + fr * fc_pitch;
            for (int fc = lane; fc < ncols; fc += 32) dst[fc] = __ldg(src + reflect101(c0 + fc, W));
        }
    }
    __syncthreads();
    // horizontal Gaussian at the two source columns of every output column + horizontal lerp
    {
        const int ox = tid & (tw - 1);
        const int x = min(x0 + ox, w - 1);
        const int lc = sx[x] - sx[x0];
        const float t = tx[x];
        for (int fr = tid / tw; fr < nrows; fr += BR_NT / tw) {
            const unsigned char* row = sS + fr * fc_pitch + off + lc;
            float a = 0.f, b = 0.f;
            float v = (float)row[0];
            for (int j = 0; j < ksz; j++) {
                float g = sG[j];
                float nv = (float)row[j + 1];
                a = fmaf(g, v, a);
                b = fmaf(g, nv, b);
                v = nv;
            }
            sT[fr * tw + ox] = a * (1.f - t) + b * t;
        }
    }
    __syncthreads();
    // vertical Gaussian at the two source rows + vertical lerp
    {
        const int ox = tid & (tw - 1);
        const int x = x0 + ox;
        for (int oy = tid / tw; oy < th; oy += BR_NT / tw) {
            const int y = y0 + oy;
            if (x < w && y < h) {
                const int lr = sy[y] - sy[y0];
                const float t = ty[y];
                const float* col = sT + lr * tw + ox;
                float a = 0.f, b = 0.f;
                float v = col[0];
                for (int j = 0; j < ksz; j++) {
                    float g = sG[j];
                    float nv = col[(j + 1) * tw];
                    a = fmaf(g, v, a);
                    b = fmaf(g, nv, b);
                    v = nv;
                }
                img[(size_t)y * w + x] = a * (1.f - t) + b * t;
            }
        }
    }
}

// Polynomial expansion (FarnebackPolyExp): separable (2n+1)^2 window, replicate borders.
// Tile 64 x 16 outputs, 256 threads; every thread produces 4 consecutive outputs in each pass
// from a register window, so shared-memory reads per output drop from (2n+1) to (2n+4)/4.
// The tile's first pixel is subtracted before accumulation: the five stored coefficients are
// invariant to a constant offset (the fit of a constant has zero derivative terms), and the
// smaller magnitudes keep fp32 accumulation at the accuracy of cv2's double accumulators.
// FUSE3: the pyramid image of a level with identity resize and a 3-tap Gaussian (the finest level)
// is computed on the fly from the gray frame while the tile is loaded (row filter then column
// filter, BORDER_REFLECT_101, same rounding order as the separate passes) and also written out.
__device__ __forceinline__ float fb_blur3(const uint8_t* __restrict__ gray, int x, int y, int w, int h, float k0,
                                          float k1) {
    int xl = x > 0 ? x - 1 : (w > 1 ? 1 : 0), xr = x < w - 1 ? x + 1 : (w > 1 ? w - 2 : 0);
    int yu = y > 0 ? y - 1 : (h > 1 ? 1 : 0), yd = y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0);
    const uint8_t* r0 = gray + (size_t)yu * w;
    const uint8_t* r1 = gray + (size_t)y * w;
    const uint8_t* r2 = gray + (size_t)yd * w;
    float h0 = fmaf(k0, (float)r0[xr], fmaf(k1, (float)r0[x], k0 * (float)r0[xl]));
    float h1 = fmaf(k0, (float)r1[xr], fmaf(k1, (float)r1[x], k0 * (float)r1[xl]));
    float h2 = fmaf(k0, (float)r2[xr], fmaf(k1, (float)r2[x], k0 * (float)r2[xl]));
    return fmaf(k0, h2, fmaf(k1, h1, k0 * h0));
}

template <int N, typename RT, bool FUSE3>
__global__ void __launch_bounds__(256, 5) k_fb_polyexp(const float* __restrict__ img, const uint8_t* __restrict__ gray,
                                                    float* __restrict__ img_out, RT* __restrict__ R, int w, int h,
                                                    PolyCoef pc, float k0, float k1) {
    constexpr int TX = 64, TY = 16, SW = TX + 2 * N, SH = TY + 2 * N, SP = SW + 1;
    __shared__ float sI[SH * SP];
    __shared__ float sV[3][TY * SP];
    int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    int tid = threadIdx.x;
    // FUSE3 fast path for tiles whose footprint lies inside the frame: the gray tile is staged in shared
    // memory as float with aligned 32-bit loads (4 pixels each, one conversion per byte), and every thread
    // blurs 4 neighbouring pixels from an 18-value register window.  With u8 inputs and taps 1/4, 1/2, 1/4
    // every product and partial sum is exact in fp32, so the evaluation order cannot change the result.
    constexpr int GW = 20;             // staged words per row: columns x0 - 8 .. x0 + 71
    constexpr int GP = 4 * GW + 1;     // pitch
    constexpr int GOFF = 8 - (N + 1);  // staged column of tile-local column lx = 0
    constexpr int NQ = (SW + 3) / 4;   // 4-pixel groups per row
    const bool interior = FUSE3 && N <= 7 && x0 >= 8 && y0 - N - 1 >= 0 && x0 + TX + 8 <= w &&
                          y0 + TY + N + 1 <= h && (w & 3) == 0 && (reinterpret_cast<uintptr_t>(gray) & 3) == 0;
    if (interior) {
        static_assert(N > 7 || (SH + 2) * GP <= 3 * TY * SP, "staged gray tile must fit in the sV scratch");
        float* sG = &sV[0][0];
        {   // all of a thread's words requested before the first is unpacked (one wait for memory, not one per word)
            constexpr int NW = ((SH + 2) * GW + 255) / 256;
            uint32_t v[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
                int i = tid + 256 * k, r = i / GW, q = i - r * GW;
                if (i < (SH + 2) * GW)
                    v[k] = __ldg(reinterpret_cast<const uint32_t*>(gray + (size_t)(y0 - N - 1 + r) * w + (x0 - 8)) + q);
            }
#pragma unroll
            for (int k = 0; k < NW; k++) {
                int i = tid + 256 * k, r = i / GW, q = i - r * GW;
                if (i < (SH + 2) * GW) {
                    float* d = sG + r * GP + 4 * q;
                    d[0] = (float)(v[k] & 255u);
                    d[1] = (float)((v[k] >> 8) & 255u);
                    d[2] = (float)((v[k] >> 16) & 255u);
                    d[3] = (float)(v[k] >> 24);
                }
            }
        }
        __syncthreads();
        float c0i;
        {
            const float* g = sG + N * GP + N + GOFF;
            float h0 = fmaf(k0, g[2], fmaf(k1, g[1], k0 * g[0]));
            float h1 = fmaf(k0, g[GP + 2], fmaf(k1, g[GP + 1], k0 * g[GP]));
            float h2 = fmaf(k0, g[2 * GP + 2], fmaf(k1, g[2 * GP + 1], k0 * g[2 * GP]));
            c0i = fmaf(k0, h2, fmaf(k1, h1, k0 * h0));
        }
        for (int i = tid; i < SH * NQ; i += 256) {
            int ly = i / NQ, lx = (i - ly * NQ) * 4;
            const float* g = sG + ly * GP + lx + GOFF;
            float hz[3][4];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                float t[6];
#pragma unroll
                for (int j = 0; j < 6; j++) t[j] = g[r * GP + j];
#pragma unroll
                for (int o = 0; o < 4; o++) hz[r][o] = fmaf(k0, t[o + 2], fmaf(k1, t[o + 1], k0 * t[o]));
            }
            const bool row_in = ly >= N && ly < N + TY;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                if (lx + o < SW) {
                    float v = fmaf(k0, hz[2][o], fmaf(k1, hz[1][o], k0 * hz[0][o]));
                    if (img_out && row_in && lx + o >= N && lx + o < N + TX)
                        img_out[(size_t)(y0 + ly - N) * w + (x0 + lx + o - N)] = v;
                    sI[ly * SP + lx + o] = v - c0i;
                }
            }
        }
    }
    float c0 = interior ? 0.f
               : FUSE3  ? fb_blur3(gray, min(x0, w - 1), min(y0, h - 1), w, h, k0, k1)
                        : __ldg(img + (size_t)min(y0, h - 1) * w + min(x0, w - 1));
    if (FUSE3) {
        for (int i = tid; i < (interior ? 0 : SH * SW); i += 256) {
            int ly = i / SW, lx = i - ly * SW;
            int gy = clampi(y0 + ly - N, 0, h - 1), gx = clampi(x0 + lx - N, 0, w - 1);
            float v = fb_blur3(gray, gx, gy, w, h, k0, k1);
            // interior of the tile: publish the pyramid image (debug hook / other consumers)
            if (img_out && ly >= N && ly < N + TY && lx >= N && lx < N + TX && y0 + ly - N < h && x0 + lx - N < w)
                img_out[(size_t)gy * w + gx] = v;
            sI[ly * SP + lx] = v - c0;
        }
    } else {
        // the footprint in batches of BATCH loads per thread, all requested before the first is used: a thread waits
        // for memory once per batch instead of once per element (the loop is 7.5 elements long for n = 5)
        constexpr int BATCH = 4;
        for (int i0 = tid; i0 < SH * SW; i0 += 256 * BATCH) {
            float v[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; k++) {
                int i = i0 + 256 * k;
                int ly = i / SW, lx = i - ly * SW;
                int gy = clampi(y0 + ly - N, 0, h - 1), gx = clampi(x0 + lx - N, 0, w - 1);
                if (i < SH * SW) v[k] = __ldg(img + (size_t)gy * w + gx);
            }
#pragma unroll
            for (int k = 0; k < BATCH; k++) {
                int i = i0 + 256 * k;
                int ly = i / SW, lx = i - ly * SW;
                if (i < SH * SW) sI[ly * SP + lx] = v[k] - c0;
            }
        }
    }
    __syncthreads();
    // vertical pass: SW columns x (TY / 4) row groups
    for (int i = tid; i < SW * (TY / 4); i += 256) {
        int lx = i % SW, gy = (i / SW) * 4;
        float win[4 + 2 * N];
#pragma unroll
        for (int j = 0; j < 4 + 2 * N; j++) win[j] = sI[(gy + j) * SP + lx];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            float r0 = win[o + N] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; k++) {
                float up = win[o + N - k], dn = win[o + N + k];
                float p = up + dn;
                r0 = fmaf(pc.g[k], p, r0);
                r1 = fmaf(pc.xg[k], dn - up, r1);
                r2 = fmaf(pc.xxg[k], p, r2);
            }
            sV[0][(gy + o) * SP + lx] = r0;
            sV[1][(gy + o) * SP + lx] = r1;
            sV[2][(gy + o) * SP + lx] = r2;
        }
    }
    __syncthreads();
    // horizontal pass: TY rows x (TX / 4) column groups = 256 work items
    {
        // a warp covers 8 column groups x 4 rows: with the odd pitch SP the 32 lanes read 32
        // distinct banks (16 groups x 2 rows would collide pairwise)
        static_assert(TX == 64 && TY == 16 && (SP & 1) == 1, "lane mapping of the horizontal pass");
        const int lane = tid & 31, wrp = tid >> 5;
        int ly = (lane >> 3) + 4 * (wrp >> 1), gx = ((lane & 7) + 8 * (wrp & 1)) * 4;
        float w0[4 + 2 * N], w1[4 + 2 * N], w2[4 + 2 * N];
#pragma unroll
        for (int j = 0; j < 4 + 2 * N; j++) {
            w0[j] = sV[0][ly * SP + gx + j];
            w1[j] = sV[1][ly * SP + gx + j];
            w2[j] = sV[2][ly * SP + gx + j];
        }
        int y = y0 + ly;
        size_t plane = (size_t)w * h;
        float out[5][4];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            float b1 = w0[o + N] * pc.g[0], b2 = 0.f, b3 = w1[o + N] * pc.g[0], b4 = 0.f, b5 = w2[o + N] * pc.g[0],
                  b6 = 0.f;
#pragma unroll
            for (int k = 1; k <= N; k++) {
                float lo0 = w0[o + N - k], hi0 = w0[o + N + k];
                float lo1 = w1[o + N - k], hi1 = w1[o + N + k];
                float lo2 = w2[o + N - k], hi2 = w2[o + N + k];
                float tg = hi0 + lo0;
                b1 = fmaf(tg, pc.g[k], b1);
                b4 = fmaf(tg, pc.xxg[k], b4);
                b2 = fmaf(hi0 - lo0, pc.xg[k], b2);
                b3 = fmaf(hi1 + lo1, pc.g[k], b3);
                b6 = fmaf(hi1 - lo1, pc.xg[k], b6);
                b5 = fmaf(hi2 + lo2, pc.g[k], b5);
            }
            out[0][o] = b3 * pc.ig11;                        // d/dy
            out[1][o] = b2 * pc.ig11;                        // d/dx
            out[2][o] = fmaf(b1, pc.ig03, b5 * pc.ig33);     // yy
            out[3][o] = fmaf(b1, pc.ig03, b4 * pc.ig33);     // xx
            out[4][o] = b6 * pc.ig55;                        // xy
        }
        int x = x0 + gx;
        if (y < h) {
            size_t at = (size_t)y * w + x;
            // "4+1" layout: one 128-bit store per pixel for the first four coefficients
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (x + o < w) store_quad(R, at + o, make_float4(out[0][o], out[1][o], out[2][o], out[3][o]));
            if (sizeof(RT) == 4 && x + 4 <= w && (w & 3) == 0) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(R) + 4 * plane + at) =
                    make_float4(out[4][0], out[4][1], out[4][2], out[4][3]);
            } else {
#pragma unroll
                for (int o = 0; o < 4; o++)
                    if (x + o < w) store_c4(R, plane, at + o, out[4][o]);
            }
        }
    }
}

// ---- polynomial expansion of the finest level with the 3x3 Gaussian FOLDED into the expansion filters ----------
// R = PolyExp(Blur3(gray)) and every operator is linear and separable, so on tiles whose footprint lies inside the
// frame the expansion taps are convolved with [1/4, 1/2, 1/4] once on the host (radius n + 1; symmetric taps stay
// symmetric, the antisymmetric x*g taps stay antisymmetric) and the blurred image is never formed: no blur phase,
// no second staging buffer.  Tiles touching the frame border keep the two-stage evaluation (reflected blur, then
// replicated expansion borders).  Rounding differs from the two-stage form in the last bits only (1e-7 relative).
template <int M, int SP, bool CENTER2>
__device__ __forceinline__ void pe_vertical(const float* __restrict__ sI, float* __restrict__ sV, const PolyCoef& pc,
                                            int tid) {
    constexpr int TX = 64, TY = 16, SW = TX + 2 * M;
    for (int i = tid; i < SW * (TY / 4); i += 256) {
        int lx = i % SW, gy = (i / SW) * 4;
        float win[4 + 2 * M];
#pragma unroll
        for (int j = 0; j < 4 + 2 * M; j++) win[j] = sI[(gy + j) * SP + lx];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            float r0 = win[o + M] * pc.g[0], r1 = 0.f, r2 = CENTER2 ? win[o + M] * pc.xxg[0] : 0.f;
#pragma unroll
            for (int k = 1; k <= M; k++) {
                float up = win[o + M - k], dn = win[o + M + k];
                float p = up + dn;
                r0 = fmaf(pc.g[k], p, r0);
                r1 = fmaf(pc.xg[k], dn - up, r1);
                r2 = fmaf(pc.xxg[k], p, r2);
            }
            sV[(gy + o) * SP + lx] = r0;
            sV[TY * SP + (gy + o) * SP + lx] = r1;
            sV[2 * TY * SP + (gy + o) * SP + lx] = r2;
        }
    }
}

template <int M, int SP, bool CENTER2, typename RT>
__device__ __forceinline__ void pe_horizontal(const float* __restrict__ sV, RT* __restrict__ R, const PolyCoef& pc, int tid,
                                              int x0, int y0, int w, int h) {
    constexpr int TX = 64, TY = 16;
    // a warp covers 8 column groups x 4 rows: with the odd pitch the 32 lanes read 32 distinct banks
    const int lane = tid & 31, wrp = tid >> 5;
    const int ly = (lane >> 3) + 4 * (wrp >> 1), gx = ((lane & 7) + 8 * (wrp & 1)) * 4;
    float w0[4 + 2 * M], w1[4 + 2 * M], w2[4 + 2 * M];
#pragma unroll
    for (int j = 0; j < 4 + 2 * M; j++) {
        w0[j] = sV[ly * SP + gx + j];
        w1[j] = sV[TY * SP + ly * SP + gx + j];
        w2[j] = sV[2 * TY * SP + ly * SP + gx + j];
    }
    const int y = y0 + ly;
    const size_t plane = (size_t)w * h;
    float out[5][4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        float b1 = w0[o + M] * pc.g[0], b2 = 0.f, b3 = w1[o + M] * pc.g[0], b5 = w2[o + M] * pc.g[0], b6 = 0.f;
        float b4 = CENTER2 ? w0[o + M] * pc.xxg[0] : 0.f;
#pragma unroll
        for (int k = 1; k <= M; k++) {
            float lo0 = w0[o + M - k], hi0 = w0[o + M + k];
            float lo1 = w1[o + M - k], hi1 = w1[o + M + k];
            float lo2 = w2[o + M - k], hi2 = w2[o + M + k];
            float tg = hi0 + lo0;
            b1 = fmaf(tg, pc.g[k], b1);
            b4 = fmaf(tg, pc.xxg[k], b4);
            b2 = fmaf(hi0 - lo0, pc.xg[k], b2);
            b3 = fmaf(hi1 + lo1, pc.g[k], b3);
            b6 = fmaf(hi1 - lo1, pc.xg[k], b6);
            b5 = fmaf(hi2 + lo2, pc.g[k], b5);
        }
        out[0][o] = b3 * pc.ig11;                     // d/dy
        out[1][o] = b2 * pc.ig11;                     // d/dx
        out[2][o] = fmaf(b1, pc.ig03, b5 * pc.ig33);  // yy
        out[3][o] = fmaf(b1, pc.ig03, b4 * pc.ig33);  // xx
        out[4][o] = b6 * pc.ig55;                     // xy
    }
    const int x = x0 + gx;
    if (y < h) {
        size_t at = (size_t)y * w + x;
#pragma unroll
        for (int o = 0; o < 4; o++)
            if (x + o < w) store_quad(R, at + o, make_float4(out[0][o], out[1][o], out[2][o], out[3][o]));
        if (sizeof(RT) == 4 && x + 4 <= w && (w & 3) == 0) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(R) + 4 * plane + at) =
                make_float4(out[4][0], out[4][1], out[4][2], out[4][3]);
        } else {
#pragma unroll
            for (int o = 0; o < 4; o++)
                if (x + o < w) store_c4(R, plane, at + o, out[4][o]);
        }
    }
}

// pc: expansion taps (radius N); pcf: the same taps convolved with [k0, k1, k0] (radius N + 1)
template <int N, typename RT>
__global__ void __launch_bounds__(256, 5) k_fb_polyexp_folded(const uint8_t* __restrict__ gray, RT* __restrict__ R, int w,
                                                              int h, const PolyCoef pc, const PolyCoef pcf, float k0,
                                                              float k1) {
    constexpr int TX = 64, TY = 16, MF = N + 1, SP = TX + 2 * MF + 1;
    __shared__ float sI[(TY + 2 * MF) * SP];
    __shared__ float sV[3 * TY * SP];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    const bool interior = MF <= 8 && x0 >= 8 && y0 - MF >= 0 && x0 + TX + 8 <= w && y0 + TY + MF <= h && (w & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(gray) & 3) == 0;
    if (interior) {
        // gray footprint rows y0 - MF .. y0 + TY + MF - 1, columns x0 - MF .. x0 + TX + MF - 1, through aligned
        // 32-bit loads of columns x0 - 8 .. x0 + TX + 7; the tile's first pixel is subtracted (see k_fb_polyexp)
        constexpr int SH = TY + 2 * MF, GW = (TX + 16) / 4, SKIP = 8 - MF;
        const float c0 = (float)__ldg(gray + (size_t)y0 * w + x0);
        {   // all of a thread's words requested before the first is unpacked (one wait for memory, not one per word)
            constexpr int NW = (SH * GW + 255) / 256;
            uint32_t v[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
                int i = tid + 256 * k, r = i / GW, q = i - r * GW;
                if (i < SH * GW)
                    v[k] = __ldg(reinterpret_cast<const uint32_t*>(gray + (size_t)(y0 - MF + r) * w + (x0 - 8)) + q);
            }
#pragma unroll
            for (int k = 0; k < NW; k++) {
                int i = tid + 256 * k, r = i / GW, q = i - r * GW;
                if (i < SH * GW) {
                    float* d = sI + r * SP + 4 * q - SKIP;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int lx = 4 * q + j - SKIP;
                        if (lx >= 0 && lx < TX + 2 * MF) d[j] = (float)((v[k] >> (8 * j)) & 255u) - c0;
                    }
                }
            }
        }
        __syncthreads();
        pe_vertical<MF, SP, true>(sI, sV, pcf, tid);
        __syncthreads();
        pe_horizontal<MF, SP, true, RT>(sV, R, pcf, tid, x0, y0, w, h);
    } else {
        constexpr int SW = TX + 2 * N, SH = TY + 2 * N;
        const float c0 = fb_blur3(gray, min(x0, w - 1), min(y0, h - 1), w, h, k0, k1);
        for (int i = tid; i < SH * SW; i += 256) {
            int ly = i / SW, lx = i - ly * SW;
            int gy = clampi(y0 + ly - N, 0, h - 1), gx = clampi(x0 + lx - N, 0, w - 1);
            sI[ly * SP + lx] = fb_blur3(gray, gx, gy, w, h, k0, k1) - c0;
        }
        __syncthreads();
        pe_vertical<N, SP, false>(sI, sV, pc, tid);
        __syncthreads();
        pe_horizontal<N, SP, false, RT>(sV, R, pc, tid, x0, y0, w, h);
    }
}

// ---------------------------------------------------------------------------------------------
// solve kernels (variant 1: materialised M / vertical sums)
// ---------------------------------------------------------------------------------------------
// resize(prevFlow, INTER_LINEAR) * (1 / pyr_scale)
__device__ __forceinline__ float2 fb_upsample_one(const float2* __restrict__ r0, const float2* __restrict__ r1, int x0,
                                                  int sw, float a, float b, float mul) {
    int x1 = min(x0 + 1, sw - 1);
    float2 p00 = __ldg(r0 + x0), p01 = __ldg(r0 + x1), p10 = __ldg(r1 + x0), p11 = __ldg(r1 + x1);
    float r0x = p00.x * (1.f - a) + p01.x * a, r0y = p00.y * (1.f - a) + p01.y * a;
    float r1x = p10.x * (1.f - a) + p11.x * a, r1y = p10.y * (1.f - a) + p11.y * a;
    return make_float2((r0x * (1.f - b) + r1x * b) * mul, (r0y * (1.f - b) + r1y * b) * mul);
}

// two horizontally adjacent outputs per thread: one 128-bit store, tables read as 64-bit pairs
__global__ void __launch_bounds__(256) k_fb_upsample_flow(const float2* __restrict__ src, float2* __restrict__ dst,
                                                          const int* __restrict__ sx, const float* __restrict__ tx,
                                                          const int* __restrict__ sy, const float* __restrict__ ty,
                                                          int sw, int sh, int w, int h, float mul) {
    int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    int y = blockIdx.y;
    if (x >= w) return;
    int y0 = sy[y];
    int y1 = min(y0 + 1, sh - 1);
    float b = ty[y];
    const float2* r0 = src + (size_t)y0 * sw;
    const float2* r1 = src + (size_t)y1 * sw;
    float2* out = dst + (size_t)y * w + x;
    if (x + 1 < w && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        int2 xs = __ldg(reinterpret_cast<const int2*>(sx + x));
        float2 ts = __ldg(reinterpret_cast<const float2*>(tx + x));
        float2 u = fb_upsample_one(r0, r1, xs.x, sw, ts.x, b, mul);
        float2 v = fb_upsample_one(r0, r1, xs.y, sw, ts.y, b, mul);
        *reinterpret_cast<float4*>(out) = make_float4(u.x, u.y, v.x, v.y);
    } else {
        out[0] = fb_upsample_one(r0, r1, sx[x], sw, tx[x], b, mul);
        if (x + 1 < w) out[1] = fb_upsample_one(r0, r1, sx[x + 1], sw, tx[x + 1], b, mul);
    }
}

template <typename RT>
__global__ void __launch_bounds__(256) k_fb_update_matrices(const RT* __restrict__ R0, const RT* __restrict__ R1,
                                                            const float2* __restrict__ flow, float* __restrict__ M,
                                                            int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t plane = (size_t)w * h, at = (size_t)y * w + x;
    float2 f = flow ? flow[at] : make_float2(0.f, 0.f);
    float m[5];
    fb_update_matrix<RT>(R0, R1, plane, w, h, x, y, f, m);
#pragma unroll
    for (int c = 0; c < 5; c++) M[c * plane + at] = m[c];
}

__global__ void __launch_bounds__(256) k_fb_box_v(const float* __restrict__ M, double* __restrict__ VS, int w, int h,
                                                  int m) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t plane = (size_t)w * h;
    double acc[5] = {0., 0., 0., 0., 0.};
    for (int k = -m; k <= m; k++) {
        size_t at = (size_t)clampi(y + k, 0, h - 1) * w + x;
#pragma unroll
        for (int c = 0; c < 5; c++) acc[c] += (double)M[c * plane + at];
    }
#pragma unroll
    for (int c = 0; c < 5; c++) VS[c * plane + (size_t)y * w + x] = acc[c];
}

__global__ void __launch_bounds__(256) k_fb_box_h_solve(const double* __restrict__ VS, float2* __restrict__ flow, int w,
                                                        int h, int m, double scale, int clip) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t plane = (size_t)w * h;
    double acc[5] = {0., 0., 0., 0., 0.};
    for (int k = -m; k <= m; k++) {
        size_t at = (size_t)y * w + clampi(x + k, 0, w - 1);
#pragma unroll
        for (int c = 0; c < 5; c++) acc[c] += VS[c * plane + at];
    }
    float2 f = fb_solve(acc, scale);
    if (clip) {
        f.x = fminf(fmaxf(f.x, (float)(-x)), (float)(w - 1 - x));
        f.y = fminf(fmaxf(f.y, (float)(-y)), (float)(h - 1 - y));
    }
    flow[(size_t)y * w + x] = f;
}

static int g_fb_two_pass = 0;  // tf_farneback_tune(1, ..): separate blur passes instead of the fused kernel

#include "fb_iter.cuh"
#include "fb_tile.cuh"
#include "fb_half.cuh"
#include "fb_stage.cuh"
#include "fb_tma.cuh"
#include "fb_ring.cuh"
#include "fb_pack.cuh"

int g_fbh_rows = 0;
int g_fbh_rows_min_px = 0;
int g_fbr_rows = 0;
static int g_fbr_min_px = 1000000;  // variant 8 uses the ring kernel on levels of at least this many pixels (key 4)
extern "C" int tf_farneback_tune(int key, int value) {
    if (key == 0) g_fbh_rows = value;
    else if (key == 1) g_fb_two_pass = value;
    else if (key == 2) g_fbh_rows_min_px = value;
    else if (key == 3) g_fbr_rows = value;
    else if (key == 4) g_fbr_min_px = value;
    else return fail(TF_ERR_INVALID_ARG, "tf_farneback_tune: unknown key %d", key);
    return TF_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
    TF_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), v.size() * sizeof(T)));
    TF_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TF_OK;
}

extern "C" int tf_farneback_destroy(tf_farneback* h) {
    if (!h) return TF_OK;
    for (auto& l : h->lv) {
        cudaFree(l.gk); cudaFree(l.sx); cudaFree(l.tx); cudaFree(l.sy); cudaFree(l.ty);
        cudaFree(l.img);
        for (int i = 0; i < FB_SLOTS; i++) cudaFree(l.R[i]);
        for (int i = 0; i < FB_LANES; i++) { cudaFree(l.flow[i]); cudaFree(l.flow2[i]); }
        cudaFree(l.fsx); cudaFree(l.ftx); cudaFree(l.fsy); cudaFree(l.fty);
    }
    cudaFree(h->T); cudaFree(h->M); cudaFree(h->VS);
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    for (int i = 0; i < FB_SLOTS; i++) {
        for (auto e : h->ev_level[i]) cudaEventDestroy(e);
        for (int l = 0; l < FB_LANES; l++)
            if (h->ev_read[i][l]) cudaEventDestroy(h->ev_read[i][l]);
    }
    delete h;
    return TF_OK;
}

extern "C" int tf_farneback_create(tf_farneback** out, int height, int width, double pyr_scale, int levels,
                                   int winsize, int iterations, int poly_n, double poly_sigma, int flags, int r_fp16) {
    TF_REQUIRE(out, TF_ERR_INVALID_ARG, "tf_farneback_create: null out");
    TF_REQUIRE(height >= 2 && width >= 2 && (size_t)height * width < (1u << 30), TF_ERR_SHAPE,
               "tf_farneback_create: bad shape %dx%d", height, width);
    TF_REQUIRE(pyr_scale > 0 && pyr_scale < 1, TF_ERR_INVALID_ARG, "fb_pyr_scale must be in (0, 1), got %g", pyr_scale);
    TF_REQUIRE(levels >= 0 && iterations >= 1, TF_ERR_INVALID_ARG, "fb_levels >= 0 and fb_iterations >= 1 required");
    TF_REQUIRE(poly_n >= 1 && poly_n <= FB_MAX_POLY_N, TF_ERR_INVALID_ARG, "fb_poly_n must be in [1, %d], got %d",
               FB_MAX_POLY_N, poly_n);
    TF_REQUIRE(winsize >= 3 && winsize / 2 <= FB_MAX_WIN_RADIUS, TF_ERR_INVALID_ARG,
               "fb_winsize must be in [3, %d], got %d (winsize < 3 hits an OpenCV running-sum quirk that is "
               "outside the parity scope)", 2 * FB_MAX_WIN_RADIUS + 1, winsize);
    TF_REQUIRE(flags == 0, TF_ERR_INVALID_ARG, "only fb_flags == 0 is supported (got %d)", flags);
    if (int e = require_sm100()) return e;
    tf_farneback* h = new (std::nothrow) tf_farneback();
    TF_REQUIRE(h, TF_ERR_CUDA, "out of host memory");
    h->H = height; h->W = width; h->pyr_scale = pyr_scale; h->levels_req = levels; h->winsize = winsize;
    h->iterations = iterations; h->poly_n = poly_n; h->poly_sigma = poly_sigma; h->flags = flags;
    h->r_fp16 = r_fp16 ? 1 : 0;
    h->publish_finest = 0;
    h->T = h->M = nullptr;
    h->VS = nullptr;
    h->aux = nullptr;
    h->ev_start = nullptr;
    memset(h->ev_read, 0, sizeof(h->ev_read));
    memset(h->has_frame, 0, sizeof(h->has_frame));
    prepare_poly(poly_n, poly_sigma, h->pc);
    {
        // fold [k0, k1, k0] = [1/4, 1/2, 1/4] into the taps: c[k] = k0 t[k-1] + k1 t[k] + k0 t[k+1], t even or odd in k
        const PolyCoef& a = h->pc;
        PolyCoef& f = h->pcf;
        f = a;
        f.n = poly_n + 1;
        auto even = [&](const float* t, int k) { int m = abs(k); return m <= poly_n ? (double)t[m] : 0.0; };
        auto odd = [&](const float* t, int k) { int m = abs(k); return m <= poly_n ? (k < 0 ? -(double)t[m] : (double)t[m]) : 0.0; };
        for (int k = 0; k <= poly_n + 1; k++) {
            f.g[k] = (float)(0.25 * even(a.g, k - 1) + 0.5 * even(a.g, k) + 0.25 * even(a.g, k + 1));
            f.xxg[k] = (float)(0.25 * even(a.xxg, k - 1) + 0.5 * even(a.xxg, k) + 0.25 * even(a.xxg, k + 1));
            f.xg[k] = (float)(0.25 * odd(a.xg, k - 1) + 0.5 * odd(a.xg, k) + 0.25 * odd(a.xg, k + 1));
        }
    }
    // level crop (min_size 32) exactly as optflowgf.cpp
    int k = 0;
    double scale = 1;
    for (; k < levels; k++) {
        scale *= pyr_scale;
        if (width * scale < 32 || height * scale < 32) break;
    }
    int maxw = 0;
    auto bail = [&](int e) { tf_farneback_destroy(h); return e; };
    for (int lvl = k; lvl >= 0; lvl--) {
        FbLevel L;
        memset(&L, 0, sizeof(L));
        scale = 1;
        for (int i = 0; i < lvl; i++) scale *= pyr_scale;
        L.k = lvl;
        L.sigma = (1. / scale - 1) * 0.5;
        L.ksz = std::max(cv_round(L.sigma * 5) | 1, 3);
        L.w = cv_round(width * scale);
        L.h = cv_round(height * scale);
        if (L.w < 2 || L.h < 2) return bail(fail(TF_ERR_SHAPE, "pyramid level %d is %dx%d", lvl, L.w, L.h));
        if (L.ksz / 2 >= std::min(width, height))
            return bail(fail(TF_ERR_SHAPE, "frame too small for the level-%d Gaussian (%d taps)", lvl, L.ksz));
        maxw = std::max(maxw, L.w);
        std::vector<int> s;
        std::vector<float> t;
        if (int e = upload(&L.gk, gaussian_kernel(L.ksz, L.sigma))) return bail(e);
        // tile of the fused blur+resize kernel: the largest of 64x16, 64x8, 32x8 level pixels whose footprint
        // (gray bytes + horizontal-pass floats) stays under 17 KB of shared memory; 32x8 otherwise
        std::vector<int> sxv, syv;
        std::vector<float> txv, tyv;
        linear_table(L.w, width, sxv, txv);
        linear_table(L.h, height, syv, tyv);
        const int cand[3][2] = {{64, 16}, {64, 8}, {BR_TX, BR_TY}};
        for (int ci = 0; ci < 3; ci++) {
            const int tw = cand[ci][0], th = cand[ci][1];
            int fc = 0, fr = 0;
            for (int a0 = 0; a0 < L.w; a0 += tw)
                fc = std::max(fc, sxv[std::min(a0 + tw - 1, L.w - 1)] - sxv[a0] + 2 * (L.ksz / 2) + 2);
            for (int a0 = 0; a0 < L.h; a0 += th)
                fr = std::max(fr, syv[std::min(a0 + th - 1, L.h - 1)] - syv[a0] + 2 * (L.ksz / 2) + 2);
            L.br_tw = tw;
            L.br_th = th;
            L.br_rows = fr;
            L.br_pitch = ((fc + 3) & ~3) + 4;  // + room for the alignment offset of the 32-bit loads
            L.br_smem = (size_t)fr * tw * 4 + (size_t)L.ksz * 4 + (size_t)fr * L.br_pitch;
            if (L.br_smem <= 17 * 1024) break;
        }
        if (L.br_smem > 96 * 1024) L.br_rows = 0;  // absurdly wide Gaussians: keep the two-pass kernels
        s = sxv; t = txv;
        if (int e = upload(&L.sx, s)) return bail(e);
        if (int e = upload(&L.tx, t)) return bail(e);
        s = syv; t = tyv;
        if (int e = upload(&L.sy, s)) return bail(e);
        if (int e = upload(&L.ty, t)) return bail(e);
        size_t n = (size_t)L.w * L.h;
        size_t rbytes = n * 5 * (h->r_fp16 ? 2 : 4);
        if (cudaMalloc(&L.img, n * 4) != cudaSuccess || cudaMalloc(&L.R[0], rbytes) != cudaSuccess ||
            cudaMalloc(&L.R[1], rbytes) != cudaSuccess || cudaMalloc(&L.flow[0], n * 8) != cudaSuccess ||
            cudaMalloc(&L.flow2[0], n * 8) != cudaSuccess)
            return bail(fail(TF_ERR_CUDA, "cudaMalloc failed for pyramid level %d (%dx%d)", lvl, L.w, L.h));
        if (!h->lv.empty()) {
            const FbLevel& C = h->lv.back();  // next-coarser level
            linear_table(L.w, C.w, s, t);
            if (int e = upload(&L.fsx, s)) return bail(e);
            if (int e = upload(&L.ftx, t)) return bail(e);
            linear_table(L.h, C.h, s, t);
            if (int e = upload(&L.fsy, s)) return bail(e);
            if (int e = upload(&L.fty, t)) return bail(e);
        }
        h->lv.push_back(L);
    }
    if (cudaMalloc(&h->T, (size_t)height * maxw * 4) != cudaSuccess)
        return bail(fail(TF_ERR_CUDA, "cudaMalloc failed for the blur intermediate"));
    *out = h;
    return TF_OK;
}

extern "C" int tf_farneback_set_debug(tf_farneback* h, int on) {
    TF_REQUIRE(h, TF_ERR_INVALID_ARG, "tf_farneback_set_debug: null handle");
    h->publish_finest = on != 0;
    return TF_OK;
}

extern "C" int tf_farneback_num_levels(const tf_farneback* h) { return h ? (int)h->lv.size() : 0; }

extern "C" int tf_farneback_level_size(const tf_farneback* h, int li, int* width, int* height) {
    TF_REQUIRE(h && li >= 0 && li < (int)h->lv.size(), TF_ERR_INVALID_ARG, "tf_farneback_level_size: bad level %d", li);
    if (width) *width = h->lv[li].w;
    if (height) *height = h->lv[li].h;
    return TF_OK;
}

template <typename RT>
static int launch_polyexp(const tf_farneback* h, const FbLevel& L, int slot, const uint8_t* gray_fused, cudaStream_t st) {
    dim3 grid(ceil_div(L.w, 64), ceil_div(L.h, 16));
    RT* R = reinterpret_cast<RT*>(L.R[slot]);
    ScopedKernelTimer timer(&L == &h->lv.back() ? TFK_FB_POLYEXP_FINEST : -1, st);
    switch (h->poly_n) {
#define TF_PE(N)                                                                                              \
    case N:                                                                                               \
        if (gray_fused && !h->publish_finest && !g_fb_two_pass)                                           \
            k_fb_polyexp_folded<N, RT><<<grid, 256, 0, st>>>(gray_fused, R, L.w, L.h, h->pc, h->pcf, 0.25f, 0.5f); \
        else if (gray_fused)                                                                              \
            k_fb_polyexp<N, RT, true><<<grid, 256, 0, st>>>(nullptr, gray_fused, h->publish_finest ? L.img : nullptr, R, \
                                                             L.w, L.h, h->pc, 0.25f, 0.5f);                      \
        else                                                                                              \
            k_fb_polyexp<N, RT, false><<<grid, 256, 0, st>>>(L.img, nullptr, nullptr, R, L.w, L.h, h->pc, 0.f, 0.f);  \
        break;
            TF_PE(1) TF_PE(2) TF_PE(3) TF_PE(4) TF_PE(5) TF_PE(6) TF_PE(7) TF_PE(8) TF_PE(9) TF_PE(10)
#undef TF_PE
        default: return fail(TF_ERR_INVALID_ARG, "unsupported poly_n %d", h->poly_n);
    }
    TF_LAUNCHED();
    return TF_OK;
}

// One launch for the blur + resize of every level that needs it; false if some level must use the two passes.
static bool fused_blur_ok(const tf_farneback* h) {
    if (g_fb_two_pass || h->lv.size() > FB_MAX_FUSED_LEVELS) return false;
    for (auto& L : h->lv)
        if (L.br_rows <= 0) return false;
    return true;
}

static int blur_levels(tf_farneback* h, const uint8_t* gray, cudaStream_t st) {
    BrArgs args;
    args.n = 0;
    size_t smem = 0;
    int tiles = 0;
    for (auto& L : h->lv) {  // coarse -> fine
        bool fuse = L.w == h->W && L.h == h->H && L.ksz == 3 && L.sigma <= 0;
        if (fuse) continue;  // the finest level's 3x3 blur lives in the polynomial expansion kernel
        BrLevel& B = args.lv[args.n++];
        B.gk = L.gk; B.sx = L.sx; B.tx = L.tx; B.sy = L.sy; B.ty = L.ty; B.img = L.img;
        B.w = L.w; B.h = L.h; B.ksz = L.ksz; B.fr_max = L.br_rows; B.fc_pitch = L.br_pitch;
        B.tw = L.br_tw; B.th = L.br_th;
        B.tiles_x = ceil_div(L.w, L.br_tw);
        tiles += B.tiles_x * ceil_div(L.h, L.br_th);
        B.tile_end = tiles;
        smem = std::max(smem, L.br_smem);
    }
    if (!args.n) return TF_OK;
    static bool attr_set = false;
    if (!attr_set) {
        TF_CUDA(cudaFuncSetAttribute(k_fb_blur_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    k_fb_blur_resize<<<tiles, BR_NT, smem, st>>>(gray, h->H, h->W, args);
    TF_LAUNCHED();
    return TF_OK;
}

static int prepare_level(tf_farneback* h, FbLevel& L, int slot, const uint8_t* gray, cudaStream_t st,
                         bool blurred = false) {
    // the finest level (identity resize, sigma 0 -> fixed [1/4, 1/2, 1/4] taps) is fused into polyexp
    bool fuse = L.w == h->W && L.h == h->H && L.ksz == 3 && L.sigma <= 0 && h->W >= 2 && h->H >= 2;
    if (!fuse && blurred) {
        // the batched blur launch (blur_levels) already produced L.img
    } else if (!fuse) {
        k_fb_hpass<<<dim3(ceil_div(L.w, 256), h->H), 256, 0, st>>>(gray, h->T, L.gk, L.sx, L.tx, h->H, h->W, L.w, L.ksz);
        TF_LAUNCHED();
        k_fb_vpass<<<dim3(ceil_div(L.w, 256), L.h), 256, 0, st>>>(h->T, L.img, L.gk, L.sy, L.ty, h->H, L.w, L.h, L.ksz);
        TF_LAUNCHED();
    }
    const uint8_t* g = fuse ? gray : nullptr;
    return h->r_fp16 ? launch_polyexp<__half>(h, L, slot, g, st) : launch_polyexp<float>(h, L, slot, g, st);
}

// slot 2 / lane 1 are allocated on first use (cudaMalloc synchronises the device once)
static int ensure_slot(tf_farneback* h, int slot) {
    for (auto& L : h->lv)
        if (!L.R[slot]) {
            size_t rbytes = (size_t)L.w * L.h * 5 * (h->r_fp16 ? 2 : 4);
            if (cudaMalloc(&L.R[slot], rbytes) != cudaSuccess)
                return fail(TF_ERR_CUDA, "cudaMalloc failed for slot %d of a %dx%d level", slot, L.w, L.h);
        }
    return TF_OK;
}

static int ensure_lane(tf_farneback* h, int lane) {
    for (auto& L : h->lv)
        if (!L.flow[lane]) {
            size_t n = (size_t)L.w * L.h;
            if (cudaMalloc(&L.flow[lane], n * 8) != cudaSuccess || cudaMalloc(&L.flow2[lane], n * 8) != cudaSuccess)
                return fail(TF_ERR_CUDA, "cudaMalloc failed for lane %d of a %dx%d level", lane, L.w, L.h);
        }
    return TF_OK;
}

static bool slot_ok(int s) { return s >= 0 && s < FB_SLOTS; }

extern "C" int tf_farneback_reserve(tf_farneback* h, int slots, int lanes) {
    TF_REQUIRE(h, TF_ERR_INVALID_ARG, "tf_farneback_reserve: null handle");
    TF_REQUIRE(slots >= 0 && slots <= FB_SLOTS && lanes >= 0 && lanes <= FB_LANES, TF_ERR_INVALID_ARG,
               "tf_farneback_reserve: at most %d slots and %d lanes", FB_SLOTS, FB_LANES);
    for (int s = 0; s < slots; s++)
        if (int e = ensure_slot(h, s)) return e;
    for (int l = 0; l < lanes; l++)
        if (int e = ensure_lane(h, l)) return e;
    return TF_OK;
}

extern "C" int tf_farneback_prepare(tf_farneback* h, int slot, const uint8_t* gray, void* stream) {
    TF_REQUIRE(h && gray, TF_ERR_INVALID_ARG, "tf_farneback_prepare: null argument");
    TF_REQUIRE(slot_ok(slot), TF_ERR_INVALID_ARG, "tf_farneback_prepare: slot must be in [0, %d)", FB_SLOTS);
    if (int e = ensure_slot(h, slot)) return e;
    h->has_frame[slot] = true;
    cudaStream_t st = as_stream(stream);
    const bool batched = fused_blur_ok(h);
    if (batched)
        if (int e = blur_levels(h, gray, st)) return e;
    for (auto& L : h->lv)
        if (int e = prepare_level(h, L, slot, gray, st, batched)) return e;
    return TF_OK;
}

template <typename RT>
static int solve_impl(tf_farneback* h, int sl, int sr, float2* flow_out, int variant, int clip, cudaStream_t st,
                      bool wait_levels = false, int lane = 0) {
    int m = h->winsize / 2;
    double scale = 1.0 / ((double)h->winsize * h->winsize);
    if (variant == 1 && !h->M) {
        size_t n = (size_t)h->lv.back().w * h->lv.back().h;
        TF_CUDA(cudaMalloc(&h->M, n * 5 * sizeof(float)));
        TF_CUDA(cudaMalloc(&h->VS, n * 5 * sizeof(double)));
    }
    const FbLevel* prev = nullptr;
    for (size_t li = 0; li < h->lv.size(); li++) {
        FbLevel& L = h->lv[li];
        bool finest = li + 1 == h->lv.size();
        if (wait_levels) {  // this level's R of both frames (built on the auxiliary stream, maybe for the other lane)
            TF_CUDA(cudaStreamWaitEvent(st, h->ev_level[sl][li], 0));
            TF_CUDA(cudaStreamWaitEvent(st, h->ev_level[sr][li], 0));
        }
        float2* final_buf = finest ? flow_out : L.flow[lane];
        float2* other_buf = finest ? L.flow[lane] : L.flow2[lane];
        const RT* R0 = reinterpret_cast<const RT*>(L.R[sl]);
        const RT* R1 = reinterpret_cast<const RT*>(L.R[sr]);
        dim3 grid(ceil_div(L.w, 256), L.h);
        bool zero_init = prev == nullptr;
        // where the initial flow of this level goes: the buffer iteration 0 reads from
        float2* init_buf = (variant == 1) ? final_buf : (((h->iterations - 1) & 1) ? final_buf : other_buf);
        if (prev) {
            k_fb_upsample_flow<<<dim3(ceil_div(ceil_div(L.w, 2), 256), L.h), 256, 0, st>>>(prev->flow[lane], init_buf, L.fsx, L.ftx, L.fsy, L.fty, prev->w, prev->h,
                                                     L.w, L.h, (float)(1.0 / h->pyr_scale));
            TF_LAUNCHED();
        }
        if (variant == 1) {
            float2* flow = final_buf;
            {
                ScopedKernelTimer timer(finest ? TFK_FB_UM_FINEST : -1, st);
                k_fb_update_matrices<RT><<<grid, 256, 0, st>>>(R0, R1, zero_init ? nullptr : flow, h->M, L.w, L.h);
            }
            TF_LAUNCHED();
            for (int it = 0; it < h->iterations; it++) {
                bool last = it + 1 == h->iterations;
                {
                    ScopedKernelTimer timer(finest ? TFK_FB_BOXV_FINEST : -1, st);
                    k_fb_box_v<<<grid, 256, 0, st>>>(h->M, h->VS, L.w, L.h, m);
                }
                TF_LAUNCHED();
                {
                    ScopedKernelTimer timer(finest ? TFK_FB_BOXH_FINEST : -1, st);
                    k_fb_box_h_solve<<<grid, 256, 0, st>>>(h->VS, flow, L.w, L.h, m, scale, clip && finest && last);
                }
                TF_LAUNCHED();
                if (!last) {
                    ScopedKernelTimer timer(finest ? TFK_FB_UM_FINEST : -1, st);
                    k_fb_update_matrices<RT><<<grid, 256, 0, st>>>(R0, R1, flow, h->M, L.w, L.h);
                }
                if (!last) TF_LAUNCHED();
            }
        } else if (variant == 3) {
            if (int e = fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st))
                return e;
        } else if (variant >= 12 && variant <= 16) {
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_tma<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st,
                                             15 - variant)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant >= 9 && variant <= 11) {
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_stage<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st,
                                               variant - 9)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant >= 18 && variant <= 22) {
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_ring<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st,
                                              variant == 18 ? 256 : variant == 19 ? 320 : variant == 20 ? 384
                                              : variant == 21 ? -256 : -320)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant == 8) {
            // default: the half-buffer kernel with the 4-column phase C and the conflict-free phase B (variant 24) where a
            // level gives it enough CTAs (159 vs 185 us at 4K, 41 vs 52 us at 1080p, 15.1 vs 18.6 us at 960x540), the
            // rolling-tile kernel below
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_half<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, 24, st)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant == 25 || (variant >= 27 && variant <= 31)) {
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_pack<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, variant, st)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant >= 23 && variant <= 24) {
            const bool big = (size_t)L.w * L.h >= (size_t)400000;
            int e = big ? fb_iterate_half<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, variant, st)
                        : fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest, st);
            if (e) return e;
        } else if (variant >= 4) {
            if (int e = fb_iterate_half<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest, finest,
                                            variant, st))
                return e;
        } else {
            if (int e = fb_iterate_fused<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip && finest,
                                             variant == 2, finest, st))
                return e;
        }
        prev = &L;
    }
    return TF_OK;
}

extern "C" int tf_farneback_solve(tf_farneback* h, int slot_left, int slot_right, float* flow, int variant, int clip,
                                  void* stream) {
    TF_REQUIRE(h && flow, TF_ERR_INVALID_ARG, "tf_farneback_solve: null argument");
    TF_REQUIRE(slot_ok(slot_left) && slot_ok(slot_right) && h->has_frame[slot_left] && h->has_frame[slot_right],
               TF_ERR_INVALID_ARG, "tf_farneback_solve: slots must be prepared slots in [0, %d)", FB_SLOTS);
    TF_REQUIRE(variant >= 0 && variant <= 31 && variant != 26, TF_ERR_INVALID_ARG, "tf_farneback_solve: unknown variant %d", variant);
    TF_REQUIRE(((uintptr_t)flow & 7) == 0, TF_ERR_INVALID_ARG, "tf_farneback_solve: flow must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    float2* out = reinterpret_cast<float2*>(flow);
    return h->r_fp16 ? solve_impl<__half>(h, slot_left, slot_right, out, variant, clip, st)
                     : solve_impl<float>(h, slot_left, slot_right, out, variant, clip, st);
}

extern "C" int tf_farneback_step_lane(tf_farneback* h, int lane, int new_slot, const uint8_t* gray, int slot_left,
                                      int slot_right, float* flow, int variant, int clip, void* stream) {
    TF_REQUIRE(h && gray && flow, TF_ERR_INVALID_ARG, "tf_farneback_step: null argument");
    TF_REQUIRE(lane >= 0 && lane < FB_LANES, TF_ERR_INVALID_ARG, "tf_farneback_step: lane must be in [0, %d)", FB_LANES);
    TF_REQUIRE(slot_ok(new_slot) && slot_ok(slot_left) && slot_ok(slot_right), TF_ERR_INVALID_ARG,
               "tf_farneback_step: slots must be in [0, %d)", FB_SLOTS);
    TF_REQUIRE((new_slot == slot_left) != (new_slot == slot_right), TF_ERR_INVALID_ARG,
               "tf_farneback_step: the new frame must be exactly one side of the pair");
    TF_REQUIRE(variant >= 0 && variant <= 31 && variant != 26, TF_ERR_INVALID_ARG, "tf_farneback_step: unknown variant %d", variant);
    TF_REQUIRE(variant != 1 || lane == 0, TF_ERR_INVALID_ARG,
               "tf_farneback_step: the unfused reference kernels (variant 1) share their scratch, lane 0 only");
    TF_REQUIRE(((uintptr_t)flow & 7) == 0, TF_ERR_INVALID_ARG, "tf_farneback_step: flow must be 8-byte aligned");
    const int old_slot = new_slot == slot_left ? slot_right : slot_left;
    TF_REQUIRE(h->has_frame[old_slot], TF_ERR_INVALID_ARG, "tf_farneback_step: slot %d was never prepared", old_slot);
    if (int e = ensure_slot(h, new_slot)) return e;
    h->has_frame[new_slot] = true;
    if (int e = ensure_lane(h, lane)) return e;
    cudaStream_t st = as_stream(stream);
    if (!h->aux) {
        TF_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
        TF_CUDA(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
        for (int s = 0; s < FB_SLOTS; s++) {
            h->ev_level[s].resize(h->lv.size());
            for (auto& e : h->ev_level[s]) TF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (int l = 0; l < FB_LANES; l++)
                TF_CUDA(cudaEventCreateWithFlags(&h->ev_read[s][l], cudaEventDisableTiming));
        }
    }
    // the auxiliary stream may overwrite new_slot only after everything already queued on the caller's
    // stream (`gray` was produced there; this lane's earlier solves) and the other lanes' solves that read
    // the slot have finished
    TF_CUDA(cudaEventRecord(h->ev_start, st));
    TF_CUDA(cudaStreamWaitEvent(h->aux, h->ev_start, 0));
    for (int l = 0; l < FB_LANES; l++)
        if (l != lane) TF_CUDA(cudaStreamWaitEvent(h->aux, h->ev_read[new_slot][l], 0));
    const bool batched = fused_blur_ok(h);
    if (batched)
        if (int e = blur_levels(h, gray, h->aux)) return e;
    for (size_t li = 0; li < h->lv.size(); li++) {
        if (int e = prepare_level(h, h->lv[li], new_slot, gray, h->aux, batched)) return e;
        TF_CUDA(cudaEventRecord(h->ev_level[new_slot][li], h->aux));
    }
    float2* out = reinterpret_cast<float2*>(flow);
    int e = h->r_fp16 ? solve_impl<__half>(h, slot_left, slot_right, out, variant, clip, st, true, lane)
                      : solve_impl<float>(h, slot_left, slot_right, out, variant, clip, st, true, lane);
    if (e) return e;
    TF_CUDA(cudaEventRecord(h->ev_read[slot_left][lane], st));
    TF_CUDA(cudaEventRecord(h->ev_read[slot_right][lane], st));
    return TF_OK;
}

extern "C" int tf_farneback_step(tf_farneback* h, int new_slot, const uint8_t* gray, int slot_left, int slot_right,
                                 float* flow, int variant, int clip, void* stream) {
    return tf_farneback_step_lane(h, 0, new_slot, gray, slot_left, slot_right, flow, variant, clip, stream);
}

extern "C" int tf_farneback_run(tf_farneback* h, const uint8_t* left, const uint8_t* right, float* flow, int variant,
                                void* stream) {
    if (int e = tf_farneback_prepare(h, 0, left, stream)) return e;
    if (int e = tf_farneback_prepare(h, 1, right, stream)) return e;
    return tf_farneback_solve(h, 0, 1, flow, variant, 0, stream);
}

// test hook: "4+1" storage -> five float planes (5, h, w)
template <typename RT>
__global__ void __launch_bounds__(256) k_r_to_planes(const RT* __restrict__ R, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 q = load_quad(R, i);
    out[i] = q.x;
    out[n + i] = q.y;
    out[2 * n + i] = q.z;
    out[3 * n + i] = q.w;
    out[4 * n + i] = load_c4(R, n, i);
}

extern "C" int tf_farneback_debug_read(tf_farneback* h, int slot, int li, int what, float* out, void* stream) {
    TF_REQUIRE(h && out, TF_ERR_INVALID_ARG, "tf_farneback_debug_read: null argument");
    TF_REQUIRE(li >= 0 && li < (int)h->lv.size(), TF_ERR_INVALID_ARG, "tf_farneback_debug_read: bad level %d", li);
    TF_REQUIRE(slot_ok(slot) && h->lv[li].R[slot], TF_ERR_INVALID_ARG, "tf_farneback_debug_read: bad slot");
    cudaStream_t st = as_stream(stream);
    FbLevel& L = h->lv[li];
    size_t n = (size_t)L.w * L.h;
    if (what == 0) {
        bool fused = L.w == h->W && L.h == h->H && L.ksz == 3 && L.sigma <= 0;
        TF_REQUIRE(!fused || h->publish_finest, TF_ERR_INVALID_ARG,
                   "the finest level's image only exists after tf_farneback_set_debug(handle, 1) + prepare");
        TF_CUDA(cudaMemcpyAsync(out, L.img, n * 4, cudaMemcpyDeviceToDevice, st));
    } else if (what == 1) {
        unsigned blocks = (unsigned)((n + 255) / 256);
        if (h->r_fp16)
            k_r_to_planes<__half><<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(L.R[slot]), out, n);
        else
            k_r_to_planes<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(L.R[slot]), out, n);
        TF_LAUNCHED();
    } else if (what == 2) {
        TF_REQUIRE(li + 1 < (int)h->lv.size(), TF_ERR_INVALID_ARG,
                   "tf_farneback_debug_read: the finest level's flow is the solve output");
        TF_CUDA(cudaMemcpyAsync(out, L.flow[0], n * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        return fail(TF_ERR_INVALID_ARG, "tf_farneback_debug_read: unknown selector %d", what);
    }
    return TF_OK;
}

extern "C" double tf_farneback_algorithmic_bytes(const tf_farneback* h, int reuse_r) {
    // (L+1)*N u8 reads + 20*S R writes (new frame) + T*S*(20+20+8+8);
    // without reuse the second frame's pyramid + R are added.
    if (!h) return 0;
    double N = (double)h->H * h->W, S = 0;
    for (auto& L : h->lv) S += (double)L.w * L.h;
    double levels = (double)h->lv.size();
    double b = levels * N + 20.0 * S + h->iterations * S * 56.0;
    if (!reuse_r) b += levels * N + 20.0 * S;
    return b;
}
