// Fused Farneback iteration, "rolling tile" kernel (solve variant 3, the default):
// update-matrices + (2m+1)^2 box sums + 2x2 solve in ONE kernel; the 5-channel matrix M lives
// only in shared memory (SURVEY.md 8d byte model: per iteration read R0, R1, flow; write flow).
//
// A CTA (512 threads, 2 CTAs per SM) owns a strip of TX = 64 output columns (+ m halo columns each side) and
// walks down a chunk of rows in tiles of TY = 32 rows.  Per tile:
//   A  each thread owns one halo'd column and computes M for every sixth new row of it (R0
//      coalesced, R1 4-tap bilinear gather); the 2m bottom matrix rows of the previous tile are
//      moved to the top of the buffer instead of being recomputed, so the vertical halo is paid
//      once per chunk, not once per tile;
//   B  one thread per (column, channel) turns M into vertical window sums IN PLACE, 8 outputs at
//      a time: a direct (2m+1)-term sum, then 7 slides (fp32 drift is bounded by 7 steps);
//   C  one thread per (row, 8-column segment) forms the horizontal window sums the same way for
//      all 5 channels, solves the 2x2 system with error-free fp32 products (the determinant
//      cancels), and stores its 8 flow vectors as four 128-bit words.
// The window radius is a template parameter so every shared-memory offset is an immediate.
// Layout [channel][row][PITCH], PITCH = TX + 2m rounded up to 1 (mod 8): phase B (lanes =
// consecutive columns) and phase C (lanes = 4 segments x 8 rows) are bank-conflict free.
// Borders replicate (cv2 clamps srow / vsum), so M is evaluated at clamped coordinates.
#pragma once
#include "fb_math.cuh"

#define FBT_TX 64
#define FBT_TY 32
#define FBT_NT 512
#define FBT_NOUT 8

__device__ __forceinline__ float fbt_diff_of_products(float a, float b, float c, float d) {
    // a*b - c*d evaluated as if in higher precision (Kahan's FMA trick)
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    float dop = fmaf(a, b, -cd);
    return dop + err;
}

__device__ __forceinline__ float2 fbt_solve(const float* s, float scale) {
    float g11 = s[0] * scale, g12 = s[1] * scale, g22 = s[2] * scale, h1 = s[3] * scale, h2 = s[4] * scale;
    float det = fbt_diff_of_products(g11, g22, g12, g12) + 1e-3f;
    float nx = fbt_diff_of_products(g11, h2, g12, h1);
    float ny = fbt_diff_of_products(g22, h1, g12, h2);
    return make_float2(__fdiv_rn(nx, det), __fdiv_rn(ny, det));
}

template <int MR>
struct FbtGeom {
    static constexpr int WIN = 2 * MR + 1;
    static constexpr int NR = FBT_TY + 2 * MR;
    static constexpr int COLS = FBT_TX + 2 * MR;
    static constexpr int PITCH = COLS + ((9 - (COLS & 7)) & 7);  // == 1 (mod 8)
    static constexpr int CHS = NR * PITCH;
    static constexpr int RPT = FBT_NT / COLS;                    // matrix rows per phase-A pass
    static constexpr size_t SMEM = (size_t)5 * CHS * sizeof(float);
};

template <typename RT, int MR>
__global__ void __launch_bounds__(FBT_NT, 2) k_fb_iter_tile(const RT* __restrict__ R0, const RT* __restrict__ R1,
                                                            const float2* __restrict__ flow_in,
                                                            float2* __restrict__ flow_out, int w, int h, float scale,
                                                            int rows_per_cta, int clip) {
    using G = FbtGeom<MR>;
    extern __shared__ __align__(16) float ring[];  // [5][NR][PITCH]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FBT_TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const size_t plane = (size_t)w * h;

    // phase A ownership: column lxA, rows rA, rA + RPT, ...
    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < G::RPT;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);

    for (int ty = y0; ty < y1; ty += FBT_TY) {
        const int nout = min(FBT_TY, y1 - ty);  // output rows of this tile
        int lr_begin = 0;
        if (ty != y0) {
            // keep the last 2m matrix rows of the previous tile: rows [TY, TY + 2m) -> [0, 2m)
            // (rows are contiguous with stride PITCH, so each channel is one linear block)
#pragma unroll
            for (int c = 0; c < 5; c++)
                for (int i = tid; i < 2 * MR * G::PITCH; i += FBT_NT)
                    ring[c * G::CHS + i] = ring[c * G::CHS + FBT_TY * G::PITCH + i];
            lr_begin = 2 * MR;
            __syncthreads();
        }
        // ---- phase A: matrix rows local [lr_begin, nout + 2m), global row = ty - m + local ----
        if (activeA) {
            // software pipeline: the flow vector and the R0 coefficients of the NEXT row are requested
            // before the current row's R1 gather, so only one memory latency is exposed per pixel
            const int lr_end = nout + 2 * MR;
            int lr = lr_begin + rA;
            float2 f_nx = make_float2(0.f, 0.f);
            float a_nx[5];
            int gy_nx = 0;
            if (lr < lr_end) {
                gy_nx = clampi(ty - MR + lr, 0, h - 1);
                size_t at = (size_t)gy_nx * w + gxA;
                if (flow_in) f_nx = __ldg(flow_in + at);
                fb_load_r0<RT>(R0, plane, at, a_nx);
            }
            for (; lr < lr_end; lr += G::RPT) {
                float2 f = f_nx;
                float a[5];
#pragma unroll
                for (int c = 0; c < 5; c++) a[c] = a_nx[c];
                const int gy = gy_nx;
                if (lr + G::RPT < lr_end) {
                    gy_nx = clampi(ty - MR + lr + G::RPT, 0, h - 1);
                    size_t at = (size_t)gy_nx * w + gxA;
                    if (flow_in) f_nx = __ldg(flow_in + at);
                    fb_load_r0<RT>(R0, plane, at, a_nx);
                }
                float mm[5];
                fb_update_matrix_pre<RT>(a, R1, plane, w, h, gxA, gy, f, mm);
                float* dst = ring + lr * G::PITCH + lxA;
#pragma unroll
                for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
            }
        }
        __syncthreads();
        // ---- phase B: vertical window sums in place; output row j overwrites matrix row j ----
        for (int item = tid; item < 5 * G::COLS; item += FBT_NT) {
            int c = item / G::COLS, lx = item - c * G::COLS;
            float* col = ring + c * G::CHS + lx;
#pragma unroll
            for (int j0 = 0; j0 < FBT_TY; j0 += 8) {
                if (j0 < nout) {
                    float v[8];
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < G::WIN; k++) s += col[(j0 + k) * G::PITCH];
                    v[0] = s;
#pragma unroll
                    for (int j = 1; j < 8; j++) {
                        s += col[(j0 + j + G::WIN - 1) * G::PITCH] - col[(j0 + j - 1) * G::PITCH];
                        v[j] = s;
                    }
                    // rows j0..j0+7 are dead as matrix values now (later blocks start at j0 + 8, the next
                    // tile reuses rows >= TY only); rows past nout may hold garbage and are never read
#pragma unroll
                    for (int j = 0; j < 8; j++) col[(j0 + j) * G::PITCH] = v[j];
                }
            }
        }
        __syncthreads();
        // ---- phase C: horizontal window sums + solve; lane = (4 segments) x (8 rows) ----
        {
            const int lane = tid & 31, warp = tid >> 5;
            const int seg = (lane & 3) + 4 * (warp & 1);   // 8 segments of 8 columns
            const int row = (lane >> 2) + 8 * (warp >> 1);  // 32 rows
            if (row < nout && tid < 256) {
                const float* rp = ring + row * G::PITCH + seg * FBT_NOUT;
                float s[5];
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < G::WIN; k++) a += rp[c * G::CHS + k];
                    s[c] = a;
                }
                const int y = ty + row;
                const int xg = x0 + seg * FBT_NOUT;
                float2 out[FBT_NOUT];
#pragma unroll
                for (int o = 0; o < FBT_NOUT; o++) {
                    if (o > 0) {
#pragma unroll
                        for (int c = 0; c < 5; c++) s[c] += rp[c * G::CHS + o + G::WIN - 1] - rp[c * G::CHS + o - 1];
                    }
                    float2 fl = fbt_solve(s, scale);
                    if (clip) {
                        int x = xg + o;
                        fl.x = fminf(fmaxf(fl.x, (float)(-x)), (float)(w - 1 - x));
                        fl.y = fminf(fmaxf(fl.y, (float)(-y)), (float)(h - 1 - y));
                    }
                    out[o] = fl;
                }
                float2* dst = flow_out + (size_t)y * w + xg;
                if (xg + FBT_NOUT <= w && (w & 1) == 0) {
                    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                    for (int o = 0; o < FBT_NOUT; o += 2)
                        d4[o >> 1] = make_float4(out[o].x, out[o].y, out[o + 1].x, out[o + 1].y);
                } else {
#pragma unroll
                    for (int o = 0; o < FBT_NOUT; o++)
                        if (xg + o < w) dst[o] = out[o];
                }
            }
        }
        __syncthreads();  // the next tile's row move / phase A overwrite what phase C just read
    }
}

template <typename RT, int MR>
static int fb_launch_tile(const RT* R0, const RT* R1, const float2* in, float2* dst, int w, int h, float scale,
                          int clip, cudaStream_t st) {
    using G = FbtGeom<MR>;
    auto kern = k_fb_iter_tile<RT, MR>;
    static bool attr_set = false;
    if (!attr_set) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        attr_set = true;
    }
    // Chunk height: a CTA costs about (rows + 2m) matrix rows (the vertical halo is paid once per chunk)
    // and the CTAs resident on one SM share its throughput, so the kernel takes about
    // (CTAs per SM) x (rows + 2m): whole CTAs while the grid is under two per SM, CTAs/SMs plus half a CTA
    // of tail once the hardware scheduler can balance.  Pick the cheapest multiple of the tile height.
    int strips = ceil_div(w, FBT_TX);
    int sms = sm_count();
    int rows = FBT_TY;
    double best = 1e30;
    for (int r = FBT_TY; r < h + FBT_TY; r += FBT_TY) {
        int ctas = strips * ceil_div(h, r);
        double per_sm = ctas <= 2 * sms ? (double)ceil_div(ctas, sms) : (double)ctas / sms + 0.5;
        double cost = per_sm * (std::min(r, h) + 2 * MR);
        if (cost < best) {
            best = cost;
            rows = r;
        }
    }
    dim3 grid(strips, ceil_div(h, rows));
    kern<<<grid, FBT_NT, G::SMEM, st>>>(R0, R1, in, dst, w, h, scale, rows, clip);
    return TF_OK;
}

template <typename RT>
static int fb_iterate_tile(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf,
                           float2* other_buf, bool zero_init, int clip, bool finest, cudaStream_t st) {
    int m = h->winsize / 2;
    float scale = (float)(1.0 / ((double)h->winsize * h->winsize));
    int T = h->iterations;
    for (int it = 0; it < T; it++) {
        float2* dst = ((T - 1 - it) & 1) ? other_buf : final_buf;
        float2* src = ((T - 1 - it) & 1) ? final_buf : other_buf;
        const float2* in = (it == 0 && zero_init) ? nullptr : src;
        int c = clip && it + 1 == T;
        int e = TF_OK;
        {
            ScopedKernelTimer timer(finest ? TFK_FB_ITER_FINEST : -1, st);
            switch (m) {
#define TF_FBT(MR) case MR: e = fb_launch_tile<RT, MR>(R0, R1, in, dst, L.w, L.h, scale, c, st); break;
                TF_FBT(1) TF_FBT(2) TF_FBT(3) TF_FBT(4) TF_FBT(5) TF_FBT(6) TF_FBT(7) TF_FBT(8)
                TF_FBT(9) TF_FBT(10) TF_FBT(11) TF_FBT(12) TF_FBT(13) TF_FBT(14) TF_FBT(15) TF_FBT(16)
#undef TF_FBT
                default: return fail(TF_ERR_INVALID_ARG, "unsupported window radius %d", m);
            }
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
