// Fused Farneback iteration, "half-buffer" kernel (solve variants 4-6; the default for window radius >= 4):
// update-matrices + (2m+1)^2 box sums + 2x2 solve in ONE kernel, M only in shared memory, like the rolling
// tile kernel (fb_tile.cuh), reorganised around what bounds it on B200 -- instruction issue and L1/shared
// bandwidth, not HBM:
//
//  * a tile is TY = 2m output rows; shared memory holds two HALVES of 2m matrix rows each.  Output row j of
//    the tile needs matrix rows j .. j + 2m of (old half ++ new half) = old[j ..] and new[.. j], so the
//    vertical window sum is  suffix(old)[j] + prefix(new)[j] : 2(2m - 1) + 2m additions per 2m outputs, no
//    subtraction (no drift, no cancellation), and nothing is moved between tiles: the halves swap roles;
//  * phase A (matrix rows of the new half): a thread owns one halo'd column and every NG-th row; R0 / flow
//    of its next row are requested before the current row's R1 gather.  (Tried and dropped, all slower on
//    B200: cp.async staging of the taps into per-thread shared slots, 215 us; runs of consecutive rows
//    per thread that keep the bottom taps as the next row's top taps, 225-290 us -- register rotation and
//    a lower L1 hit rate cost more than the saved loads.)
//  * phase B: one thread per (column pair, channel): 2m old values into registers, suffix sums, then stream
//    the new half (64-bit shared accesses); the sums overwrite the old half (dead after this tile);
//  * phase C: one thread per (row, 4-column segment) in the default configuration (SEG = 4: a warp = 8 segments x
//    4 rows, every warp of the CTA has work; SEG = 8 is the earlier split): the (SEG + 2m)-wide window is read with
//    128-bit shared loads, horizontal window sums (direct sum + slides), 2x2 solve with error-free fp32 products,
//    128-bit stores.
// PITCH = 4 (mod 8) floats: rows are 16-byte aligned and a quarter-warp of phase C (4 segments x 2 rows)
// touches 8 distinct 16-byte bank groups.
#pragma once
#include "fb_tile.cuh"

template <int MR, int TX, bool VEC>
struct FbhGeom {
    static constexpr int TY = 2 * MR;                            // output rows per tile = rows per half
    static constexpr int WIN = 2 * MR + 1;
    static constexpr int COLS = TX + 2 * MR;
    static constexpr int NQ = 2 + (MR + 1) / 2;                  // 128-bit words of a phase-C window
    static constexpr int NEED = (TX - 8 + 4 * NQ) > COLS ? (TX - 8 + 4 * NQ) : COLS;
    // VEC: == 4 (mod 8), wide enough for the 128-bit window reads; scalar: == 1 (mod 8)
    static constexpr int PITCH = VEC ? NEED + ((12 - (NEED & 7)) & 7) : COLS + ((9 - (COLS & 7)) & 7);
    static constexpr int HALF = TY * PITCH;
    static constexpr int CHS = 2 * HALF;                         // channel stride
    static constexpr size_t SMEM = (size_t)5 * CHS * sizeof(float);
};

// (s0 s1; s1 s2) x = (s3 s4) with the window sums left unscaled: dividing numerator and denominator of
// cv2's expression by scale^2 moves the scale into the regulariser (reg = 1e-3 / scale^2).
__device__ __forceinline__ float2 fbh_solve(float s0, float s1, float s2, float s3, float s4, float reg) {
    float det = fbt_diff_of_products(s0, s2, s1, s1) + reg;
    float nx = fbt_diff_of_products(s0, s4, s1, s3);
    float ny = fbt_diff_of_products(s2, s3, s1, s4);
    float r = __frcp_rn(det);
    return make_float2(nx * r, ny * r);
}

// R0 and the flow are read once per CTA: keep them out of L1 so that the R1 taps (each used by ~4 neighbouring
// pixels) stay resident.  g_fbh_stream (tf_farneback_tune key 2) switches the hint off for comparison.
__device__ __forceinline__ float4 fbh_ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 fbh_ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float fbh_ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// one row of bilinear taps: pixels (x1, y) and (x1 + 1, y) of R in the "4+1" layout
struct FbhTaps {
    float4 q0, q1;
    float e0, e1;
};
__device__ __forceinline__ FbhTaps fbh_load_taps(const float4* __restrict__ Rq, const float* __restrict__ Re,
                                                 unsigned px) {
    FbhTaps t;
    t.q0 = __ldg(Rq + px);
    t.q1 = __ldg(Rq + px + 1);
    t.e0 = __ldg(Re + px);
    t.e1 = __ldg(Re + px + 1);
    return t;
}

// FarnebackUpdateMatrices for one pixel from already-fetched operands (same arithmetic as
// fb_update_matrix_pre): a = R0 at (x, y); top / bot = the taps of R1 at rows y1 and y1 + 1.
__device__ __forceinline__ void fbh_matrix(const float* a, float2 f, int x, int y, int w, int h, bool in,
                                           const FbhTaps& top, const FbhTaps& bot, float* m) {
    float dx = f.x, dy = f.y;
    float fx = (float)x + dx, fy = (float)y + dy;
    fx -= floorf(fx);
    fy -= floorf(fy);
    float r2, r3, r4, r5, r6;
    if (in) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * top.q0.x + a01 * top.q1.x + a10 * bot.q0.x + a11 * bot.q1.x;
        r3 = a00 * top.q0.y + a01 * top.q1.y + a10 * bot.q0.y + a11 * bot.q1.y;
        r4 = a00 * top.q0.z + a01 * top.q1.z + a10 * bot.q0.z + a11 * bot.q1.z;
        r5 = a00 * top.q0.w + a01 * top.q1.w + a10 * bot.q0.w + a11 * bot.q1.w;
        r6 = a00 * top.e0 + a01 * top.e1 + a10 * bot.e0 + a11 * bot.e1;
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

// ---- phase B: vertical window sums of (old half ++ new half) into the old half ----
template <typename G, int NT, bool VEC>
__device__ __forceinline__ void fbh_phase_b(float* __restrict__ old_half, const float* __restrict__ new_half, int tid) {
    if (VEC) {  // two columns per thread, 64-bit accesses
        for (int item = tid; item < 5 * (G::COLS / 2); item += NT) {
            int c = item / (G::COLS / 2), lx = 2 * (item - c * (G::COLS / 2));
            float2* oc = reinterpret_cast<float2*>(old_half + c * G::CHS + lx);
            const float2* nc = reinterpret_cast<const float2*>(new_half + c * G::CHS + lx);
            float2 v[G::TY];
#pragma unroll
            for (int j = 0; j < G::TY; j++) v[j] = oc[j * (G::PITCH / 2)];
#pragma unroll
            for (int j = G::TY - 2; j >= 0; j--) {
                v[j].x += v[j + 1].x;
                v[j].y += v[j + 1].y;
            }
            float2 p = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < G::TY; j++) {
                float2 nv = nc[j * (G::PITCH / 2)];
                p = j == 0 ? nv : make_float2(p.x + nv.x, p.y + nv.y);
                oc[j * (G::PITCH / 2)] = make_float2(v[j].x + p.x, v[j].y + p.y);
            }
        }
    } else {
        for (int item = tid; item < 5 * G::COLS; item += NT) {
            int c = item / G::COLS, lx = item - c * G::COLS;
            float* oc = old_half + c * G::CHS + lx;
            const float* nc = new_half + c * G::CHS + lx;
            float v[G::TY];
#pragma unroll
            for (int j = 0; j < G::TY; j++) v[j] = oc[j * G::PITCH];
#pragma unroll
            for (int j = G::TY - 2; j >= 0; j--) v[j] += v[j + 1];
            float p = 0.f;
#pragma unroll
            for (int j = 0; j < G::TY; j++) {
                float nv = nc[j * G::PITCH];
                p = j == 0 ? nv : p + nv;
                oc[j * G::PITCH] = v[j] + p;
            }
        }
    }
}

// phase B of fb_half.cuh with PAIRS column pairs per channel (PAIRS >= COLS / 2, 2 * PAIRS <= PITCH): the padding
// pairs are summed too (finite or not, phase C never reads them) so that consecutive items stay on consecutive banks
// across a channel boundary
template <typename G, int NT, int PAIRS>
__device__ __forceinline__ void fbr_phase_b(float* __restrict__ old_half, const float* __restrict__ new_half, int tid) {
    for (int item = tid; item < 5 * PAIRS; item += NT) {
        int c = item / PAIRS, lx = 2 * (item - c * PAIRS);
        float2* oc = reinterpret_cast<float2*>(old_half + c * G::CHS + lx);
        const float2* nc = reinterpret_cast<const float2*>(new_half + c * G::CHS + lx);
        float2 v[G::TY];
#pragma unroll
        for (int j = 0; j < G::TY; j++) v[j] = oc[j * (G::PITCH / 2)];
#pragma unroll
        for (int j = G::TY - 2; j >= 0; j--) {
            v[j].x += v[j + 1].x;
            v[j].y += v[j + 1].y;
        }
        float2 p = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < G::TY; j++) {
            float2 nv = nc[j * (G::PITCH / 2)];
            p = j == 0 ? nv : make_float2(p.x + nv.x, p.y + nv.y);
            oc[j * (G::PITCH / 2)] = make_float2(v[j].x + p.x, v[j].y + p.y);
        }
    }
}

// smallest number of column pairs >= base with 2 * pairs == chs (mod 32) that still fits the pitch; else base
constexpr int fbr_pick_pairs(int base, int pitch, int chs) {
    for (int p = base; 2 * p <= pitch; p++)
        if ((2 * p - chs) % 32 == 0) return p;
    return base;
}
template <int MR, int TX>
struct FbrPairs {
    using G = FbhGeom<MR, TX, true>;
    static constexpr int PAIRS = fbr_pick_pairs((G::COLS + 1) / 2, G::PITCH, G::CHS);
};

// ---- phase C: horizontal window sums + 2x2 solve + store for the tile's nout rows ----
template <typename G, int TX, int NT, bool VEC, int SEG = 8>
__device__ __forceinline__ void fbh_phase_c(const float* __restrict__ old_half, float2* __restrict__ flow_out, int tid,
                                            int x0, int ty, int nout, int w, int h, float reg, int clip) {
    if (SEG == 4) {
        // finer split (variant 17): one thread per (row, 4-column segment), a warp = 8 segments x 4 rows, so that a
        // 64-column strip has 8 warp items per tile instead of 4 (every warp of the CTA has work); the window is
        // 4 + 2m floats = five 128-bit loads per channel, a quarter-warp reads 8 consecutive 16-byte groups
        static_assert(SEG != 4 || VEC, "the 4-column split uses the vector layout");
        constexpr int NSEG = TX / 4;
        constexpr int WX = NSEG / 8;
        constexpr int NQ4 = (4 + G::WIN - 1 + 3) / 4;
        constexpr int NWI = ((G::TY + 3) / 4) * WX;
        const int lane = tid & 31;
        for (int wi = tid >> 5; wi < NWI; wi += NT / 32) {
            const int seg = (lane & 7) + 8 * (wi % WX);
            const int row = (lane >> 3) + 4 * (wi / WX);
            if (row < nout) {
                const int y = ty + row;
                const int xg = x0 + seg * 4;
                float2* dst = flow_out + (size_t)y * w + xg;
                const bool wide = xg + 4 <= w && (w & 1) == 0;
                const float4* rp = reinterpret_cast<const float4*>(old_half + row * G::PITCH + seg * 4);
                float sum[5][4];
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    float win[4 * NQ4];
#pragma unroll
                    for (int k = 0; k < NQ4; k++) {
                        float4 t = rp[c * (G::CHS / 4) + k];
                        win[4 * k] = t.x; win[4 * k + 1] = t.y; win[4 * k + 2] = t.z; win[4 * k + 3] = t.w;
                    }
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < G::WIN; k++) acc += win[k];
                    sum[c][0] = acc;
#pragma unroll
                    for (int o = 1; o < 4; o++) {
                        acc += win[o + G::WIN - 1] - win[o - 1];
                        sum[c][o] = acc;
                    }
                }
#pragma unroll
                for (int o = 0; o < 4; o += 2) {
                    float2 u = fbh_solve(sum[0][o], sum[1][o], sum[2][o], sum[3][o], sum[4][o], reg);
                    float2 v = fbh_solve(sum[0][o + 1], sum[1][o + 1], sum[2][o + 1], sum[3][o + 1], sum[4][o + 1], reg);
                    if (clip) {
                        u.x = fminf(fmaxf(u.x, (float)(-(xg + o))), (float)(w - 1 - (xg + o)));
                        u.y = fminf(fmaxf(u.y, (float)(-y)), (float)(h - 1 - y));
                        v.x = fminf(fmaxf(v.x, (float)(-(xg + o + 1))), (float)(w - 2 - (xg + o)));
                        v.y = fminf(fmaxf(v.y, (float)(-y)), (float)(h - 1 - y));
                    }
                    if (wide) {
                        reinterpret_cast<float4*>(dst)[o >> 1] = make_float4(u.x, u.y, v.x, v.y);
                    } else {
                        if (xg + o < w) dst[o] = u;
                        if (xg + o + 1 < w) dst[o + 1] = v;
                    }
                }
            }
        }
    } else {
        constexpr int NSEG = TX / 8;
        constexpr int WX = NSEG / 4;  // warp items per group of 8 rows
        constexpr int NWI = ((G::TY + 7) / 8) * WX;
        const int lane = tid & 31;
        for (int wi = tid >> 5; wi < NWI; wi += NT / 32) {
            const int seg = (lane & 3) + 4 * (wi % WX);
            const int row = (lane >> 2) + 8 * (wi / WX);
            if (row < nout) {
                const int y = ty + row;
                const int xg = x0 + seg * 8;
                float2* dst = flow_out + (size_t)y * w + xg;
                const bool wide = xg + 8 <= w && (w & 1) == 0;
                // clip (last iteration only) and store outputs o, o + 1 of the segment
                auto emit = [&](int o, float2 u, float2 v) {
                    if (clip) {
                        u.x = fminf(fmaxf(u.x, (float)(-(xg + o))), (float)(w - 1 - (xg + o)));
                        u.y = fminf(fmaxf(u.y, (float)(-y)), (float)(h - 1 - y));
                        v.x = fminf(fmaxf(v.x, (float)(-(xg + o + 1))), (float)(w - 2 - (xg + o)));
                        v.y = fminf(fmaxf(v.y, (float)(-y)), (float)(h - 1 - y));
                    }
                    if (wide) {
                        reinterpret_cast<float4*>(dst)[o >> 1] = make_float4(u.x, u.y, v.x, v.y);
                    } else {
                        if (xg + o < w) dst[o] = u;
                        if (xg + o + 1 < w) dst[o + 1] = v;
                    }
                };
                if (!VEC) {  // scalar sliding windows, all channels abreast
                    const float* sp = old_half + row * G::PITCH + seg * 8;
                    float s5[5];
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < G::WIN; k++) acc += sp[c * G::CHS + k];
                        s5[c] = acc;
                    }
                    float2 res[8];
#pragma unroll
                    for (int o = 0; o < 8; o++) {
                        if (o > 0) {
#pragma unroll
                            for (int c = 0; c < 5; c++)
                                s5[c] += sp[c * G::CHS + o + G::WIN - 1] - sp[c * G::CHS + o - 1];
                        }
                        res[o] = fbh_solve(s5[0], s5[1], s5[2], s5[3], s5[4], reg);
                    }
#pragma unroll
                    for (int o = 0; o < 8; o += 2) emit(o, res[o], res[o + 1]);
                } else {  // 128-bit window reads, channel by channel
                    const float4* rp = reinterpret_cast<const float4*>(old_half + row * G::PITCH + seg * 8);
                    float sum[5][8];
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        float win[4 * G::NQ];
#pragma unroll
                        for (int k = 0; k < G::NQ; k++) {
                            float4 t = rp[c * (G::CHS / 4) + k];
                            win[4 * k] = t.x; win[4 * k + 1] = t.y; win[4 * k + 2] = t.z; win[4 * k + 3] = t.w;
                        }
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < G::WIN; k++) acc += win[k];
                        sum[c][0] = acc;
#pragma unroll
                        for (int o = 1; o < 8; o++) {
                            acc += win[o + G::WIN - 1] - win[o - 1];
                            sum[c][o] = acc;
                        }
                    }
#pragma unroll
                    for (int o = 0; o < 8; o += 2)
                        emit(o, fbh_solve(sum[0][o], sum[1][o], sum[2][o], sum[3][o], sum[4][o], reg),
                             fbh_solve(sum[0][o + 1], sum[1][o + 1], sum[2][o + 1], sum[3][o + 1], sum[4][o + 1], reg));
                }
            }
        }
    }
}

template <int MR, int TX, int NT, int WANT, bool VEC>
struct FbhCfg {
    using G = FbhGeom<MR, TX, VEC>;
    static constexpr int NG = NT / G::COLS;  // phase-A row groups
    static constexpr int FIT = (int)((227 * 1024) / (G::SMEM + 1024));
    static constexpr int CTAS = FIT < 1 ? 1 : (FIT > WANT ? WANT : FIT);
};

template <int MR, int TX, int NT, int WANT, bool VEC, int SEG = 8, bool BFIX = false>
__global__ void __launch_bounds__(NT, (FbhCfg<MR, TX, NT, WANT, VEC>::CTAS))
    k_fb_iter_half(const float4* __restrict__ R0q, const float* __restrict__ R0e, const float4* __restrict__ R1q,
                   const float* __restrict__ R1e, const float2* __restrict__ flow_in, float2* __restrict__ flow_out,
                   int w, int h, float reg, int rows_per_cta, int clip) {
    using G = FbhGeom<MR, TX, VEC>;
    using S = FbhCfg<MR, TX, NT, WANT, VEC>;
    static_assert(NT >= G::COLS, "one thread per halo'd column needed");
    static_assert(TX % 32 == 0, "a warp covers whole rows of segments");
    extern __shared__ __align__(16) float ring[];  // [5][2 halves][TY][PITCH]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;  // halves 0 .. ntiles (half 0 = prologue)

    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < S::NG;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);
    const unsigned uw = (unsigned)w;

    for (int hh = 0; hh <= ntiles; hh++) {
        float* new_half = ring + (hh & 1) * G::HALF;
        // ---- phase A: matrix rows of half hh = global rows y0 - m + hh * TY + [0, TY), clamped ----
        if (activeA) {
            const int gy_base = y0 - MR + hh * G::TY;
            int r = rA;
            int gy_nx = clampi(gy_base + r, 0, h - 1);
            unsigned at = (unsigned)gy_nx * uw + (unsigned)gxA;
            float2 f_nx = flow_in ? fbh_ld_stream(flow_in + at) : make_float2(0.f, 0.f);
            float4 q_nx = fbh_ld_stream(R0q + at);
            float e_nx = fbh_ld_stream(R0e + at);
#pragma unroll 1
            for (; r < G::TY; r += S::NG) {
                const float2 f = f_nx;
                const float a[5] = {q_nx.x, q_nx.y, q_nx.z, q_nx.w, e_nx};
                const int gy = gy_nx;
                // cvFloor; the float->int conversion saturates, so absurd displacements land outside the image
                int x1 = __float2int_rd((float)gxA + f.x), yy1 = __float2int_rd((float)gy + f.y);
                const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1);
                FbhTaps top, bot;
                if (in) {
                    unsigned q = (unsigned)yy1 * uw + (unsigned)x1;
                    top = fbh_load_taps(R1q, R1e, q);
                    bot = fbh_load_taps(R1q, R1e, q + uw);
                }
                // next row's R0 / flow, requested after this row's gather (the gather is the load this row waits for)
                if (r + S::NG < G::TY) {
                    gy_nx = clampi(gy_base + r + S::NG, 0, h - 1);
                    at = (unsigned)gy_nx * uw + (unsigned)gxA;
                    if (flow_in) f_nx = fbh_ld_stream(flow_in + at);
                    q_nx = fbh_ld_stream(R0q + at);
                    e_nx = fbh_ld_stream(R0e + at);
                }
                float mm[5];
                fbh_matrix(a, f, gxA, gy, w, h, in, top, bot, mm);
                float* dst = new_half + r * G::PITCH + lxA;
#pragma unroll
                for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
            }
        }
        if (hh == 0) continue;
        float* old_half = ring + ((hh & 1) ^ 1) * G::HALF;
        const int ty = y0 + (hh - 1) * G::TY;
        const int nout = min(G::TY, y1 - ty);
        __syncthreads();
        if constexpr (BFIX) fbr_phase_b<G, NT, FbrPairs<MR, TX>::PAIRS>(old_half, new_half, tid);
        else fbh_phase_b<G, NT, VEC>(old_half, new_half, tid);
        __syncthreads();
        fbh_phase_c<G, TX, NT, VEC, SEG>(old_half, flow_out, tid, x0, ty, nout, w, h, reg, clip);
        __syncthreads();  // the next tile's phase A overwrites the half phase C just read
    }
}

// tuning knobs for experiments (0 = heuristic): rows per CTA, via tf_farneback_tune
extern int g_fbh_rows;
extern int g_fbh_rows_min_px;  // the rows override applies to levels of at least this many pixels (key 2)

template <int MR, int TX, int NT, int WANT, bool VEC, int SEG = 8, bool BFIX = false>
static int fb_launch_half(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                          int clip, cudaStream_t st) {
    using G = FbhGeom<MR, TX, VEC>;
    auto kern = k_fb_iter_half<MR, TX, NT, WANT, VEC, SEG, BFIX>;
    static int resident = 0;
    if (!resident) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, NT, G::SMEM));
        if (resident < 1) return fail(TF_ERR_CUDA, "k_fb_iter_half<%d,%d> does not fit an SM", MR, TX);
    }
    // Chunk height, from sweeps at the 4K pyramid's sizes (tools/fb_level_sweep.py): the kernel runs best with the
    // grid at a whole number k of resident waves (CTAs of one wave are in step, their phases overlap best when a
    // second wave interleaves) and chunks of roughly 100 rows (taller chunks amortise the 2m-row prologue, shorter
    // ones balance the tail): 4K -> 20 chunks of 112 rows (2.03 waves), 1080p -> 20 chunks of 56 rows (1.01).
    int strips = ceil_div(w, TX);
    int sms = sm_count();
    int rows;
    if (g_fbh_rows > 0 && (size_t)w * h >= (size_t)g_fbh_rows_min_px) {
        rows = ceil_div(g_fbh_rows, G::TY) * G::TY;
    } else {
        const double wave = (double)resident * sms;
        int k = std::max(1, (int)lround((double)h * strips / (wave * 100.0)));
        int chunks = std::max(1, (int)lround(k * wave / strips));
        rows = std::max(G::TY, ceil_div(ceil_div(h, chunks), G::TY) * G::TY);
    }
    dim3 grid(strips, ceil_div(h, rows));
    float reg = (float)(1e-3 / (scale * scale));
    const size_t plane = (size_t)w * h;
    kern<<<grid, NT, G::SMEM, st>>>(reinterpret_cast<const float4*>(R0), R0 + 4 * plane,
                                     reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg, rows, clip);
    return TF_OK;
}

// variant 4: scalar B / C, 4 CTAs / SM (64 registers); 5: 128-column strips, 512 threads, 2 CTAs / SM, vector B / C;
// 6: vector B / C, 4 CTAs / SM; 7: scalar B / C, 3 CTAs / SM; 17: as 6 with phase C split into 4-column segments so
// that all 8 warps of a CTA have work in it (the configuration variant 8 uses: 157 vs 169 us at 4K).
// Window radii below 4 (tiles of fewer than 8 rows) stay on the rolling-tile kernel.
template <typename RT>
static int fb_iterate_half(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf,
                           float2* other_buf, bool zero_init, int clip, bool finest, int variant, cudaStream_t st) {
    int m = h->winsize / 2;
    // tiles of fewer than 8 rows and half-precision R storage stay on the rolling-tile kernel
    if (m < 4 || sizeof(RT) != 4)
        return fb_iterate_tile<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip, finest, st);
    const float* R0f = reinterpret_cast<const float*>(R0);
    const float* R1f = reinterpret_cast<const float*>(R1);
    double scale = 1.0 / ((double)h->winsize * h->winsize);
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
#define TF_FBH(MR)                                                                                      \
    case MR:                                                                                            \
        e = variant == 5   ? fb_launch_half<MR, 128, 512, 2, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)  \
            : variant == 6 ? fb_launch_half<MR, 64, 256, 4, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 17 ? fb_launch_half<MR, 64, 256, 4, true, 4>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 23 ? fb_launch_half<MR, 128, 512, 2, true, 4, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 24 ? fb_launch_half<MR, 64, 256, 4, true, 4, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 7 ? fb_launch_half<MR, 64, 256, 3, false>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
                           : fb_launch_half<MR, 64, 256, 4, false>(R0f, R1f, in, dst, L.w, L.h, scale, c, st); \
        break;
                TF_FBH(4) TF_FBH(5) TF_FBH(6) TF_FBH(7) TF_FBH(8) TF_FBH(9) TF_FBH(10) TF_FBH(11) TF_FBH(12)
                TF_FBH(13) TF_FBH(14) TF_FBH(15) TF_FBH(16)
#undef TF_FBH
                default: return fail(TF_ERR_INVALID_ARG, "unsupported window radius %d", m);
            }
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
