// Fused Farneback iteration, "packed half-buffer" kernel (solve variants 25 and 27-31, checked experiments -- not the
// default; what they showed is in DESIGN.md 5c): the arithmetic and the tile walk of
// fb_half.cuh's default configuration (64-column strips, 256 threads, 4 CTAs / SM, halves of 2m matrix rows, 4-column
// phase C) with the instruction count cut where the default spends issue slots on work that is not FP math
// (the kernel's time is its warp-instruction count / ~2.1 per cycle and SM):
//
//  * the matrix channels live in shared memory as PAIRS -- (G11, G12) and (G22, h1) interleaved per pixel as float2
//    planes, h2 as a float plane -- so that every 64-bit shared access of phase B and every half of a 128-bit access
//    of phase C is a naturally aligned register pair and the window sums of two channels are ONE packed instruction
//    (sm_100's FADD2: same flops per clock as FADD, half the issue slots; tools/probe/f32x2_rate.cu).  Each channel is
//    still summed in exactly the order of the default kernel, so the flows are bit-identical to it;
//  * phase A stores 2 x 64 bit + 32 bit per matrix pixel instead of 5 x 32 bit;
//  * the 5-pixel border attenuation sits behind a CTA-uniform test (most CTAs never touch the border), the
//    zero-initial-flow case is a template parameter, the column part of every address is hoisted.
//
// Bank layout: pair planes have a row pitch of PP float2 with PP = 2 (mod 4), the single plane PS floats with PS = 16
// (mod 32); a quarter-warp of phase C is 4 segments x 2 rows and touches 8 distinct 16-byte bank groups in both.  The
// plane stride is padded so that phase B's items (consecutive columns, then the next plane) stay on consecutive banks
// across a plane boundary.
#pragma once
#include "fb_math.cuh"

namespace fbp {

template <int MR>
struct Geom {
    static constexpr int TX = 64, NT = 256;
    static constexpr int TY = 2 * MR;
    static constexpr int WIN = 2 * MR + 1;
    static constexpr int COLS = TX + 2 * MR;
    static constexpr int NG = NT / COLS;                                   // phase-A row groups
    static constexpr int PP = COLS + ((2 - (COLS & 3)) + 4) % 4;           // float2 per pair-plane row, == 2 (mod 4)
    static constexpr int NQ2 = (4 + 2 * MR) / 2;                           // 128-bit words of a pair-plane window
    static constexpr int NQ4 = (4 + 2 * MR + 3) / 4;                       // 128-bit words of a single-plane window
    static constexpr int NEED = (TX - 4 + 4 * NQ4) > COLS ? (TX - 4 + 4 * NQ4) : COLS;
    static constexpr int PS = NEED + ((16 - (NEED & 31)) + 32) % 32;       // floats per single-plane row, == 16 (mod 32)
    static constexpr int PHALF = TY * PP * 2;                              // floats per half of a pair plane
    static constexpr int PRAW = 2 * PHALF;
    static constexpr int PSTR = PRAW + (((2 * COLS - PRAW) % 32) + 32) % 32;  // pair-plane stride in floats
    static constexpr int SHALF = TY * PS;
    static constexpr int TOTAL = 2 * PSTR + 2 * SHALF;
    static constexpr size_t SMEM = (size_t)TOTAL * sizeof(float);
    static constexpr int ITEMS_B = 2 * COLS + COLS / 2;
    static_assert(PSTR % 4 == 0 && PS % 4 == 0 && (2 * PP) % 4 == 0, "rows and planes stay 16-byte aligned");
    static_assert(COLS % 2 == 0, "column pairs in the single plane");
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// a*b - c*d with the rounding error of c*d carried along (Kahan); same operations as fbt_diff_of_products
__device__ __forceinline__ float dop(float a, float b, float c, float d) {
    float cd = c * d;
    float err = fmaf(-c, d, cd);
    float v = fmaf(a, b, -cd);
    return v + err;
}
__device__ __forceinline__ float2 solve(float s0, float s1, float s2, float s3, float s4, float reg) {
    float det = dop(s0, s2, s1, s1) + reg;
    float nx = dop(s0, s4, s1, s3);
    float ny = dop(s2, s3, s1, s4);
    float r = __frcp_rn(det);
    return make_float2(nx * r, ny * r);
}

__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// ---- phase B: vertical window sums of (old half ++ new half) into the old half; an item is one column of a pair
// plane (two channels) or one column pair of the single plane: 64-bit shared accesses and packed adds either way ----
template <int TY, int PITCH2>
__device__ __forceinline__ void phase_b_item(float2* __restrict__ oc, const float2* __restrict__ nc) {
    float2 v[TY];
#pragma unroll
    for (int j = 0; j < TY; j++) v[j] = oc[j * PITCH2];
#pragma unroll
    for (int j = TY - 2; j >= 0; j--) v[j] = __fadd2_rn(v[j], v[j + 1]);
    float2 p = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < TY; j++) {
        float2 nv = nc[j * PITCH2];
        p = j == 0 ? nv : __fadd2_rn(p, nv);
        oc[j * PITCH2] = __fadd2_rn(v[j], p);
    }
}
template <typename G>
__device__ __forceinline__ void phase_b(float* __restrict__ ring, int old_sel, int tid) {
    for (int item = tid; item < G::ITEMS_B; item += G::NT) {
        if (item < 2 * G::COLS) {
            const int plane = item >= G::COLS ? 1 : 0;
            float* base = ring + plane * G::PSTR;
            const int col = item - plane * G::COLS;
            phase_b_item<G::TY, G::PP>(reinterpret_cast<float2*>(base + old_sel * G::PHALF) + col,
                                       reinterpret_cast<const float2*>(base + (old_sel ^ 1) * G::PHALF) + col);
        } else {
            float* base = ring + 2 * G::PSTR;
            const int pair = item - 2 * G::COLS;
            phase_b_item<G::TY, G::PS / 2>(reinterpret_cast<float2*>(base + old_sel * G::SHALF) + pair,
                                           reinterpret_cast<const float2*>(base + (old_sel ^ 1) * G::SHALF) + pair);
        }
    }
}

// ---- phase C: horizontal window sums + 2x2 solve + store; one thread per (row, 4-column segment), a warp = 4 segments
// x 8 rows ----
template <typename G, bool HALF_READS = false>
__device__ __forceinline__ void phase_c(const float* __restrict__ ring, int old_sel, float2* __restrict__ flow_out,
                                        int tid, int x0, int ty, int nout, int w, int h, float reg, int clip) {
    constexpr int NWI = ((G::TY + 7) / 8) * (G::TX / 16);
    const int lane = tid & 31;
    for (int wi = tid >> 5; wi < NWI; wi += G::NT / 32) {
        const int seg = (lane & 3) + 4 * (wi % (G::TX / 16));
        const int row = (lane >> 2) + 8 * (wi / (G::TX / 16));
        if (row >= nout) continue;
        const int y = ty + row;
        const int xg = x0 + seg * 4;
        float s4[4];
        {   // h2: single plane
            const float4* rp = reinterpret_cast<const float4*>(ring + 2 * G::PSTR + old_sel * G::SHALF + row * G::PS + seg * 4);
            float win[4 * G::NQ4];
#pragma unroll
            for (int k = 0; k < G::NQ4; k++) {
                float4 t = rp[k];
                win[4 * k] = t.x; win[4 * k + 1] = t.y; win[4 * k + 2] = t.z; win[4 * k + 3] = t.w;
            }
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < G::WIN; k++) acc += win[k];
            s4[0] = acc;
#pragma unroll
            for (int o = 1; o < 4; o++) {
                acc += win[o + G::WIN - 1] - win[o - 1];
                s4[o] = acc;
            }
        }
        float2 sp[2][4];  // [pair plane][output]: (G11, G12), (G22, h1)
#pragma unroll
        for (int p = 0; p < (HALF_READS ? 1 : 2); p++) {  // HALF_READS: timing experiment (mode 6), wrong flows
            const float4* rp = reinterpret_cast<const float4*>(ring + p * G::PSTR + old_sel * G::PHALF + row * (2 * G::PP) + seg * 8);
            float2 win[2 * G::NQ2];
#pragma unroll
            for (int k = 0; k < G::NQ2; k++) {
                float4 t = rp[k];
                win[2 * k] = make_float2(t.x, t.y);
                win[2 * k + 1] = make_float2(t.z, t.w);
            }
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < G::WIN; k++) acc = __fadd2_rn(acc, win[k]);
            sp[p][0] = acc;
#pragma unroll
            for (int o = 1; o < 4; o++) {
                acc = __fadd2_rn(acc, __fadd2_rn(win[o + G::WIN - 1], neg2(win[o - 1])));
                sp[p][o] = acc;
            }
        }
        if (HALF_READS) {
#pragma unroll
            for (int o = 0; o < 4; o++) sp[1][o] = make_float2(sp[0][o].y, sp[0][o].x);
        }
        float2* dst = flow_out + (size_t)y * w + xg;
        const bool wide = xg + 4 <= w && (w & 1) == 0;
#pragma unroll
        for (int o = 0; o < 4; o += 2) {
            float2 u = solve(sp[0][o].x, sp[0][o].y, sp[1][o].x, sp[1][o].y, s4[o], reg);
            float2 v = solve(sp[0][o + 1].x, sp[0][o + 1].y, sp[1][o + 1].x, sp[1][o + 1].y, s4[o + 1], reg);
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

// ---- phase A: matrix rows of one half = global rows gy_base + [0, TY), clamped; a thread owns one halo'd column and
// every NG-th row.  edge: the CTA touches the 5-pixel attenuation border (CTA-uniform; most CTAs do not) ----
template <typename G, bool HAS_FLOW>
__device__ __forceinline__ void phase_a(float* __restrict__ ring, int new_sel, const float4* __restrict__ R0q,
                                        const float* __restrict__ R0e, const float4* __restrict__ R1q,
                                        const float* __restrict__ R1e, const float2* __restrict__ flow_in, int w, int h,
                                        int gy_base, int gxA, int lxA, int rA, float sxA, bool edge) {
    const unsigned uw = (unsigned)w;
    const float fgx = (float)gxA;
    int r = rA;
    int gy_nx = tf::clampi(gy_base + r, 0, h - 1);
    unsigned at = (unsigned)gy_nx * uw + (unsigned)gxA;
    float2 f_nx = HAS_FLOW ? ld_stream(flow_in + at) : make_float2(0.f, 0.f);
    float4 q_nx = ld_stream(R0q + at);
    float e_nx = ld_stream(R0e + at);
    float2* d01 = reinterpret_cast<float2*>(ring + new_sel * G::PHALF) + r * G::PP + lxA;
    float* d4 = ring + 2 * G::PSTR + new_sel * G::SHALF + r * G::PS + lxA;
#pragma unroll 1
    for (; r < G::TY; r += G::NG) {
        const float2 f = f_nx;
        const float4 a = q_nx;
        const float a4 = e_nx;
        const int gy = gy_nx;
        float fx = fgx + f.x, fy = (float)gy + f.y;
        // cvFloor; the float->int conversion saturates, so absurd displacements land outside the image
        const int x1 = __float2int_rd(fx), yy1 = __float2int_rd(fy);
        const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1);
        float4 t0, t1, b0, b1;
        float te0, te1, be0, be1;
        if (in) {
            const unsigned q = (unsigned)yy1 * uw + (unsigned)x1;
            const float4* pq = R1q + q;
            const float* pe = R1e + q;
            t0 = __ldg(pq); t1 = __ldg(pq + 1);
            te0 = __ldg(pe); te1 = __ldg(pe + 1);
            b0 = __ldg(pq + uw); b1 = __ldg(pq + uw + 1);
            be0 = __ldg(pe + uw); be1 = __ldg(pe + uw + 1);
        }
        // next row's R0 / flow, requested after this row's gather (the gather is the load this row waits for)
        if (r + G::NG < G::TY) {
            gy_nx = tf::clampi(gy_base + r + G::NG, 0, h - 1);
            at = (unsigned)gy_nx * uw + (unsigned)gxA;
            if (HAS_FLOW) f_nx = ld_stream(flow_in + at);
            q_nx = ld_stream(R0q + at);
            e_nx = ld_stream(R0e + at);
        }
        fx -= floorf(fx);
        fy -= floorf(fy);
        // every product / sum below is spelled out (no contraction left to the compiler) in the order the default
        // kernel's fbh_matrix compiles to, so that both kernels round identically
        float r2, r3, r4, r5, r6;
        if (in) {
            const float omx = 1.f - fx, omy = 1.f - fy;
            const float a00 = __fmul_rn(omx, omy), a01 = __fmul_rn(fx, omy), a10 = __fmul_rn(omx, fy), a11 = __fmul_rn(fx, fy);
            r2 = fmaf(a11, b1.x, fmaf(a10, b0.x, fmaf(a01, t1.x, __fmul_rn(a00, t0.x))));
            r3 = fmaf(a11, b1.y, fmaf(a10, b0.y, fmaf(a01, t1.y, __fmul_rn(a00, t0.y))));
            r4 = fmaf(a11, b1.z, fmaf(a10, b0.z, fmaf(a01, t1.z, __fmul_rn(a00, t0.z))));
            r5 = fmaf(a11, b1.w, fmaf(a10, b0.w, fmaf(a01, t1.w, __fmul_rn(a00, t0.w))));
            r6 = fmaf(a11, be1, fmaf(a10, be0, fmaf(a01, te1, __fmul_rn(a00, te0))));
            r4 = __fmul_rn(__fadd_rn(a.z, r4), 0.5f);
            r5 = __fmul_rn(__fadd_rn(a.w, r5), 0.5f);
            r6 = __fmul_rn(__fadd_rn(a4, r6), 0.25f);
        } else {
            r2 = r3 = 0.f;
            r4 = a.z;
            r5 = a.w;
            r6 = __fmul_rn(a4, 0.5f);
        }
        r2 = fmaf(__fsub_rn(a.x, r2), 0.5f, fmaf(r4, f.y, __fmul_rn(r6, f.x)));
        r3 = fmaf(__fsub_rn(a.y, r3), 0.5f, fmaf(r5, f.x, __fmul_rn(r6, f.y)));
        if (edge) {
            if (sxA != 1.f || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
                // same product, same order as fbh_matrix: ((x-low * x-high) * y-low) * y-high
                const float s = __fmul_rn(__fmul_rn(sxA, gy < 5 ? fb_border(gy) : 1.f), gy >= h - 5 ? fb_border(h - gy - 1) : 1.f);
                r2 = __fmul_rn(r2, s); r3 = __fmul_rn(r3, s); r4 = __fmul_rn(r4, s); r5 = __fmul_rn(r5, s); r6 = __fmul_rn(r6, s);
            }
        }
        const float r66 = __fmul_rn(r6, r6);
        d01[0] = make_float2(fmaf(r4, r4, r66), __fmul_rn(__fadd_rn(r4, r5), r6));           // G(1,1), G(1,2)
        d01[G::PSTR / 2] = make_float2(fmaf(r5, r5, r66), fmaf(r2, r4, __fmul_rn(r3, r6)));  // G(2,2), h(1)
        d4[0] = fmaf(r3, r5, __fmul_rn(r2, r6));                                             // h(2)
        d01 += G::NG * G::PP;
        d4 += G::NG * G::PS;
    }
}

// ---- phase A, row pairs (MODE 3): a thread owns one halo'd column and every NG-th PAIR of consecutive rows.  The bottom
// tap row of the first pixel is the top tap row of the second whenever both sample the same integer displacement (the
// common case: the flow is smooth), so a pair fetches three tap rows instead of four -- a quarter fewer L1 wavefronts
// for the gather, which is what the kernel is bound by.  A pixel whose neighbour below samples elsewhere
// fetches its own two rows (per lane).  Same arithmetic per pixel, so the flows stay bit-identical. ----
struct Taps {
    float4 q0, q1;
    float e0, e1;
};
__device__ __forceinline__ Taps load_taps(const float4* __restrict__ Rq, const float* __restrict__ Re, unsigned px) {
    Taps t;
    t.q0 = __ldg(Rq + px);
    t.q1 = __ldg(Rq + px + 1);
    t.e0 = __ldg(Re + px);
    t.e1 = __ldg(Re + px + 1);
    return t;
}
// one matrix pixel from fetched operands: (fx, fy) = sample position before the floor is removed
template <typename G>
__device__ __forceinline__ void matrix_pixel(float4 a, float a4, float2 f, float fx, float fy, bool in, int gy, int h,
                                             float sxA, bool edge, const Taps& T, const Taps& B,
                                             float2* __restrict__ d01, float* __restrict__ d4) {
    fx -= floorf(fx);
    fy -= floorf(fy);
    float r2, r3, r4, r5, r6;
    if (in) {
        const float omx = 1.f - fx, omy = 1.f - fy;
        const float a00 = __fmul_rn(omx, omy), a01 = __fmul_rn(fx, omy), a10 = __fmul_rn(omx, fy), a11 = __fmul_rn(fx, fy);
        r2 = fmaf(a11, B.q1.x, fmaf(a10, B.q0.x, fmaf(a01, T.q1.x, __fmul_rn(a00, T.q0.x))));
        r3 = fmaf(a11, B.q1.y, fmaf(a10, B.q0.y, fmaf(a01, T.q1.y, __fmul_rn(a00, T.q0.y))));
        r4 = fmaf(a11, B.q1.z, fmaf(a10, B.q0.z, fmaf(a01, T.q1.z, __fmul_rn(a00, T.q0.z))));
        r5 = fmaf(a11, B.q1.w, fmaf(a10, B.q0.w, fmaf(a01, T.q1.w, __fmul_rn(a00, T.q0.w))));
        r6 = fmaf(a11, B.e1, fmaf(a10, B.e0, fmaf(a01, T.e1, __fmul_rn(a00, T.e0))));
        r4 = __fmul_rn(__fadd_rn(a.z, r4), 0.5f);
        r5 = __fmul_rn(__fadd_rn(a.w, r5), 0.5f);
        r6 = __fmul_rn(__fadd_rn(a4, r6), 0.25f);
    } else {
        r2 = r3 = 0.f;
        r4 = a.z;
        r5 = a.w;
        r6 = __fmul_rn(a4, 0.5f);
    }
    r2 = fmaf(__fsub_rn(a.x, r2), 0.5f, fmaf(r4, f.y, __fmul_rn(r6, f.x)));
    r3 = fmaf(__fsub_rn(a.y, r3), 0.5f, fmaf(r5, f.x, __fmul_rn(r6, f.y)));
    if (edge) {
        if (sxA != 1.f || (unsigned)(gy - 5) >= (unsigned)(h - 10)) {
            const float s = __fmul_rn(__fmul_rn(sxA, gy < 5 ? fb_border(gy) : 1.f), gy >= h - 5 ? fb_border(h - gy - 1) : 1.f);
            r2 = __fmul_rn(r2, s); r3 = __fmul_rn(r3, s); r4 = __fmul_rn(r4, s); r5 = __fmul_rn(r5, s); r6 = __fmul_rn(r6, s);
        }
    }
    const float r66 = __fmul_rn(r6, r6);
    d01[0] = make_float2(fmaf(r4, r4, r66), __fmul_rn(__fadd_rn(r4, r5), r6));           // G(1,1), G(1,2)
    d01[G::PSTR / 2] = make_float2(fmaf(r5, r5, r66), fmaf(r2, r4, __fmul_rn(r3, r6)));  // G(2,2), h(1)
    d4[0] = fmaf(r3, r5, __fmul_rn(r2, r6));                                             // h(2)
}

template <typename G, bool HAS_FLOW>
__device__ __forceinline__ void phase_a_pair(float* __restrict__ ring, int new_sel, const float4* __restrict__ R0q,
                                             const float* __restrict__ R0e, const float4* __restrict__ R1q,
                                             const float* __restrict__ R1e, const float2* __restrict__ flow_in, int w,
                                             int h, int gy_base, int gxA, int lxA, int rA, float sxA, bool edge) {
    static_assert(G::TY % 2 == 0, "rows come in pairs");
    const unsigned uw = (unsigned)w;
    const float fgx = (float)gxA;
    int r = 2 * rA;  // first row of the pair
    int gy0_nx = tf::clampi(gy_base + r, 0, h - 1), gy1_nx = tf::clampi(gy_base + r + 1, 0, h - 1);
    unsigned at0 = (unsigned)gy0_nx * uw + (unsigned)gxA, at1 = (unsigned)gy1_nx * uw + (unsigned)gxA;
    float2 f0_nx = HAS_FLOW ? ld_stream(flow_in + at0) : make_float2(0.f, 0.f);
    float2 f1_nx = HAS_FLOW ? ld_stream(flow_in + at1) : make_float2(0.f, 0.f);
    float4 q_nx = ld_stream(R0q + at0);
    float e_nx = ld_stream(R0e + at0);
    float2* d01 = reinterpret_cast<float2*>(ring + new_sel * G::PHALF) + r * G::PP + lxA;
    float* d4 = ring + 2 * G::PSTR + new_sel * G::SHALF + r * G::PS + lxA;
#pragma unroll 1
    for (; r < G::TY; r += 2 * G::NG) {
        const float2 f0 = f0_nx, f1 = f1_nx;
        const float4 a0 = q_nx;
        const float a04 = e_nx;
        const int gy0 = gy0_nx, gy1 = gy1_nx;
        const float fx0 = fgx + f0.x, fy0 = (float)gy0 + f0.y;
        const float fx1 = fgx + f1.x, fy1 = (float)gy1 + f1.y;
        // cvFloor; the float->int conversion saturates, so absurd displacements land outside the image
        const int xa = __float2int_rd(fx0), ya = __float2int_rd(fy0);
        const int xb = __float2int_rd(fx1), yb = __float2int_rd(fy1);
        const bool in0 = (unsigned)xa < (unsigned)(w - 1) && (unsigned)ya < (unsigned)(h - 1);
        const bool in1 = (unsigned)xb < (unsigned)(w - 1) && (unsigned)yb < (unsigned)(h - 1);
        const bool share = in0 && xb == xa && yb == ya + 1;
        Taps U, V;  // first pixel: top U, bottom V; second pixel: top V, bottom U
        if (in0) {
            const unsigned q = (unsigned)ya * uw + (unsigned)xa;
            U = load_taps(R1q, R1e, q);
            V = load_taps(R1q, R1e, q + uw);
        }
        // second pixel's R0 (arrives while the first pixel waits for its taps)
        const unsigned atb = (unsigned)gy1 * uw + (unsigned)gxA;
        const float4 a1 = ld_stream(R0q + atb);
        const float a14 = ld_stream(R0e + atb);
        // next pair's flows and first R0
        if (r + 2 * G::NG < G::TY) {
            gy0_nx = tf::clampi(gy_base + r + 2 * G::NG, 0, h - 1);
            gy1_nx = tf::clampi(gy_base + r + 2 * G::NG + 1, 0, h - 1);
            at0 = (unsigned)gy0_nx * uw + (unsigned)gxA;
            at1 = (unsigned)gy1_nx * uw + (unsigned)gxA;
            if (HAS_FLOW) {
                f0_nx = ld_stream(flow_in + at0);
                f1_nx = ld_stream(flow_in + at1);
            }
            q_nx = ld_stream(R0q + at0);
            e_nx = ld_stream(R0e + at0);
        }
        matrix_pixel<G>(a0, a04, f0, fx0, fy0, in0, gy0, h, sxA, edge, U, V, d01, d4);
        if (in1) {
            const unsigned qb = (unsigned)yb * uw + (unsigned)xb;
            if (!share) V = load_taps(R1q, R1e, qb);
            U = load_taps(R1q, R1e, qb + uw);
        }
        matrix_pixel<G>(a1, a14, f1, fx1, fy1, in1, gy1, h, sxA, edge, V, U, d01 + G::PP, d4 + G::PS);
        d01 += 2 * G::NG * G::PP;
        d4 += 2 * G::NG * G::PS;
    }
}

// ---- phase A, row pairs with ONE wait per pair (MODE 7): as phase_a_pair, but the three tap rows of a pair and both
// R0 pixels are requested together, so the longest thread of a tile waits for memory 3 times (7 pairs over 3 row groups)
// instead of 5 (rows one at a time) or 6 (phase_a_pair).  Only the next pair's FLOWS are prefetched (the addresses depend
// on them); a lane whose second pixel samples elsewhere fetches its own two rows after the first pixel (extra wait). ----
template <typename G, bool HAS_FLOW>
__device__ __forceinline__ void phase_a_pair1(float* __restrict__ ring, int new_sel, const float4* __restrict__ R0q,
                                              const float* __restrict__ R0e, const float4* __restrict__ R1q,
                                              const float* __restrict__ R1e, const float2* __restrict__ flow_in, int w,
                                              int h, int gy_base, int gxA, int lxA, int rA, float sxA, bool edge) {
    const unsigned uw = (unsigned)w;
    const float fgx = (float)gxA;
    int r = 2 * rA;
    int gy0 = tf::clampi(gy_base + r, 0, h - 1), gy1 = tf::clampi(gy_base + r + 1, 0, h - 1);
    float2 f0 = make_float2(0.f, 0.f), f1 = f0;
    if (HAS_FLOW) {
        f0 = ld_stream(flow_in + ((unsigned)gy0 * uw + (unsigned)gxA));
        f1 = ld_stream(flow_in + ((unsigned)gy1 * uw + (unsigned)gxA));
    }
    float2* d01 = reinterpret_cast<float2*>(ring + new_sel * G::PHALF) + r * G::PP + lxA;
    float* d4 = ring + 2 * G::PSTR + new_sel * G::SHALF + r * G::PS + lxA;
#pragma unroll 1
    for (; r < G::TY; r += 2 * G::NG) {
        const float fx0 = fgx + f0.x, fy0 = (float)gy0 + f0.y;
        const float fx1 = fgx + f1.x, fy1 = (float)gy1 + f1.y;
        const int xa = __float2int_rd(fx0), ya = __float2int_rd(fy0);
        const int xb = __float2int_rd(fx1), yb = __float2int_rd(fy1);
        const bool in0 = (unsigned)xa < (unsigned)(w - 1) && (unsigned)ya < (unsigned)(h - 1);
        const bool in1 = (unsigned)xb < (unsigned)(w - 1) && (unsigned)yb < (unsigned)(h - 1);
        const bool share = in0 && xb == xa && yb == ya + 1;
        Taps U, V, W;
        if (in0) {
            const unsigned q = (unsigned)ya * uw + (unsigned)xa;
            U = load_taps(R1q, R1e, q);
            V = load_taps(R1q, R1e, q + uw);
        }
        const unsigned qb = (unsigned)yb * uw + (unsigned)xb;
        if (in1 && share) W = load_taps(R1q, R1e, qb + uw);
        const unsigned ata = (unsigned)gy0 * uw + (unsigned)gxA, atb = (unsigned)gy1 * uw + (unsigned)gxA;
        const float4 a0 = ld_stream(R0q + ata), a1 = ld_stream(R0q + atb);
        const float a04 = ld_stream(R0e + ata), a14 = ld_stream(R0e + atb);
        matrix_pixel<G>(a0, a04, f0, fx0, fy0, in0, gy0, h, sxA, edge, U, V, d01, d4);
        if (in1 && !share) {
            V = load_taps(R1q, R1e, qb);
            W = load_taps(R1q, R1e, qb + uw);
        }
        const float2 g1 = f1;
        const int gyb = gy1;
        // the next pair's flows (its addresses need them first), requested once the first pixel's taps are dead: the second
        // pixel's arithmetic and the other warps of the scheduler cover their latency
        if (r + 2 * G::NG < G::TY) {
            gy0 = tf::clampi(gy_base + r + 2 * G::NG, 0, h - 1);
            gy1 = tf::clampi(gy_base + r + 2 * G::NG + 1, 0, h - 1);
            if (HAS_FLOW) {
                f0 = ld_stream(flow_in + ((unsigned)gy0 * uw + (unsigned)gxA));
                f1 = ld_stream(flow_in + ((unsigned)gy1 * uw + (unsigned)gxA));
            }
        }
        matrix_pixel<G>(a1, a14, g1, fx1, fy1, in1, gyb, h, sxA, edge, V, W, d01 + G::PP, d4 + G::PS);
        d01 += 2 * G::NG * G::PP;
        d4 += 2 * G::NG * G::PS;
    }
}

// MODE 1: rows one at a time; 3: row pairs; 4 / 5 / 6: timing experiments (phase A only / phases B + C only / B + C with
// phase C reading 14 instead of 23 128-bit words)
template <int MR, bool HAS_FLOW, int MODE>
__global__ void __launch_bounds__(256, (MODE == 7 ? 3 : 4))   // mode 7 holds three tap rows: 80 registers, 3 CTAs / SM
    k_fb_iter_pack(const float4* __restrict__ R0q, const float* __restrict__ R0e, const float4* __restrict__ R1q,
                   const float* __restrict__ R1e, const float2* __restrict__ flow_in, float2* __restrict__ flow_out,
                   int w, int h, float reg, int rows_per_cta, int clip) {
    using G = Geom<MR>;
    static_assert(G::NT >= G::COLS, "one thread per halo'd column needed");
    extern __shared__ __align__(16) float ring[];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * G::TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;  // halves 0 .. ntiles (half 0 = prologue)

    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < G::NG && ((MODE != 3 && MODE != 7) || 2 * rA < G::TY);
    const int gxA = tf::clampi(x0 - MR + lxA, 0, w - 1);
    // does any matrix pixel of this CTA lie in the 5-pixel attenuation border?  (CTA-uniform)
    const bool edge = x0 - MR < 5 || x0 + G::TX + MR > w - 5 || y0 - MR < 5 || y1 + MR > h - 5;
    const float sxA = __fmul_rn(gxA < 5 ? fb_border(gxA) : 1.f, gxA >= w - 5 ? fb_border(w - gxA - 1) : 1.f);

    for (int hh = 0; hh <= ntiles; hh++) {
        const int new_sel = hh & 1;
        if (activeA && MODE != 5 && MODE != 6) {  // (modes 4 / 5: timing experiments, phase A only / phases B + C only)
            const int gy_base = y0 - MR + hh * G::TY;
            if constexpr (MODE == 7)
                phase_a_pair1<G, HAS_FLOW>(ring, new_sel, R0q, R0e, R1q, R1e, flow_in, w, h, gy_base, gxA, lxA, rA, sxA, edge);
            else if constexpr (MODE == 3)
                phase_a_pair<G, HAS_FLOW>(ring, new_sel, R0q, R0e, R1q, R1e, flow_in, w, h, gy_base, gxA, lxA, rA, sxA, edge);
            else
                phase_a<G, HAS_FLOW>(ring, new_sel, R0q, R0e, R1q, R1e, flow_in, w, h, gy_base, gxA, lxA, rA, sxA, edge);
        }
        if (hh == 0) continue;
        const int old_sel = new_sel ^ 1;
        const int ty = y0 + (hh - 1) * G::TY;
        const int nout = min(G::TY, y1 - ty);
        __syncthreads();
        if (MODE != 4) phase_b<G>(ring, old_sel, tid);
        __syncthreads();
        if (MODE != 4) phase_c<G, MODE == 6>(ring, old_sel, flow_out, tid, x0, ty, nout, w, h, reg, clip);
        __syncthreads();  // the next tile's phase A overwrites the half phase C just read
    }
}

}  // namespace fbp

#ifndef FBP_KERNEL_ONLY
template <int MR, int MODE>
static int fb_launch_pack(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                          int clip, cudaStream_t st) {
    using G = fbp::Geom<MR>;
    auto kern1 = fbp::k_fb_iter_pack<MR, true, MODE>;
    auto kern0 = fbp::k_fb_iter_pack<MR, false, MODE>;
    static int resident = 0;
    if (!resident) {
        TF_CUDA(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        TF_CUDA(cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern1, G::NT, G::SMEM));
        if (resident < 1) return fail(TF_ERR_CUDA, "k_fb_iter_pack<%d> does not fit an SM", MR);
    }
    // chunk height: the rule of fb_launch_half (a whole number of resident waves, roughly 100 rows per chunk)
    int strips = ceil_div(w, G::TX);
    int rows;
    if (g_fbh_rows > 0 && (size_t)w * h >= (size_t)g_fbh_rows_min_px) {
        rows = ceil_div(g_fbh_rows, G::TY) * G::TY;
    } else {
        const double wave = (double)resident * sm_count();
        int k = std::max(1, (int)lround((double)h * strips / (wave * 100.0)));
        int chunks = std::max(1, (int)lround(k * wave / strips));
        rows = std::max(G::TY, ceil_div(ceil_div(h, chunks), G::TY) * G::TY);
    }
    dim3 grid(strips, ceil_div(h, rows));
    float reg = (float)(1e-3 / (scale * scale));
    const size_t plane = (size_t)w * h;
    if (in)
        kern1<<<grid, G::NT, G::SMEM, st>>>(reinterpret_cast<const float4*>(R0), R0 + 4 * plane,
                                            reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg, rows, clip);
    else
        kern0<<<grid, G::NT, G::SMEM, st>>>(reinterpret_cast<const float4*>(R0), R0 + 4 * plane,
                                            reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg, rows, clip);
    return TF_OK;
}

// timing experiments (phase A only / phases B + C only / B + C with fewer phase-C reads: wrong flows by construction),
// window radius 7 only
template <int MR, int MODE>
static int fb_launch_pack_exp(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                              int clip, cudaStream_t st) {
    if constexpr (MR == 7) return fb_launch_pack<7, MODE>(R0, R1, in, dst, w, h, scale, clip, st);
    else return fail(TF_ERR_INVALID_ARG, "variants 28 - 30 are timing experiments for winsize 15");
}

// variants 25 (rows one at a time) / 27 (row pairs) / 28 - 30 (timing experiments); window radii below 4 and half-precision R storage stay on
// the rolling-tile kernel, like fb_iterate_half
template <typename RT>
static int fb_iterate_pack(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf,
                           float2* other_buf, bool zero_init, int clip, bool finest, int variant, cudaStream_t st) {
    int m = h->winsize / 2;
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
#define TF_FBP(MR)                                                                                \
    case MR:                                                                                      \
        e = variant == 27   ? fb_launch_pack<MR, 3>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)    \
            : variant == 28 ? fb_launch_pack_exp<MR, 4>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 29 ? fb_launch_pack_exp<MR, 5>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 30 ? fb_launch_pack_exp<MR, 6>(R0f, R1f, in, dst, L.w, L.h, scale, c, st) \
            : variant == 31 ? fb_launch_pack<MR, 7>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)     \
                            : fb_launch_pack<MR, 1>(R0f, R1f, in, dst, L.w, L.h, scale, c, st);   \
        break;
                TF_FBP(4) TF_FBP(5) TF_FBP(6) TF_FBP(7) TF_FBP(8) TF_FBP(9) TF_FBP(10) TF_FBP(11) TF_FBP(12)
                TF_FBP(13) TF_FBP(14) TF_FBP(15) TF_FBP(16)
#undef TF_FBP
                default: return fail(TF_ERR_INVALID_ARG, "unsupported window radius %d", m);
            }
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
#endif
