// Fused Farneback iteration, "ring" kernel (solve variants 18-20; variant 8 uses it on the large levels):
// the half-buffer kernel of fb_half.cuh with the R1 operand of phase A -- the bilinear gather that made that
// kernel L1TEX-bound (46 of its 60 global wavefronts per warp-row, every R1 pixel fetched ~4x through L1) --
// staged in shared memory by the tensor-copy engine as a ROLLING RING OF ROWS:
//
//   * R1 rows enter shared memory once per CTA, in sub-blocks of SB = m rows x BW quads (two
//     cp.async.bulk.tensor.2d copies per sub-block: the interleaved quads as a 2-D map of 8-byte elements -- a
//     row of 88 quads is 176 elements, inside the 256-element box limit that stopped the float32 maps of
//     fb_tma.cuh at 32-column strips -- and the fifth-coefficient plane as float32), each signalled on the
//     slot's mbarrier.  Five slots: a half (2m matrix rows) samples the three sub-blocks 2hh .. 2hh + 2, the two
//     sub-blocks of the NEXT half are already in flight, and the two a half has finished with are re-issued
//     right after its phase A for the half after next -- a copy has a whole half (A + B + C) to land, and there
//     is no producer warp and no operand buffer per half (35 ring rows instead of 2 x 19 box rows);
//   * the ring is displaced by the integer flow at the chunk's centre (ox, oy), with a margin of 4 / 5 columns
//     and (m - 1) / 2 rows around the displaced footprint; a tap outside the resident rows / columns (flow
//     varying by more than the margin inside one chunk) takes the global gather of fb_half.cuh, so the result
//     never depends on the staging -- flows are BIT-IDENTICAL to the half-buffer kernel;
//   * coordinates may leave the image: the copy engine zero-fills, the box is always complete (constant
//     expect_tx) and out-of-image entries are never sampled (the `in` test of the gather excludes them);
//   * a bilinear tap is then one 128-bit + one 32-bit SHARED load (a shared wavefront delivers 128 B, a global
//     one 64 B, and there is no tag stage): 20 wavefronts per warp-row of matrix pixels instead of 46;
//   * R0 and the flow stay coalesced streaming global loads, software-pipelined one row ahead;
//   * phases B and C are those of fb_half.cuh; phase B walks 40 column pairs per channel instead of 39 so that
//     a warp that straddles two channels still touches 32 distinct banks (the channel stride is 16 mod 32).
// Requirements (else the caller uses the half-buffer kernel): fp32 R, window radius 7, w % 4 == 0.
#pragma once
#include "fb_tma.cuh"

template <int MR, int TX>
struct FbrGeom {
    using G = FbhGeom<MR, TX, true>;
    static constexpr int SB = MR;                                  // rows of a sub-block
    static constexpr int NSLOT = 5;
    static constexpr int RR = NSLOT * SB;                          // ring rows
    static constexpr int RES = 3 * SB;                             // rows resident for a half
    static constexpr int GU = (RES - 1 - G::TY) / 2;               // rows above the displaced footprint
    static constexpr int BW = (G::COLS + 1 + 8 + 7) & ~7;          // quads per ring row (multiple of 8: 128-byte slots)
    static constexpr int GL = (BW - G::COLS - 1) / 2;              // columns left of the displaced footprint
    static constexpr int BE = (BW + 3 + 31) & ~31;                 // fifth-plane floats per ring row (multiple of 32)
    static constexpr size_t QBYTES = (size_t)RR * BW * 16;
    static constexpr size_t EBYTES = (size_t)RR * BE * 4;
    static constexpr uint32_t TX_BYTES = (uint32_t)(SB * BW * 16 + SB * BE * 4);
    static constexpr size_t SMEM = QBYTES + EBYTES + G::SMEM + 64; // + the five mbarriers
    static_assert((SB * BW * 16) % 128 == 0 && (SB * BE * 4) % 128 == 0, "tensor copies land on 128-byte boundaries");
    static_assert(QBYTES % 128 == 0 && EBYTES % 128 == 0, "ring sections are 128-byte aligned");
};

struct FbrMaps {
    CUtensorMap r1q, r1e;
};

template <int MR, int TX, int NT>
__global__ void __launch_bounds__(NT, 2)
    k_fb_iter_ring(const __grid_constant__ FbrMaps maps, const float4* __restrict__ R0q, const float* __restrict__ R0e,
                   const float4* __restrict__ R1q, const float* __restrict__ R1e, const float2* __restrict__ flow_in,
                   float2* __restrict__ flow_out, int w, int h, float reg, int rows_per_cta, int clip) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbrGeom<MR, TX>;
    constexpr int NG = NT / G::COLS;
    static_assert(NT >= G::COLS, "one thread per halo'd column needed");
    extern __shared__ __align__(128) float smem_f[];
    char* base = reinterpret_cast<char*>(smem_f);
    const float4* boxq = reinterpret_cast<const float4*>(base);                 // [RR][BW] quads of R1
    const float* boxe = reinterpret_cast<const float*>(base + B::QBYTES);       // [RR][BE] fifth plane of R1
    float* ring = reinterpret_cast<float*>(base + B::QBYTES + B::EBYTES);       // [5][2 halves][TY][PITCH] matrix ring
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + B::QBYTES + B::EBYTES + G::SMEM);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;  // halves 0 .. ntiles (half 0 = prologue)
    const int nsb = 2 * ntiles + 3;                    // sub-blocks 0 .. 2 ntiles + 2
    const unsigned uw = (unsigned)w;

    // ring origin: the chunk's footprint displaced by the integer flow at its centre, minus the margins
    int ox = 0, oy = 0;
    if (flow_in) {
        const int cy = clampi((y0 + y1) >> 1, 0, h - 1), cx = clampi(x0 + TX / 2, 0, w - 1);
        const float2 fc = __ldg(flow_in + (unsigned)cy * uw + (unsigned)cx);
        ox = __float2int_rd(fminf(fmaxf(fc.x, -4096.f), 4096.f));
        oy = __float2int_rd(fminf(fmaxf(fc.y, -4096.f), 4096.f));
    }
    const int bx0 = x0 - MR - B::GL + ox;      // image column of ring column 0
    const int ex0 = bx0 & ~3;                  // the float32 map starts on a 16-byte boundary
    const int ry0 = y0 - MR - B::GU + oy;      // image row of sub-block 0, row 0

    auto issue = [&](int sb) {                 // one thread: the two tensor copies of sub-block sb into slot sb % 5
        const int s = sb % B::NSLOT;
        const uint32_t bar = fbh_smem_u32(bars + s);
        fbs_mbar_expect_tx(bar, B::TX_BYTES);
        fbm_tensor_g2s(fbh_smem_u32(boxq + s * B::SB * B::BW), &maps.r1q, 2 * bx0, ry0 + B::SB * sb, bar);
        fbm_tensor_g2s(fbh_smem_u32(boxe + s * B::SB * B::BE), &maps.r1e, ex0, ry0 + B::SB * sb, bar);
    };
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.r1q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.r1e)) : "memory");
        for (int s = 0; s < B::NSLOT; s++) fbs_mbar_init(fbh_smem_u32(bars + s), 1);
    }
    __syncthreads();   // the mbarriers exist; the copies below are awaited through them, not through a CTA barrier
    if (tid == 0)
        for (int sb = 0; sb < B::NSLOT && sb < nsb; sb++) issue(sb);

    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < NG;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);

    for (int hh = 0; hh <= ntiles; hh++) {
        float* new_half = ring + (hh & 1) * G::HALF;
        const int gy_base = y0 - MR + hh * G::TY;
        // first row's R0 / flow: requested before the wait for the ring
        int r = rA;
        int gy_nx = clampi(gy_base + r, 0, h - 1);
        unsigned at = (unsigned)gy_nx * uw + (unsigned)gxA;
        float2 f_nx = make_float2(0.f, 0.f);
        float4 q_nx = make_float4(0.f, 0.f, 0.f, 0.f);
        float e_nx = 0.f;
        if (activeA) {
            if (flow_in) f_nx = fbh_ld_stream(flow_in + at);
            q_nx = fbh_ld_stream(R0q + at);
            e_nx = fbh_ld_stream(R0e + at);
        }
        // the three sub-blocks this half samples (2hh was awaited by the previous half)
        for (int sb = hh == 0 ? 0 : 2 * hh + 1; sb <= 2 * hh + 2; sb++)
            if (!fbm_mbar_wait(fbh_smem_u32(bars + sb % B::NSLOT), (uint32_t)((sb / B::NSLOT) & 1))) return;
        // ---- phase A: matrix rows of half hh = global rows y0 - m + hh * TY + [0, TY), clamped ----
        if (activeA) {
            const int byh = ry0 + G::TY * hh;                       // image row of the half's first resident ring row
            const int rbase = B::SB * ((2 * hh) % B::NSLOT);        // its ring row
#pragma unroll 1
            for (; r < G::TY; r += NG) {
                const float2 f = f_nx;
                const float a[5] = {q_nx.x, q_nx.y, q_nx.z, q_nx.w, e_nx};
                const int gy = gy_nx;
                // next row's R0 / flow first: they are the only loads of the fast path that leave the SM
                if (r + NG < G::TY) {
                    gy_nx = clampi(gy_base + r + NG, 0, h - 1);
                    at = (unsigned)gy_nx * uw + (unsigned)gxA;
                    if (flow_in) f_nx = fbh_ld_stream(flow_in + at);
                    q_nx = fbh_ld_stream(R0q + at);
                    e_nx = fbh_ld_stream(R0e + at);
                }
                // cvFloor; the float->int conversion saturates, so absurd displacements land outside the image
                const int x1 = __float2int_rd((float)gxA + f.x), yy1 = __float2int_rd((float)gy + f.y);
                const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1);
                const int bx = x1 - bx0, by = yy1 - byh;
                const bool staged = (unsigned)bx < (unsigned)(B::BW - 1) && (unsigned)by < (unsigned)(B::RES - 1);
                int r0 = by + rbase;
                r0 -= r0 >= B::RR ? B::RR : 0;
                const int r1 = r0 + 1 == B::RR ? 0 : r0 + 1;
                FbhTaps top, bot;
                float mm[5];
                // Warp-uniform choice: the common path holds shared loads only, so the wait for the taps is not tied
                // to the scoreboard of the global loads just issued for the next row.
                if (__ballot_sync(__activemask(), in && !staged) == 0) {
                    if (in) {
                        const float4* qt = boxq + r0 * B::BW + bx;
                        const float4* qb = boxq + r1 * B::BW + bx;
                        const float* et = boxe + r0 * B::BE + (x1 - ex0);
                        const float* eb = boxe + r1 * B::BE + (x1 - ex0);
                        top.q0 = qt[0]; top.q1 = qt[1]; bot.q0 = qb[0]; bot.q1 = qb[1];
                        top.e0 = et[0]; top.e1 = et[1]; bot.e0 = eb[0]; bot.e1 = eb[1];
                    }
                    fbh_matrix(a, f, gxA, gy, w, h, in, top, bot, mm);
                } else {
                    if (in) {
                        if (staged) {
                            const float4* qt = boxq + r0 * B::BW + bx;
                            const float4* qb = boxq + r1 * B::BW + bx;
                            const float* et = boxe + r0 * B::BE + (x1 - ex0);
                            const float* eb = boxe + r1 * B::BE + (x1 - ex0);
                            top.q0 = qt[0]; top.q1 = qt[1]; bot.q0 = qb[0]; bot.q1 = qb[1];
                            top.e0 = et[0]; top.e1 = et[1]; bot.e0 = eb[0]; bot.e1 = eb[1];
                        } else {  // outside the resident part of the ring: gather from global memory
                            const unsigned q = (unsigned)yy1 * uw + (unsigned)x1;
                            top = fbh_load_taps(R1q, R1e, q);
                            bot = fbh_load_taps(R1q, R1e, q + uw);
                        }
                    }
                    fbh_matrix(a, f, gxA, gy, w, h, in, top, bot, mm);
                }
                float* dst = new_half + r * G::PITCH + lxA;
#pragma unroll
                for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
            }
        }
        __syncthreads();  // M of this half complete; sub-blocks 2hh and 2hh + 1 are dead
        if (tid == 0) {
            fbs_fence_proxy_async();  // the slots were read through the generic proxy until the barrier above
            if (2 * hh + 5 < nsb) issue(2 * hh + 5);
            if (2 * hh + 6 < nsb) issue(2 * hh + 6);
        }
        if (hh == 0) continue;
        float* old_half = ring + ((hh & 1) ^ 1) * G::HALF;
        const int ty = y0 + (hh - 1) * G::TY;
        const int nout = min(G::TY, y1 - ty);
        fbr_phase_b<G, NT, FbrPairs<MR, TX>::PAIRS>(old_half, new_half, tid);
        __syncthreads();
        fbh_phase_c<G, TX, NT, true, 4>(old_half, flow_out, tid, x0, ty, nout, w, h, reg, clip);
        __syncthreads();  // the next half's phase A overwrites the half phase C just read
    }
}

// ---- ring kernel, ROW PAIRS (variants 21 / 22) -------------------------------------------------------------------
// Phase A with one thread per (halo'd column, PAIR of consecutive matrix rows): when the two pixels sample the same
// columns one row apart -- the common case, the flow is smooth -- the lower tap row of the first pixel is the upper
// tap row of the second, so a pair costs 3 tap rows (6 x 128-bit + 6 x 32-bit shared loads) instead of 4, and its
// address arithmetic, bounds tests and loop overhead are paid once.  The R0 / flow operands of the NEXT pair are
// requested before the current pair's taps (two rows of look-ahead), and those of a half's first pair before phases B
// and C of the previous half, so that no global load is waited for inside phase A.  7 row pairs per half.
struct FbrOps {
    float2 f;
    float4 q;
    float e;
};
__device__ __forceinline__ FbrOps fbr_load_ops(const float4* __restrict__ R0q, const float* __restrict__ R0e,
                                               const float2* __restrict__ flow_in, unsigned at) {
    FbrOps o;
    o.f = flow_in ? fbh_ld_stream(flow_in + at) : make_float2(0.f, 0.f);
    o.q = fbh_ld_stream(R0q + at);
    o.e = fbh_ld_stream(R0e + at);
    return o;
}

template <int MR, int TX, int NT>
__global__ void __launch_bounds__(NT, 2)
    k_fb_iter_ring2(const __grid_constant__ FbrMaps maps, const float4* __restrict__ R0q, const float* __restrict__ R0e,
                    const float4* __restrict__ R1q, const float* __restrict__ R1e, const float2* __restrict__ flow_in,
                    float2* __restrict__ flow_out, int w, int h, float reg, int rows_per_cta, int clip) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbrGeom<MR, TX>;
    constexpr int NG = NT / G::COLS;
    constexpr int NP = G::TY / 2;  // row pairs per half
    static_assert(NT >= G::COLS, "one thread per halo'd column needed");
    extern __shared__ __align__(128) float smem_f[];
    char* base = reinterpret_cast<char*>(smem_f);
    const float4* boxq = reinterpret_cast<const float4*>(base);
    const float* boxe = reinterpret_cast<const float*>(base + B::QBYTES);
    float* ring = reinterpret_cast<float*>(base + B::QBYTES + B::EBYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + B::QBYTES + B::EBYTES + G::SMEM);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;
    const int nsb = 2 * ntiles + 3;
    const unsigned uw = (unsigned)w;

    int ox = 0, oy = 0;
    if (flow_in) {
        const int cy = clampi((y0 + y1) >> 1, 0, h - 1), cx = clampi(x0 + TX / 2, 0, w - 1);
        const float2 fc = __ldg(flow_in + (unsigned)cy * uw + (unsigned)cx);
        ox = __float2int_rd(fminf(fmaxf(fc.x, -4096.f), 4096.f));
        oy = __float2int_rd(fminf(fmaxf(fc.y, -4096.f), 4096.f));
    }
    const int bx0 = x0 - MR - B::GL + ox;
    const int ex0 = bx0 & ~3;
    const int ry0 = y0 - MR - B::GU + oy;

    auto issue = [&](int sb) {
        const int s = sb % B::NSLOT;
        const uint32_t bar = fbh_smem_u32(bars + s);
        fbs_mbar_expect_tx(bar, B::TX_BYTES);
        fbm_tensor_g2s(fbh_smem_u32(boxq + s * B::SB * B::BW), &maps.r1q, 2 * bx0, ry0 + B::SB * sb, bar);
        fbm_tensor_g2s(fbh_smem_u32(boxe + s * B::SB * B::BE), &maps.r1e, ex0, ry0 + B::SB * sb, bar);
    };
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.r1q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.r1e)) : "memory");
        for (int s = 0; s < B::NSLOT; s++) fbs_mbar_init(fbh_smem_u32(bars + s), 1);
    }
    __syncthreads();
    if (tid == 0)
        for (int sb = 0; sb < B::NSLOT && sb < nsb; sb++) issue(sb);

    const int lxA = tid % G::COLS, rg = tid / G::COLS;
    const bool activeA = rg < NG;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);
    const float fgx = (float)gxA;

    // operands of the pair (rows 2 pr, 2 pr + 1) of the half whose first matrix row is gyb
    auto row_at = [&](int gyb, int r) { return (unsigned)clampi(gyb + r, 0, h - 1) * uw + (unsigned)gxA; };
    FbrOps na, nb;  // the NEXT pair's operands (in flight)
    na.f = nb.f = make_float2(0.f, 0.f);
    na.q = nb.q = make_float4(0.f, 0.f, 0.f, 0.f);
    na.e = nb.e = 0.f;
    if (activeA) {
        na = fbr_load_ops(R0q, R0e, flow_in, row_at(y0 - MR, 2 * rg));
        nb = fbr_load_ops(R0q, R0e, flow_in, row_at(y0 - MR, 2 * rg + 1));
    }

    for (int hh = 0; hh <= ntiles; hh++) {
        float* new_half = ring + (hh & 1) * G::HALF;
        const int gy_base = y0 - MR + hh * G::TY;
        for (int sb = hh == 0 ? 0 : 2 * hh + 1; sb <= 2 * hh + 2; sb++)
            if (!fbm_mbar_wait(fbh_smem_u32(bars + sb % B::NSLOT), (uint32_t)((sb / B::NSLOT) & 1))) return;
        if (activeA) {
            const int byh = ry0 + G::TY * hh;
            const int rbase = B::SB * ((2 * hh) % B::NSLOT);
#pragma unroll 1
            for (int pr = rg; pr < NP; pr += NG) {
                const FbrOps oa = na, ob = nb;
                const int gya = clampi(gy_base + 2 * pr, 0, h - 1), gyb = clampi(gy_base + 2 * pr + 1, 0, h - 1);
                // next pair: this half's, or the first pair of the next half (it then waits through phases B and C)
                {
                    const bool more = pr + NG < NP;
                    const int nbase = more ? gy_base : gy_base + G::TY;
                    const int npr = more ? pr + NG : rg;
                    if (more || hh < ntiles) {
                        na = fbr_load_ops(R0q, R0e, flow_in, row_at(nbase, 2 * npr));
                        nb = fbr_load_ops(R0q, R0e, flow_in, row_at(nbase, 2 * npr + 1));
                    }
                }
                const int x1a = __float2int_rd(fgx + oa.f.x), y1a = __float2int_rd((float)gya + oa.f.y);
                const int x1b = __float2int_rd(fgx + ob.f.x), y1b = __float2int_rd((float)gyb + ob.f.y);
                const bool ina = (unsigned)x1a < (unsigned)(w - 1) && (unsigned)y1a < (unsigned)(h - 1);
                const bool inb = (unsigned)x1b < (unsigned)(w - 1) && (unsigned)y1b < (unsigned)(h - 1);
                const int bxa = x1a - bx0, bya = y1a - byh, bxb = x1b - bx0, byb = y1b - byh;
                const bool sta = (unsigned)bxa < (unsigned)(B::BW - 1) && (unsigned)bya < (unsigned)(B::RES - 1);
                const bool stb = (unsigned)bxb < (unsigned)(B::BW - 1) && (unsigned)byb < (unsigned)(B::RES - 1);
                const float aa[5] = {oa.q.x, oa.q.y, oa.q.z, oa.q.w, oa.e};
                const float ab[5] = {ob.q.x, ob.q.y, ob.q.z, ob.q.w, ob.e};
                float* dst = new_half + (2 * pr) * G::PITCH + lxA;
                float mm[5];
                auto ring_row = [&](int by) {
                    int r = by + rbase;
                    return r - (r >= B::RR ? B::RR : 0);
                };
                auto next_row = [&](int r) { return r + 1 == B::RR ? 0 : r + 1; };
                auto ld = [&](int r, int bx, int x1) {
                    FbhTaps t;
                    const float4* q = boxq + r * B::BW + bx;
                    const float* e = boxe + r * B::BE + (x1 - ex0);
                    t.q0 = q[0]; t.q1 = q[1]; t.e0 = e[0]; t.e1 = e[1];
                    return t;
                };
                if (__ballot_sync(__activemask(), (ina && !sta) || (inb && !stb)) == 0) {
                    // shared-memory taps only
                    FbhTaps top, bot;
                    const int r0a = ring_row(bya), r1a = next_row(r0a);
                    if (ina) {
                        top = ld(r0a, bxa, x1a);
                        bot = ld(r1a, bxa, x1a);
                    }
                    fbh_matrix(aa, oa.f, gxA, gya, w, h, ina, top, bot, mm);
#pragma unroll
                    for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
                    if (inb) {
                        if (ina && x1b == x1a && y1b == y1a + 1) {
                            top = bot;                         // the row below the first pixel's: already here
                            bot = ld(next_row(r1a), bxb, x1b);
                        } else {
                            const int r0b = ring_row(byb);
                            top = ld(r0b, bxb, x1b);
                            bot = ld(next_row(r0b), bxb, x1b);
                        }
                    }
                    fbh_matrix(ab, ob.f, gxA, gyb, w, h, inb, top, bot, mm);
#pragma unroll
                    for (int c = 0; c < 5; c++) dst[G::PITCH + c * G::CHS] = mm[c];
                } else {
                    // a tap of this warp lies outside the resident ring: per-pixel choice, global gather where needed
                    FbhTaps top, bot;
                    if (ina) {
                        if (sta) {
                            const int r0a = ring_row(bya);
                            top = ld(r0a, bxa, x1a);
                            bot = ld(next_row(r0a), bxa, x1a);
                        } else {
                            const unsigned q = (unsigned)y1a * uw + (unsigned)x1a;
                            top = fbh_load_taps(R1q, R1e, q);
                            bot = fbh_load_taps(R1q, R1e, q + uw);
                        }
                    }
                    fbh_matrix(aa, oa.f, gxA, gya, w, h, ina, top, bot, mm);
#pragma unroll
                    for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
                    if (inb) {
                        if (stb) {
                            const int r0b = ring_row(byb);
                            top = ld(r0b, bxb, x1b);
                            bot = ld(next_row(r0b), bxb, x1b);
                        } else {
                            const unsigned q = (unsigned)y1b * uw + (unsigned)x1b;
                            top = fbh_load_taps(R1q, R1e, q);
                            bot = fbh_load_taps(R1q, R1e, q + uw);
                        }
                    }
                    fbh_matrix(ab, ob.f, gxA, gyb, w, h, inb, top, bot, mm);
#pragma unroll
                    for (int c = 0; c < 5; c++) dst[G::PITCH + c * G::CHS] = mm[c];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            fbs_fence_proxy_async();
            if (2 * hh + 5 < nsb) issue(2 * hh + 5);
            if (2 * hh + 6 < nsb) issue(2 * hh + 6);
        }
        if (hh == 0) continue;
        float* old_half = ring + ((hh & 1) ^ 1) * G::HALF;
        const int ty = y0 + (hh - 1) * G::TY;
        const int nout = min(G::TY, y1 - ty);
        fbr_phase_b<G, NT, FbrPairs<MR, TX>::PAIRS>(old_half, new_half, tid);
        __syncthreads();
        fbh_phase_c<G, TX, NT, true, 4>(old_half, flow_out, tid, x0, ty, nout, w, h, reg, clip);
        __syncthreads();
    }
}

// ---- host side ----
static int fbr_map2d(CUtensorMap* m, const void* basep, CUtensorMapDataType dtype, uint32_t elem_bytes, uint64_t inner,
                     uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    static fbm_encode_fn encode = nullptr;
    if (!encode) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        TF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
        TF_REQUIRE(sym && q == cudaDriverEntryPointSuccess, TF_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<fbm_encode_fn>(sym);
    }
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstride[1] = {inner * elem_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, dtype, 2, const_cast<void*>(basep), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TF_OK;
}

extern int g_fbr_rows;  // rows per CTA of the ring kernel (0 = heuristic), tf_farneback_tune key 3

template <int MR, int TX, int NT, bool PAIRS = false>
static int fb_launch_ring(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                          int clip, cudaStream_t st) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbrGeom<MR, TX>;
    void (*kern)(const FbrMaps, const float4*, const float*, const float4*, const float*, const float2*, float2*, int, int,
                 float, int, int);
    if constexpr (PAIRS) kern = k_fb_iter_ring2<MR, TX, NT>;
    else kern = k_fb_iter_ring<MR, TX, NT>;
    static int resident = 0;
    if (!resident) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::SMEM));
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, NT, B::SMEM));
        if (resident < 1) return fail(TF_ERR_CUDA, "k_fb_iter_ring<%d,%d> does not fit an SM", MR, TX);
    }
    const size_t plane = (size_t)w * h;
    FbrMaps maps;
    memset(&maps, 0, sizeof(maps));
    // quads as 8-byte elements: 2w per row, a ring row of BW quads = 2 BW elements (<= 256)
    static_assert(2 * B::BW <= 256 && B::BE <= 256, "tensor-map box limit");
    if (int e = fbr_map2d(&maps.r1q, R1, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 2ull * w, h, 2 * B::BW, B::SB)) return e;
    if (int e = fbr_map2d(&maps.r1e, R1 + 4 * plane, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w, h, B::BE, B::SB)) return e;
    // Chunk height: a whole number of resident waves (the short last chunk of a column of chunks fills the gaps)
    // and chunks tall enough to amortise the 2m-row prologue and the ring's 3m extra rows.
    int strips = ceil_div(w, TX);
    int sms = sm_count();
    int rows;
    if (g_fbr_rows > 0) {
        rows = ceil_div(g_fbr_rows, G::TY) * G::TY;
    } else {
        const double wave = (double)resident * sms;
        int k = std::max(1, (int)lround((double)h * strips / (wave * 110.0)));
        int chunks = std::max(1, (int)lround(k * wave / strips));
        rows = std::max(G::TY, ceil_div(ceil_div(h, chunks), G::TY) * G::TY);
    }
    dim3 grid(strips, ceil_div(h, rows));
    float reg = (float)(1e-3 / (scale * scale));
    kern<<<grid, NT, B::SMEM, st>>>(maps, reinterpret_cast<const float4*>(R0), R0 + 4 * plane,
                                    reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg, rows, clip);
    return TF_OK;
}

// variants 18 / 19 / 20: 256 / 320 / 384 threads per CTA (3 / 4 / 4 phase-A row groups), 2 CTAs per SM;
// `threads` < 0: the row-pair kernel (variants 21 / 22: 256 / 320 threads)
template <typename RT>
static int fb_iterate_ring(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf, float2* other_buf,
                           bool zero_init, int clip, bool finest, cudaStream_t st, int threads) {
    int m = h->winsize / 2;
    if (m != 7 || sizeof(RT) != 4 || (L.w & 3) != 0)
        return fb_iterate_half<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip, finest, 17, st);
    const float* R0f = reinterpret_cast<const float*>(R0);
    const float* R1f = reinterpret_cast<const float*>(R1);
    double scale = 1.0 / ((double)h->winsize * h->winsize);
    int T = h->iterations;
    for (int it = 0; it < T; it++) {
        float2* dst = ((T - 1 - it) & 1) ? other_buf : final_buf;
        float2* src = ((T - 1 - it) & 1) ? final_buf : other_buf;
        const float2* in = (it == 0 && zero_init) ? nullptr : src;
        int c = clip && it + 1 == T;
        int e;
        {
            ScopedKernelTimer timer(finest ? TFK_FB_ITER_FINEST : -1, st);
            e = threads == 384   ? fb_launch_ring<7, 64, 384>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : threads == 320 ? fb_launch_ring<7, 64, 320>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : threads == -256 ? fb_launch_ring<7, 64, 256, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : threads == -320 ? fb_launch_ring<7, 64, 320, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                                 : fb_launch_ring<7, 64, 256>(R0f, R1f, in, dst, L.w, L.h, scale, c, st);
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
