// Fused Farneback iteration, "staged" kernel (solve variant 9, experimental): the half-buffer kernel
// (fb_half.cuh) with the R1 operand of phase A staged in shared memory by the bulk-copy engine.
//
// For every half (2m matrix rows x (TX + 2m) columns) one warp issues `cp.async.bulk` row copies of the R1 box
// the half can sample -- its own footprint displaced by the integer flow at its centre, plus a margin of GM
// pixels and the bilinear +1 -- signalled on an mbarrier (expect_tx / complete_tx).  The copies for half k + 1
// are issued right after phase A of half k (one buffer: every thread is done reading it at the barrier that
// follows phase A) and land while phases B and C run.  Phase A then takes its four bilinear taps with 128-bit
// + 32-bit SHARED loads at immediate offsets from one base address: no L1 tag lookups, no misaligned 128-bit
// global gathers (5 wavefronts per warp instead of 4), no exposed DRAM latency.  A tap that leaves the box
// (flow varying by more than GM pixels inside one half) falls back to the global gather of fb_half.cuh, so the
// result does not depend on the staging.
// Requirements (else the caller uses the half-buffer kernel): fp32 R, w % 4 == 0 (16-byte aligned rows of the
// fifth-coefficient plane).
//
// Measured on B200 (4K level 0): 243 us per launch against 174 us for the half-buffer kernel, so this is NOT the
// default.  The box (36 KB) on top of the matrix ring (47 KB) leaves room for 2 CTAs of 8 + 1 warps per SM; with
// one buffer the copies of half k + 1 cannot start before phase A of half k ends, and ncu shows them landing late
// (20 % of the stall samples sit in the mbarrier wait, 25 % on the R0 / flow loads that nothing hides any more,
// issue slots 34 % busy).  A second buffer would cut the residency to one CTA per SM.  Kept as a checked variant:
// it produces bit-identical flows.
#pragma once
#include "fb_half.cuh"

__device__ __forceinline__ uint32_t fbh_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fbs_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the bulk-copy engine
}
__device__ __forceinline__ void fbs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fbs_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int N>
__device__ __forceinline__ void fbs_bar_consumers() {  // named barrier of the N compute threads (not the producer warp)
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fbs_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "FBS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra FBS_DONE;\n"
        "bra FBS_WAIT;\n"
        "FBS_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fbs_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fbs_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MR, int TX, int GM, bool S0 = false>
struct FbsGeom {
    using G = FbhGeom<MR, TX, true>;
    static constexpr int BH = G::TY + 2 * GM + 1;            // box rows
    static constexpr int BW = G::COLS + 2 * GM + 1;          // box columns (quads)
    static constexpr int EP = ((BW + 3) & ~3) + 4;           // fifth-plane row pitch: 4-aligned start + length
    static constexpr size_t RING = G::SMEM;
    static constexpr size_t QBYTES = (size_t)BH * BW * 16;
    static constexpr size_t EBYTES = (size_t)BH * EP * 4;
    // S0: the half's own R0 quads / fifth plane / flow (TY rows x COLS columns) are staged too
    static constexpr int E0P = ((G::COLS + 3) & ~3) + 4;     // R0 fifth-plane row pitch (floats), 4-aligned start
    static constexpr int F0P = ((G::COLS + 1) & ~1) + 2;     // flow row pitch (float2), 2-aligned start
    static constexpr size_t Q0BYTES = S0 ? (size_t)G::TY * G::COLS * 16 : 0;
    static constexpr size_t E0BYTES = S0 ? (size_t)G::TY * E0P * 4 : 0;
    static constexpr size_t F0BYTES = S0 ? (size_t)G::TY * F0P * 8 : 0;
    static constexpr size_t SMEM = RING + QBYTES + EBYTES + Q0BYTES + E0BYTES + F0BYTES + 32;  // + mbarrier, box origin
    static constexpr int FIT = (int)((227 * 1024) / (SMEM + 1024));
    static constexpr int CTAS = FIT < 1 ? 1 : (FIT > 4 ? 4 : FIT);  // resident CTAs the launch bounds ask for
};

// NT compute threads + one producer warp (threads NT .. NT + 31) that only issues the bulk copies: the per-row
// copies are serialised by the uniform datapath, a compute warp doing it would hold up its CTA at the next barrier.
template <int MR, int TX, int NT, int GM, bool S0>
__global__ void __launch_bounds__(NT + 32, (FbsGeom<MR, TX, GM, S0>::CTAS))
    k_fb_iter_stage(const float4* __restrict__ R0q, const float* __restrict__ R0e, const float4* __restrict__ R1q,
                    const float* __restrict__ R1e, const float2* __restrict__ flow_in, float2* __restrict__ flow_out,
                    int w, int h, float reg, int rows_per_cta, int clip) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbsGeom<MR, TX, GM, S0>;
    constexpr int NG = NT / G::COLS;
    static_assert(NT >= G::COLS && B::BH <= 32, "one lane per box row");
    // [5][2 halves][TY][PITCH] | box quads | box fifth plane | R0 quads | R0 fifth plane | flow | mbarrier
    extern __shared__ __align__(16) float ring[];
    float4* boxq = reinterpret_cast<float4*>(reinterpret_cast<char*>(ring) + B::RING);
    float* boxe = reinterpret_cast<float*>(reinterpret_cast<char*>(boxq) + B::QBYTES);
    float4* t0q = reinterpret_cast<float4*>(reinterpret_cast<char*>(boxe) + B::EBYTES);
    float* t0e = reinterpret_cast<float*>(reinterpret_cast<char*>(t0q) + B::Q0BYTES);
    float2* t0f = reinterpret_cast<float2*>(reinterpret_cast<char*>(t0e) + B::E0BYTES);
    // ctl: [0..1] "full" mbarrier, [2..3] "empty" mbarrier, [4] bx0, [5] by0, [6] ex0
    int* ctl = reinterpret_cast<int*>(reinterpret_cast<char*>(t0f) + B::F0BYTES);
    // S0: columns of the half's own operands that exist in the image (quads exactly, the other two widened to
    // 16-byte boundaries: w % 4 == 0)
    const int tq0 = max((int)blockIdx.x * TX - MR, 0), tq1 = min((int)blockIdx.x * TX + TX + MR, w);
    const int te0 = tq0 & ~3, te1 = min((tq1 + 3) & ~3, w);
    const int tf0 = tq0 & ~1, tf1 = min((tq1 + 1) & ~1, w);
    const uint32_t bar = fbh_smem_u32(ctl), bar_empty = fbh_smem_u32(ctl + 2);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;

    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < NG;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);
    const unsigned uw = (unsigned)w;

    // Issue the box of half hh (warp 0 only).  Origin = footprint of the half displaced by the integer flow at its
    // centre, minus the margin.  Lane l copies box row l: the quad row and the fifth-plane row, both clipped to
    // the image (rows outside it are skipped and never sampled: the `in` test of the gather excludes them).
    auto issue_box = [&](int hh) {
        const int lane = tid & 31;
        const int gy_base = y0 - MR + hh * G::TY;
        int ox = 0, oy = 0;
        if (flow_in) {
            const int cy = clampi(gy_base + G::TY / 2, 0, h - 1), cx = clampi(x0 + TX / 2, 0, w - 1);
            const float2 f = __ldg(flow_in + (unsigned)cy * uw + (unsigned)cx);
            // keep the offset finite and modest: beyond this every tap takes the fallback anyway
            ox = __float2int_rd(fminf(fmaxf(f.x, -4096.f), 4096.f));
            oy = __float2int_rd(fminf(fmaxf(f.y, -4096.f), 4096.f));
        }
        const int bx0 = x0 - MR - GM + ox, by0 = gy_base - GM + oy;
        const int cq0 = max(bx0, 0), cq1 = min(bx0 + B::BW, w);               // quad columns [cq0, cq1)
        const int ce0 = max(bx0, 0) & ~3, ce1 = min((bx0 + B::BW + 3) & ~3, w);  // fifth plane, 4-aligned (w % 4 == 0)
        const int ex0 = ce0;                                                    // column of boxe[row][0]
        const int ra = max(by0, 0), rb = min(by0 + B::BH, h);                   // rows [ra, rb)
        const bool any = cq1 > cq0 && rb > ra;
        const uint32_t row0_bytes = S0 ? (uint32_t)((tq1 - tq0) * 16 + (te1 - te0) * 4 + (flow_in ? (tf1 - tf0) * 8 : 0)) : 0u;
        if (lane == 0) {
            ctl[4] = bx0; ctl[5] = by0; ctl[6] = ex0;
            const uint32_t per_row = any ? (uint32_t)((cq1 - cq0) * 16 + (ce1 - ce0) * 4) : 0u;
            fbs_mbar_expect_tx(bar, per_row * (uint32_t)max(rb - ra, 0) + row0_bytes * (uint32_t)G::TY);
        }
        fbs_fence_proxy_async();  // the box was read through the generic proxy until the barrier before this call
        __syncwarp();
        const int row = by0 + lane;
        if (any && lane < B::BH && row >= ra && row < rb) {
            fbs_bulk_g2s(fbh_smem_u32(boxq + lane * B::BW + (cq0 - bx0)), R1q + (unsigned)row * uw + (unsigned)cq0,
                         (uint32_t)(cq1 - cq0) * 16u, bar);
            fbs_bulk_g2s(fbh_smem_u32(boxe + lane * B::EP), R1e + (unsigned)row * uw + (unsigned)ce0,
                         (uint32_t)(ce1 - ce0) * 4u, bar);
        }
        if (S0 && lane < G::TY) {  // the half's own rows (clamped to the image like phase A clamps them)
            const unsigned grow = (unsigned)clampi(gy_base + lane, 0, h - 1) * uw;
            fbs_bulk_g2s(fbh_smem_u32(t0q + lane * G::COLS + (tq0 - (x0 - MR))), R0q + grow + (unsigned)tq0,
                         (uint32_t)(tq1 - tq0) * 16u, bar);
            fbs_bulk_g2s(fbh_smem_u32(t0e + lane * B::E0P), R0e + grow + (unsigned)te0, (uint32_t)(te1 - te0) * 4u, bar);
            if (flow_in)
                fbs_bulk_g2s(fbh_smem_u32(t0f + lane * B::F0P), flow_in + grow + (unsigned)tf0,
                             (uint32_t)(tf1 - tf0) * 8u, bar);
        }
    };

    if (tid == 0) {
        fbs_mbar_init(bar, 1);
        fbs_mbar_init(bar_empty, 1);
    }
    __syncthreads();
    if (tid >= NT) {  // producer warp: box of half hh as soon as the compute warps have released the buffer
        for (int hh = 0; hh <= ntiles; hh++) {
            if (hh > 0) fbs_mbar_wait(bar_empty, (uint32_t)((hh - 1) & 1));
            issue_box(hh);
        }
        return;
    }

    for (int hh = 0; hh <= ntiles; hh++) {
        float* new_half = ring + (hh & 1) * G::HALF;
        // ---- phase A: matrix rows of half hh, R1 taps from the staged box ----
        fbs_mbar_wait(bar, (uint32_t)(hh & 1));
        if (activeA) {
            const int bx0 = ctl[4], by0 = ctl[5], ex0 = ctl[6];
            const int gy_base = y0 - MR + hh * G::TY;
            // R0 / flow of the next TWO rows are in flight (the shared-memory taps leave no other latency to hide behind)
            auto fetch = [&](int rr, int& gyo, float2& fo, float4& qo, float& eo) {
                if (rr < G::TY) {
                    gyo = clampi(gy_base + rr, 0, h - 1);
                    if (S0) {  // staged: row rr of the half, column gxA
                        fo = flow_in ? t0f[rr * B::F0P + (gxA - tf0)] : make_float2(0.f, 0.f);
                        qo = t0q[rr * G::COLS + (gxA - (x0 - MR))];
                        eo = t0e[rr * B::E0P + (gxA - te0)];
                    } else {
                        const unsigned at = (unsigned)gyo * uw + (unsigned)gxA;
                        fo = flow_in ? __ldg(flow_in + at) : make_float2(0.f, 0.f);
                        qo = __ldg(R0q + at);
                        eo = __ldg(R0e + at);
                    }
                }
            };
            int r = rA;
            int gy_1 = 0, gy_2 = 0;
            float2 f_1 = make_float2(0.f, 0.f), f_2 = f_1;
            float4 q_1 = make_float4(0.f, 0.f, 0.f, 0.f), q_2 = q_1;
            float e_1 = 0.f, e_2 = 0.f;
            fetch(r, gy_1, f_1, q_1, e_1);
            fetch(r + NG, gy_2, f_2, q_2, e_2);
#pragma unroll 1
            for (; r < G::TY; r += NG) {
                const float2 f = f_1;
                const float a[5] = {q_1.x, q_1.y, q_1.z, q_1.w, e_1};
                const int gy = gy_1;
                gy_1 = gy_2; f_1 = f_2; q_1 = q_2; e_1 = e_2;
                fetch(r + 2 * NG, gy_2, f_2, q_2, e_2);
                int x1 = __float2int_rd((float)gxA + f.x), yy1 = __float2int_rd((float)gy + f.y);
                const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1);
                FbhTaps top, bot;
                if (in) {
                    const int bx = x1 - bx0, by = yy1 - by0;
                    if ((unsigned)bx < (unsigned)(B::BW - 1) && (unsigned)by < (unsigned)(B::BH - 1)) {
                        const float4* q = boxq + by * B::BW + bx;
                        const float* e = boxe + by * B::EP + (x1 - ex0);
                        top.q0 = q[0]; top.q1 = q[1]; bot.q0 = q[B::BW]; bot.q1 = q[B::BW + 1];
                        top.e0 = e[0]; top.e1 = e[1]; bot.e0 = e[B::EP]; bot.e1 = e[B::EP + 1];
                    } else {  // outside the staged box: gather from global memory
                        unsigned q = (unsigned)yy1 * uw + (unsigned)x1;
                        top = fbh_load_taps(R1q, R1e, q);
                        bot = fbh_load_taps(R1q, R1e, q + uw);
                    }
                }
                float mm[5];
                fbh_matrix(a, f, gxA, gy, w, h, in, top, bot, mm);
                float* dst = new_half + r * G::PITCH + lxA;
#pragma unroll
                for (int c = 0; c < 5; c++) dst[c * G::CHS] = mm[c];
            }
        }
        fbs_bar_consumers<NT>();  // M of this half complete; every thread is done with the box
        if (tid == 0) fbs_mbar_arrive(bar_empty);
        if (hh == 0) continue;
        float* old_half = ring + ((hh & 1) ^ 1) * G::HALF;
        const int ty = y0 + (hh - 1) * G::TY;
        const int nout = min(G::TY, y1 - ty);
        fbh_phase_b<G, NT, true>(old_half, new_half, tid);
        fbs_bar_consumers<NT>();
        fbh_phase_c<G, TX, NT, true, 4>(old_half, flow_out, tid, x0, ty, nout, w, h, reg, clip);
        fbs_bar_consumers<NT>();  // the next half's phase A overwrites the half phase C just read
    }
}

template <int MR, int TX, int NT, int GM, bool S0 = false>
static int fb_launch_stage(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                           int clip, cudaStream_t st) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbsGeom<MR, TX, GM, S0>;
    auto kern = k_fb_iter_stage<MR, TX, NT, GM, S0>;
    static int resident = 0;
    if (!resident) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::SMEM));
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, NT + 32, B::SMEM));
        if (resident < 1) return fail(TF_ERR_CUDA, "k_fb_iter_stage<%d,%d> does not fit an SM", MR, TX);
    }
    int strips = ceil_div(w, TX);
    int sms = sm_count();
    int rows;
    if (g_fbh_rows > 0) {
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
    kern<<<grid, NT + 32, B::SMEM, st>>>(reinterpret_cast<const float4*>(R0), R0 + 4 * plane,
                                         reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg, rows, clip);
    return TF_OK;
}

// variant 9: staged kernel where its requirements hold (fp32 R, default window radius, width % 4 == 0), else
// the half-buffer kernel
template <typename RT>
static int fb_iterate_stage(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf,
                            float2* other_buf, bool zero_init, int clip, bool finest, cudaStream_t st,
                            int kind = 0) {
    int m = h->winsize / 2;
    if (m != 7 || sizeof(RT) != 4 || (L.w & 3) != 0)
        return fb_iterate_half<RT>(h, L, R0, R1, final_buf, other_buf, zero_init, clip, finest, 6, st);
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
            // kind 1 (variant 10): 32-column strips, 192 + 32 threads: ring 29 KB + box 23 KB -> 4 CTAs per SM
            // kind 2 (variant 11): as 1, with the half's own R0 / flow rows staged as well -> 3 CTAs per SM
            e = kind == 2   ? fb_launch_stage<7, 32, 192, 3, true>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : kind == 1 ? fb_launch_stage<7, 32, 192, 3>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                            : fb_launch_stage<7, 64, 256, 3>(R0f, R1f, in, dst, L.w, L.h, scale, c, st);
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
