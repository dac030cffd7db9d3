// Fused Farneback iteration, "tensor-copy" kernel (solve variant 12, experimental): the staged kernel of
// fb_stage.cuh with EVERY operand of phase A -- the R1 box, the half's own R0 quads / fifth plane and its flow --
// brought into shared memory by 2-D bulk TENSOR copies (cp.async.bulk.tensor.2d, five instructions per half issued
// by one elected lane) instead of ~84 per-row bulk copies.  Why: with per-row copies the refill of the single operand
// buffer takes longer than phases B + C of the same half (variant 11: 5.7 us per 32-column half against ~2.4 us of
// instruction issue), and a CTA's halves are strictly sequential.  The tensor maps describe each operand as a 2-D
// float32 array (quads: 4w floats per row); coordinates may leave the image, the copy engine zero-fills there, so
// the box is always complete (constant expect_tx) and needs no clipping:
//   * R1 rows / columns outside the image are never sampled (the `in` test of the gather excludes them);
//   * matrix rows / columns outside the image take the CLAMPED pixel's operands (cv2 replicates the border), which
//     lie inside the tile because the tile's origin is clamped so that it always intersects the image.
// Every tile starts on a 128-byte boundary of shared memory (tensor copies require it).
// Requirements (else the caller uses the half-buffer kernel): fp32 R, window radius 7, w % 4 == 0.
#pragma once
#include <cuda.h>
#include "fb_stage.cuh"

template <int MR, int TX, int GM, int NBUF = 1>
struct FbmGeom {
    using G = FbhGeom<MR, TX, true>;
    static constexpr int BH = G::TY + 2 * GM + 1;           // R1 box rows
    static constexpr int BW = G::COLS + 2 * GM + 1;         // R1 box columns (quads)
    // a tensor copy must start on a 16-byte boundary of the innermost dimension (probed: tools/probe/tma_probe.cu), so
    // the float / float2 tiles start at the column rounded down to 4 / 2 and are up to 3 / 1 columns wider
    static constexpr int BE = (BW + 6) & ~3;                // fifth-plane box width (floats)
    static constexpr int TW = G::COLS;                      // own-tile columns
    static constexpr int TE = (TW + 6) & ~3;                // own fifth-plane tile width (floats)
    static constexpr int TF = (TW + 2) & ~1;                // own flow tile width (float2)
    static constexpr size_t al(size_t v) { return (v + 127) & ~(size_t)127; }
    static constexpr size_t RING = al(G::SMEM);
    static constexpr size_t Q1 = al((size_t)BH * BW * 16), E1 = al((size_t)BH * BE * 4);
    static constexpr size_t Q0 = al((size_t)G::TY * TW * 16), E0 = al((size_t)G::TY * TE * 4), F0 = al((size_t)G::TY * TF * 8);
    static constexpr uint32_t TX_BYTES_NOFLOW = (uint32_t)(BH * BW * 16 + BH * BE * 4 + G::TY * TW * 16 + G::TY * TE * 4);
    static constexpr uint32_t TX_BYTES_FLOW = TX_BYTES_NOFLOW + (uint32_t)(G::TY * TF * 8);
    static constexpr size_t OPB = Q1 + E1 + Q0 + E0 + F0;                // one set of operand buffers
    static constexpr size_t SMEM = RING + NBUF * OPB + 128;              // + mbarriers, box origins
    static constexpr int FIT = (int)((227 * 1024) / (SMEM + 1024));
    static constexpr int CTAS = FIT < 1 ? 1 : (FIT > 4 ? 4 : FIT);
};

struct FbmMaps {
    CUtensorMap r1q, r1e, r0q, r0e, flow;
};

__device__ __forceinline__ void fbm_tensor_g2s(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

// mbarrier wait that traps instead of spinning forever (a byte-count mismatch must not hang the device)
__device__ __forceinline__ bool fbm_mbar_wait(uint32_t bar, uint32_t parity) {
    for (unsigned tries = 0;; tries++) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return true;
        if (tries > (1u << 20)) {
            if ((threadIdx.x & 31) == 0)
                printf("k_fb_iter_tma: mbarrier %u parity %u never completed (block %d,%d thread %d)\n", bar, parity,
                       (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
            return false;
        }
    }
}

template <int MR, int TX, int NT, int GM, int NBUF>
__global__ void __launch_bounds__(NT + 32, (FbmGeom<MR, TX, GM, NBUF>::CTAS))
    k_fb_iter_tma(const __grid_constant__ FbmMaps maps, const float4* __restrict__ R1q, const float* __restrict__ R1e,
                  const float2* __restrict__ flow_in, float2* __restrict__ flow_out, int w, int h, float reg,
                  int rows_per_cta, int clip) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbmGeom<MR, TX, GM, NBUF>;
    constexpr int NG = NT / G::COLS;
    static_assert(NT >= G::COLS, "one thread per halo'd column needed");
    static_assert(NBUF == 1 || NBUF == 2, "one or two sets of operand buffers");
    extern __shared__ __align__(128) float ring[];
    char* base = reinterpret_cast<char*>(ring);
    // operand set s (half hh uses set hh % NBUF): R1 box quads | fifth plane | own quads | own fifth plane | own flow
    auto set_base = [&](int s) { return base + B::RING + (size_t)s * B::OPB; };
    // ctl: "full" mbarrier of set s at [4 s], "empty" at [4 s + 2]; origins (bx0, by0, ty0) of set s at [8 + 4 s ..]
    int* ctl = reinterpret_cast<int*>(base + B::RING + NBUF * B::OPB);
    auto bar_full = [&](int s) { return fbh_smem_u32(ctl + 4 * s); };
    auto bar_empty = [&](int s) { return fbh_smem_u32(ctl + 4 * s + 2); };
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int ntiles = (y1 - y0 + G::TY - 1) / G::TY;
    const int tx0 = x0 - MR;  // own-tile origin column (x0 < w: the tile always holds the clamped columns)

    const int lxA = tid % G::COLS, rA = tid / G::COLS;
    const bool activeA = rA < NG;
    const int gxA = clampi(x0 - MR + lxA, 0, w - 1);
    const unsigned uw = (unsigned)w;

    if (tid == 0) {
        for (int s = 0; s < NBUF; s++) {
            fbs_mbar_init(bar_full(s), 1);
            fbs_mbar_init(bar_empty(s), 1);
        }
    }
    __syncthreads();
    if (tid >= NT) {  // producer warp: one lane issues the five tensor copies of half hh once its set is free
        if (tid == NT) {
            // the flow at the centre of a half (origin of its R1 box) is requested one half ahead, so that the load is
            // not on the path between "buffer free" and the tensor copies
            auto centre = [&](int hh) {
                const int cy = clampi(y0 - MR + hh * G::TY + G::TY / 2, 0, h - 1), cx = clampi(x0 + TX / 2, 0, w - 1);
                return flow_in ? __ldg(flow_in + (unsigned)cy * uw + (unsigned)cx) : make_float2(0.f, 0.f);
            };
            float2 fc = centre(0);
            for (int hh = 0; hh <= ntiles; hh++) {
                const int s = hh % NBUF, use = hh / NBUF;  // the use-th fill of set s
                if (use > 0 && !fbm_mbar_wait(bar_empty(s), (uint32_t)((use - 1) & 1))) return;
                const uint32_t bar = bar_full(s);
                char* ob = set_base(s);
                float4* boxq = reinterpret_cast<float4*>(ob);
                float* boxe = reinterpret_cast<float*>(ob + B::Q1);
                float4* t0q = reinterpret_cast<float4*>(ob + B::Q1 + B::E1);
                float* t0e = reinterpret_cast<float*>(ob + B::Q1 + B::E1 + B::Q0);
                float2* t0f = reinterpret_cast<float2*>(ob + B::Q1 + B::E1 + B::Q0 + B::E0);
                const int gy_base = y0 - MR + hh * G::TY;
                const int ox = __float2int_rd(fminf(fmaxf(fc.x, -4096.f), 4096.f));
                const int oy = __float2int_rd(fminf(fmaxf(fc.y, -4096.f), 4096.f));
                if (hh < ntiles) fc = centre(hh + 1);
                const int bx0 = x0 - MR - GM + ox, by0 = gy_base - GM + oy;
                const int ty0 = min(gy_base, h - 1);  // keeps row h - 1 inside the tile of a half below the image
                ctl[8 + 4 * s] = bx0; ctl[9 + 4 * s] = by0; ctl[10 + 4 * s] = ty0;
                fbs_fence_proxy_async();  // the buffers were read through the generic proxy until the last barrier
                fbs_mbar_expect_tx(bar, flow_in ? B::TX_BYTES_FLOW : B::TX_BYTES_NOFLOW);
                fbm_tensor_g2s(fbh_smem_u32(t0q), &maps.r0q, 4 * tx0, ty0, bar);
                fbm_tensor_g2s(fbh_smem_u32(t0e), &maps.r0e, tx0 & ~3, ty0, bar);
                if (flow_in) fbm_tensor_g2s(fbh_smem_u32(t0f), &maps.flow, 2 * (tx0 & ~1), ty0, bar);
                fbm_tensor_g2s(fbh_smem_u32(boxq), &maps.r1q, 4 * bx0, by0, bar);
                fbm_tensor_g2s(fbh_smem_u32(boxe), &maps.r1e, bx0 & ~3, by0, bar);
            }
        }
        return;
    }

    for (int hh = 0; hh <= ntiles; hh++) {
        float* new_half = ring + (hh & 1) * G::HALF;
        const int s = hh % NBUF, use = hh / NBUF;
        if (!fbm_mbar_wait(bar_full(s), (uint32_t)(use & 1))) return;
        if (activeA) {
            const char* ob = set_base(s);
            const float4* boxq = reinterpret_cast<const float4*>(ob);
            const float* boxe = reinterpret_cast<const float*>(ob + B::Q1);
            const float4* t0q = reinterpret_cast<const float4*>(ob + B::Q1 + B::E1);
            const float* t0e = reinterpret_cast<const float*>(ob + B::Q1 + B::E1 + B::Q0);
            const float2* t0f = reinterpret_cast<const float2*>(ob + B::Q1 + B::E1 + B::Q0 + B::E0);
            const int bx0 = ctl[8 + 4 * s], by0 = ctl[9 + 4 * s], ty0 = ctl[10 + 4 * s];
            const int gy_base = y0 - MR + hh * G::TY;
            const int lx0 = gxA - tx0, le0 = gxA - (tx0 & ~3), lf0 = gxA - (tx0 & ~1), ex0 = bx0 & ~3;
#pragma unroll 1
            for (int r = rA; r < G::TY; r += NG) {
                const int gy = clampi(gy_base + r, 0, h - 1);
                const int lr = gy - ty0;
                const float2 f = flow_in ? t0f[lr * B::TF + lf0] : make_float2(0.f, 0.f);
                const float4 q0 = t0q[lr * B::TW + lx0];
                const float a[5] = {q0.x, q0.y, q0.z, q0.w, t0e[lr * B::TE + le0]};
                int x1 = __float2int_rd((float)gxA + f.x), yy1 = __float2int_rd((float)gy + f.y);
                const bool in = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1);
                FbhTaps top, bot;
                if (in) {
                    const int bx = x1 - bx0, by = yy1 - by0;
                    if ((unsigned)bx < (unsigned)(B::BW - 1) && (unsigned)by < (unsigned)(B::BH - 1)) {
                        const float4* q = boxq + by * B::BW + bx;
                        const float* e = boxe + by * B::BE + (x1 - ex0);
                        top.q0 = q[0]; top.q1 = q[1]; bot.q0 = q[B::BW]; bot.q1 = q[B::BW + 1];
                        top.e0 = e[0]; top.e1 = e[1]; bot.e0 = e[B::BE]; bot.e1 = e[B::BE + 1];
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
        fbs_bar_consumers<NT>();  // M of this half complete; every thread is done with this set of operands
        if (tid == 0) fbs_mbar_arrive(bar_empty(s));
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

// ---- host side: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point) ----
typedef CUresult (*fbm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int fbm_map2d(CUtensorMap* m, const void* basep, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    static fbm_encode_fn encode = nullptr;
    if (!encode) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        TF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
        TF_REQUIRE(sym && q == cudaDriverEntryPointSuccess, TF_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = reinterpret_cast<fbm_encode_fn>(sym);
    }
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstride[1] = {inner * 4};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(basep), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TF_OK;
}

template <int MR, int TX, int NT, int GM, int NBUF = 1>
static int fb_launch_tma(const float* R0, const float* R1, const float2* in, float2* dst, int w, int h, double scale,
                         int clip, cudaStream_t st) {
    using G = FbhGeom<MR, TX, true>;
    using B = FbmGeom<MR, TX, GM, NBUF>;
    auto kern = k_fb_iter_tma<MR, TX, NT, GM, NBUF>;
    static int resident = 0;
    if (!resident) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B::SMEM));
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, NT + 32, B::SMEM));
        if (resident < 1) return fail(TF_ERR_CUDA, "k_fb_iter_tma<%d,%d> does not fit an SM", MR, TX);
    }
    const size_t plane = (size_t)w * h;
    FbmMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (int e = fbm_map2d(&maps.r1q, R1, 4ull * w, h, 4 * B::BW, B::BH)) return e;
    if (int e = fbm_map2d(&maps.r1e, R1 + 4 * plane, w, h, B::BE, B::BH)) return e;
    if (int e = fbm_map2d(&maps.r0q, R0, 4ull * w, h, 4 * B::TW, G::TY)) return e;
    if (int e = fbm_map2d(&maps.r0e, R0 + 4 * plane, w, h, B::TE, G::TY)) return e;
    if (in)
        if (int e = fbm_map2d(&maps.flow, in, 2ull * w, h, 2 * B::TF, G::TY)) return e;
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
    kern<<<grid, NT + 32, B::SMEM, st>>>(maps, reinterpret_cast<const float4*>(R1), R1 + 4 * plane, in, dst, w, h, reg,
                                         rows, clip);
    return TF_OK;
}

// variants 12 / 13 / 14: 32-column strips (a 64-column box of quads would exceed the 256-element box limit of a 2-D
// tensor map), box margin 3 / 2 / 1 pixels around the footprint displaced by the centre flow
template <typename RT>
static int fb_iterate_tma(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf, float2* other_buf,
                          bool zero_init, int clip, bool finest, cudaStream_t st, int margin = 3) {
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
            // margin 0 / -1 (variants 15 / 16): margin 1 with TWO sets of operand buffers (the copies of half k + 1 run
            // during the whole of half k; 2 CTAs per SM), 192 or 384 compute threads
            e = margin == 0    ? fb_launch_tma<7, 32, 192, 1, 2>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : margin == -1 ? fb_launch_tma<7, 32, 384, 1, 2>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : margin == 1  ? fb_launch_tma<7, 32, 192, 1>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                : margin == 2 ? fb_launch_tma<7, 32, 192, 2>(R0f, R1f, in, dst, L.w, L.h, scale, c, st)
                              : fb_launch_tma<7, 32, 192, 3>(R0f, R1f, in, dst, L.w, L.h, scale, c, st);
        }
        if (e) return e;
        TF_LAUNCHED();
    }
    return TF_OK;
}
