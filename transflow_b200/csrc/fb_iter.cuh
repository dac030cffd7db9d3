// Fused Farneback iteration: update-matrices + (2m+1)^2 box sums + 2x2 solve in ONE kernel, so
// the 5-channel matrix M is never written to HBM (SURVEY.md 8d byte model: per iteration
// read R0, R1, flow; write flow).
//
// Streaming column strips: a CTA owns TX = 128 output columns (+ m halo columns each side,
// one thread per column) and marches down a chunk of rows.  Per matrix row each thread
//   * computes M(x, row) for its column (R0 coalesced, R1 4-tap bilinear gather),
//   * keeps the vertical window sum in DOUBLE registers (vs += new - old; `old` comes from a
//     (2m+1)-row ring in shared memory that only this thread touches -> no bank conflicts),
//   * publishes the sum for output row (row - m) to shared memory.
// Every RB = 8 rows the CTA synchronises and 64 threads run the horizontal pass: each slides
// a double accumulator over 16 consecutive outputs for all 5 channels and solves the 2x2
// system; results are staged in shared memory and stored with coalesced 8-byte writes.
// Borders replicate (cv2: vsum / srow clamping), so M is evaluated at clamped coordinates.
#pragma once
#include "fb_math.cuh"

#define FBI_TX 128
#define FBI_NT 160
#define FBI_RB 8
#define FBI_NOUT 16
#define FBI_NSEG (FBI_TX / FBI_NOUT)
// shared-memory column index with one pad element per 16 columns (bank-conflict-free phase B)
#define FBI_SKEW(c) ((c) + ((c) >> 4))
#define FBI_VS_PITCH 200  // >= FBI_SKEW(FBI_NT - 1) + 1 and == 8 (mod 32)

template <typename RT, typename VST>
__global__ void __launch_bounds__(FBI_NT) k_fb_iter(const RT* __restrict__ R0, const RT* __restrict__ R1,
                                                    const float2* __restrict__ flow_in, float2* __restrict__ flow_out,
                                                    int w, int h, int m, double scale, int rows_per_cta, int clip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int win = 2 * m + 1;
    float* ring = reinterpret_cast<float*>(smem_raw);                      // [win][5][FBI_NT]
    VST* svs = reinterpret_cast<VST*>(ring + (size_t)win * 5 * FBI_NT);    // [RB][5][FBI_VS_PITCH]
    float2* sout = reinterpret_cast<float2*>(svs + FBI_RB * 5 * FBI_VS_PITCH);  // [RB][FBI_TX]

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FBI_TX;
    const int y0 = blockIdx.y * rows_per_cta;
    const int y1 = min(h, y0 + rows_per_cta);
    const int cols = FBI_TX + 2 * m;
    const bool col_active = tid < cols;
    const int gx = clampi(x0 - m + tid, 0, w - 1);
    const size_t plane = (size_t)w * h;

    double vs[5] = {0., 0., 0., 0., 0.};
    // matrix rows r in [y0 - m, y1 + m); output row yc = r - m is complete once row r is added
    for (int rbase = y0 - m; rbase < y1 + m; rbase += FBI_RB) {
        if (col_active) {
#pragma unroll 2
            for (int i = 0; i < FBI_RB; i++) {
                int r = rbase + i;
                if (r >= y1 + m) break;
                int gy = clampi(r, 0, h - 1);
                float2 f = flow_in ? __ldg(flow_in + (size_t)gy * w + gx) : make_float2(0.f, 0.f);
                float mm[5];
                fb_update_matrix<RT>(R0, R1, plane, w, h, gx, gy, f, mm);
                int slot = (r - (y0 - m)) % win;
                float* rp = ring + (size_t)slot * 5 * FBI_NT + tid;
                bool have_old = r - win >= y0 - m;
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    float old = have_old ? rp[c * FBI_NT] : 0.f;
                    vs[c] += (double)mm[c] - (double)old;
                    rp[c * FBI_NT] = mm[c];
                }
                int yc = r - m;
                if (yc >= y0) {
                    VST* vp = svs + (size_t)((yc - y0) % FBI_RB) * 5 * FBI_VS_PITCH + FBI_SKEW(tid);
#pragma unroll
                    for (int c = 0; c < 5; c++) vp[c * FBI_VS_PITCH] = (VST)vs[c];
                }
            }
        }
        __syncthreads();
        // output rows completed by this step: yc in [max(y0, rbase - m), min(y1, rbase + RB - m))
        int oy_lo = max(y0, rbase - m), oy_hi = min(y1, rbase + FBI_RB - m);
        if (tid < FBI_RB * FBI_NSEG) {
            int rb = tid / FBI_NSEG, seg = tid % FBI_NSEG;
            int yc = oy_lo + rb;
            if (yc < oy_hi) {
                const VST* vp = svs + (size_t)((yc - y0) % FBI_RB) * 5 * FBI_VS_PITCH;
                int c0 = seg * FBI_NOUT;  // first output column (local); window centre index = c0 + m
                double s[5] = {0., 0., 0., 0., 0.};
                for (int k = 0; k < win; k++) {
                    int ci = FBI_SKEW(c0 + k);
#pragma unroll
                    for (int c = 0; c < 5; c++) s[c] += (double)vp[c * FBI_VS_PITCH + ci];
                }
                for (int o = 0; o < FBI_NOUT; o++) {
                    if (o > 0) {
                        int ca = FBI_SKEW(c0 + o + win - 1), cb = FBI_SKEW(c0 + o - 1);
#pragma unroll
                        for (int c = 0; c < 5; c++)
                            s[c] += (double)vp[c * FBI_VS_PITCH + ca] - (double)vp[c * FBI_VS_PITCH + cb];
                    }
                    float2 fl = fb_solve(s, scale);
                    if (clip) {
                        int x = x0 + c0 + o;
                        fl.x = fminf(fmaxf(fl.x, (float)(-x)), (float)(w - 1 - x));
                        fl.y = fminf(fmaxf(fl.y, (float)(-yc)), (float)(h - 1 - yc));
                    }
                    sout[rb * FBI_TX + c0 + o] = fl;
                }
            }
        }
        __syncthreads();
        int nrows = oy_hi - oy_lo;
        for (int i = tid; i < nrows * FBI_TX; i += FBI_NT) {
            int rb = i / FBI_TX, lx = i - rb * FBI_TX;
            int x = x0 + lx;
            if (x < w) flow_out[(size_t)(oy_lo + rb) * w + x] = sout[rb * FBI_TX + lx];
        }
        // the next step's phase A overwrites svs rows that phase B of this step has finished with
        // (guarded by the second __syncthreads above); sout is rewritten only after the next sync.
    }
}

template <typename VST>
static size_t fb_iter_smem_bytes(int m) {
    return (size_t)(2 * m + 1) * 5 * FBI_NT * sizeof(float) + (size_t)FBI_RB * 5 * FBI_VS_PITCH * sizeof(VST) +
           (size_t)FBI_RB * FBI_TX * sizeof(float2);
}

// One level: `iterations` launches ping-ponging between the level's two flow buffers so that the
// last iteration lands in `final_buf`.  init_in_final tells where the initial flow currently is.
template <typename RT>
static int fb_iterate_fused(tf_farneback* h, FbLevel& L, const RT* R0, const RT* R1, float2* final_buf,
                            float2* other_buf, bool zero_init, int clip, int precise, bool finest, cudaStream_t st) {
    int m = h->winsize / 2;
    double scale = 1.0 / ((double)h->winsize * h->winsize);
    int strips = ceil_div(L.w, FBI_TX);
    int chunks = std::max(1, (2 * sm_count() + strips / 2) / strips);
    int rows = std::max(32, ceil_div(L.h, chunks));
    chunks = ceil_div(L.h, rows);
    dim3 grid(strips, chunks);
    size_t smem = precise ? fb_iter_smem_bytes<double>(m) : fb_iter_smem_bytes<float>(m);
    auto kern = precise ? k_fb_iter<RT, double> : k_fb_iter<RT, float>;
    static bool attr_set[2][2] = {{false, false}, {false, false}};
    bool& done = attr_set[sizeof(RT) == 2][precise ? 1 : 0];
    if (!done) {
        TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done = true;
    }
    TF_REQUIRE(smem <= 200 * 1024, TF_ERR_INVALID_ARG, "fb_winsize %d needs %zu bytes of shared memory", h->winsize, smem);
    int T = h->iterations;
    for (int it = 0; it < T; it++) {
        // iteration `it` writes `final_buf` when (T - 1 - it) is even
        float2* dst = ((T - 1 - it) & 1) ? other_buf : final_buf;
        float2* src = ((T - 1 - it) & 1) ? final_buf : other_buf;
        const float2* in = (it == 0 && zero_init) ? nullptr : src;
        {
            ScopedKernelTimer timer(finest ? TFK_FB_ITER_FINEST : -1, st);
            kern<<<grid, FBI_NT, smem, st>>>(R0, R1, in, dst, L.w, L.h, m, scale, rows, clip && it + 1 == T);
        }
        TF_LAUNCHED();
    }
    return TF_OK;
}
