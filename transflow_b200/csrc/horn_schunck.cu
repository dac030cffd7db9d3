// Horn-Schunck optical flow on device (transflow/flow/methods/horn_schunck.py:9-45).
//
//   derivatives : 5x5 [1 4 6 4 1]/16 Gaussian (BORDER_REFLECT_101) of both frames fused with the
//                 2x2 forward-difference stencils Ex, Ey, Et (scipy 'reflect' = symmetric border)
//   sweep       : Jacobi update with the 3x3 [1 2 1; 2 0 2; 1 2 1]/12 average, fused with the
//                 reductions the early exit needs
//   early exit  : the reference stops when numpy.linalg.norm(u - u_prev, 2) < delta, i.e. on the
//                 LARGEST SINGULAR VALUE of the H x W change matrix (quirk Q4).  We bound it:
//                 max(row norm, column norm) <= sigma_max <= Frobenius norm, and only when delta
//                 falls between the bounds run power iterations on D^T D.
#include "common.cuh"

using namespace tf;

struct tf_horn_schunck {
    int H, W;
    float *Ex, *Ey, *Et;
    float2* uv[2];
    float* rowsq;   // [H]  sum over x of d^2
    float* colsq;   // [W]
    float* rowabs;  // [H]  sum over x of |d|
    float* colabs;  // [W]
    float* stats;   // [4] frobenius^2
    float* vx;      // [W] power-iteration vector
    float* vy;      // [H]
    struct HsState* state;  // device
};

#define HS_TX 32
#define HS_TY 16

__global__ void __launch_bounds__(HS_TX* HS_TY) k_hs_derivatives(const uint8_t* __restrict__ left,
                                                                 const uint8_t* __restrict__ right,
                                                                 float* __restrict__ Ex, float* __restrict__ Ey,
                                                                 float* __restrict__ Et, int H, int W) {
    // horizontal pass of both frames for absolute rows y0-2 .. y0+TY+2, columns x0 .. x0+TX
    __shared__ float sh[2][HS_TY + 5][HS_TX + 1];
    __shared__ float sb[2][HS_TY + 1][HS_TX + 1];
    const float k0 = 0.0625f, k1 = 0.25f, k2 = 0.375f;
    int x0 = blockIdx.x * HS_TX, y0 = blockIdx.y * HS_TY;
    int tid = threadIdx.y * HS_TX + threadIdx.x;
    for (int i = tid; i < (HS_TY + 5) * (HS_TX + 1); i += HS_TX * HS_TY) {
        int ly = i / (HS_TX + 1), lx = i - ly * (HS_TX + 1);
        int r = y0 - 2 + ly;
        int xx = min(x0 + lx, W - 1);
        float a = 0.f, b = 0.f;
        if (r >= 0 && r < H) {
            const uint8_t* la = left + (size_t)r * W;
            const uint8_t* lb = right + (size_t)r * W;
            int c0 = reflect101(xx - 2, W), c1 = reflect101(xx - 1, W), c3 = reflect101(xx + 1, W),
                c4 = reflect101(xx + 2, W);
            a = k0 * (float)la[c0] + k1 * (float)la[c1] + k2 * (float)la[xx] + k1 * (float)la[c3] + k0 * (float)la[c4];
            b = k0 * (float)lb[c0] + k1 * (float)lb[c1] + k2 * (float)lb[xx] + k1 * (float)lb[c3] + k0 * (float)lb[c4];
        }
        sh[0][ly][lx] = a;
        sh[1][ly][lx] = b;
    }
    __syncthreads();
    for (int i = tid; i < (HS_TY + 1) * (HS_TX + 1); i += HS_TX * HS_TY) {
        int ly = i / (HS_TX + 1), lx = i - ly * (HS_TX + 1);
        int yy = min(y0 + ly, H - 1);
        int base = y0 - 2;
        int r0 = reflect101(yy - 2, H) - base, r1 = reflect101(yy - 1, H) - base, r2 = yy - base,
            r3 = reflect101(yy + 1, H) - base, r4 = reflect101(yy + 2, H) - base;
#pragma unroll
        for (int im = 0; im < 2; im++)
            sb[im][ly][lx] = k0 * sh[im][r0][lx] + k1 * sh[im][r1][lx] + k2 * sh[im][r2][lx] + k1 * sh[im][r3][lx] +
                             k0 * sh[im][r4][lx];
    }
    __syncthreads();
    int lx = threadIdx.x, ly = threadIdx.y;
    int x = x0 + lx, y = y0 + ly;
    if (x >= W || y >= H) return;
    float a00 = sb[0][ly][lx], a01 = sb[0][ly][lx + 1], a10 = sb[0][ly + 1][lx], a11 = sb[0][ly + 1][lx + 1];
    float b00 = sb[1][ly][lx], b01 = sb[1][ly][lx + 1], b10 = sb[1][ly + 1][lx], b11 = sb[1][ly + 1][lx + 1];
    size_t at = (size_t)y * W + x;
    Ex[at] = 0.25f * (a11 - a10 + a01 - a00) + 0.25f * (b11 - b10 + b01 - b00);
    Ey[at] = 0.25f * (a11 + a10 - a01 - a00) + 0.25f * (b11 + b10 - b01 - b00);
    Et[at] = 0.25f * (b11 + b10 + b01 + b00) - 0.25f * (a11 + a10 + a01 + a00);
}

__global__ void __launch_bounds__(256) k_hs_init(const float2* __restrict__ prev, float2* __restrict__ uv, float decay,
                                                 size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 f = make_float2(0.f, 0.f);
    if (prev) {
        f = prev[i];
        f.x = __fmul_rn(decay, f.x);
        f.y = __fmul_rn(decay, f.y);
    }
    uv[i] = f;
}

// Device-side state of one tf_hs_run: the whole call is asynchronous, no host round trip per sweep.
struct HsState {
    int done;       // the early-exit test fired: later sweeps are no-ops
    int sweeps;     // sweeps executed
    int cur;        // index of the buffer holding the current (u, v)
    int ambiguous;  // the cheap bounds could not decide; the power iteration must
};

// One Jacobi sweep.  Block 32 x 8.  Accumulates, for D = u_new - u_old: row / column sums of D^2 and
// of |D| and the total of D^2 (bounds on the spectral norm, see k_hs_decide).
__global__ void __launch_bounds__(256) k_hs_sweep(const float* __restrict__ Ex, const float* __restrict__ Ey,
                                                  const float* __restrict__ Et, float2* __restrict__ uv0,
                                                  float2* __restrict__ uv1, float alpha2, int H, int W,
                                                  const HsState* __restrict__ state, int track,
                                                  float* __restrict__ rowsq, float* __restrict__ colsq,
                                                  float* __restrict__ rowabs, float* __restrict__ colabs,
                                                  float* __restrict__ stats) {
    __shared__ float scol[2][8][32];
    if (state->done) return;
    const float2* in = state->cur ? uv1 : uv0;
    float2* out = state->cur ? uv0 : uv1;
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    float d2 = 0.f, da = 0.f;
    if (x < W && y < H) {
        int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
        const float2* r0 = in + (size_t)ym * W;
        const float2* r1 = in + (size_t)y * W;
        const float2* r2 = in + (size_t)yp * W;
        float2 c00 = r0[xm], c01 = r0[x], c02 = r0[xp], c10 = r1[xm], c11 = r1[x], c12 = r1[xp], c20 = r2[xm],
               c21 = r2[x], c22 = r2[xp];
        float ua = ((c00.x + c02.x + c20.x + c22.x) + 2.f * (c01.x + c21.x + c10.x + c12.x)) * (1.f / 12.f);
        float va = ((c00.y + c02.y + c20.y + c22.y) + 2.f * (c01.y + c21.y + c10.y + c12.y)) * (1.f / 12.f);
        size_t at = (size_t)y * W + x;
        float ex = Ex[at], ey = Ey[at], et = Et[at];
        float c = (ex * ua + ey * va + et) / (alpha2 + ex * ex + ey * ey);
        float un = ua - ex * c, vn = va - ey * c;
        float d = un - c11.x;
        d2 = d * d;
        da = fabsf(d);
        out[at] = make_float2(un, vn);
    }
    if (!track) return;
    // row partials: one warp == one row of the block
    float r = d2, ra = da;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, s);
        ra += __shfl_xor_sync(0xffffffffu, ra, s);
    }
    if (threadIdx.x == 0 && y < H) {
        atomicAdd(rowsq + y, r);
        atomicAdd(rowabs + y, ra);
    }
    scol[0][threadIdx.y][threadIdx.x] = d2;
    scol[1][threadIdx.y][threadIdx.x] = da;
    __syncthreads();
    if (threadIdx.y < 2 && x < W) {
        float c = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) c += scol[threadIdx.y][j][threadIdx.x];
        atomicAdd((threadIdx.y ? colabs : colsq) + x, c);
    }
}

__device__ __forceinline__ float hs_block_max(float v, float* sh) {
    int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int k = blockDim.x >> 1; k > 0; k >>= 1) {
        if (t < k) sh[t] = fmaxf(sh[t], sh[t + k]);
        __syncthreads();
    }
    float m = sh[0];
    __syncthreads();
    return m;
}

__device__ __forceinline__ float hs_block_sum(float v, float* sh) {
    int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int k = blockDim.x >> 1; k > 0; k >>= 1) {
        if (t < k) sh[t] += sh[t + k];
        __syncthreads();
    }
    float m = sh[0];
    __syncthreads();
    return m;
}

// After a sweep: book-keeping plus the early-exit test of horn_schunck.py:43 on the device.
//   max(row 2-norm, column 2-norm) <= sigma_max(D) <= min(||D||_F, sqrt(||D||_1 * ||D||_inf))
// delta above the upper bound -> converged; at or below the lower bound -> keep sweeping; in between
// the power iteration (k_hs_power) decides.  Also clears the accumulators for the next sweep.
__global__ void __launch_bounds__(256) k_hs_decide(HsState* state, float* rowsq, float* colsq, float* rowabs,
                                                   float* colabs, float* stats, int H, int W, float delta, int check) {
    __shared__ float sh[256];
    if (state->done) return;
    float mr = 0.f, mc = 0.f, ar = 0.f, ac = 0.f, frob2 = 0.f;
    if (check) {
        for (int i = threadIdx.x; i < H; i += 256) {
            mr = fmaxf(mr, rowsq[i]);
            ar = fmaxf(ar, rowabs[i]);
            frob2 += rowsq[i];   // ||D||_F^2 = sum of the row sums (no single-address atomic in the sweep)
        }
        for (int i = threadIdx.x; i < W; i += 256) {
            mc = fmaxf(mc, colsq[i]);
            ac = fmaxf(ac, colabs[i]);
        }
        mr = hs_block_max(mr, sh);
        mc = hs_block_max(mc, sh);
        ar = hs_block_max(ar, sh);   // ||D||_inf (max absolute row sum)
        ac = hs_block_max(ac, sh);   // ||D||_1   (max absolute column sum)
        frob2 = hs_block_sum(frob2, sh);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += 256) rowsq[i] = rowabs[i] = 0.f;
    for (int i = threadIdx.x; i < W; i += 256) colsq[i] = colabs[i] = 0.f;
    if (threadIdx.x == 0) {
        stats[0] = 0.f;
        state->sweeps += 1;
        state->cur ^= 1;
        if (check) {
            float upper = fminf(sqrtf(frob2), sqrtf(ar * ac));
            float lower = sqrtf(fmaxf(mr, mc));
            if (upper < delta) state->done = 1;
            else if (lower < delta) state->ambiguous = 1;
        }
    }
}

// Rare path: power iteration on D^T D, one block (the decision is needed on the device, in stream
// order, and the common case must cost one empty launch).  ||D x|| with ||x|| = 1 is a lower bound of
// sigma_max that grows monotonically; reaching delta means "keep sweeping".
__global__ void __launch_bounds__(1024) k_hs_power(HsState* state, const float2* __restrict__ uv0,
                                                   const float2* __restrict__ uv1, float* vx, float* vy, int H, int W,
                                                   float delta, int max_steps) {
    __shared__ float sh[1024];
    if (state->done || !state->ambiguous) return;
    const float2* un = state->cur ? uv1 : uv0;  // k_hs_decide already flipped cur: un = newest
    const float2* uo = state->cur ? uv0 : uv1;
    int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < W; i += 1024) vx[i] = rsqrtf((float)W);
    __syncthreads();
    bool reaches = false;
    for (int step = 0; step < max_steps && !reaches; step++) {
        // vy = D vx (one warp per row)
        for (int y = warp; y < H; y += 32) {
            float acc = 0.f;
            for (int x = lane; x < W; x += 32) {
                size_t at = (size_t)y * W + x;
                acc += (un[at].x - uo[at].x) * vx[x];
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
            if (lane == 0) vy[y] = acc;
        }
        __syncthreads();
        float part = 0.f;
        for (int i = t; i < H; i += 1024) part += vy[i] * vy[i];
        float ny = sqrtf(hs_block_sum(part, sh));
        if (ny >= delta) {
            reaches = true;
            break;
        }
        if (ny == 0.f) break;
        // vx = D^T vy / ||.||  (one thread per column, rows strided: coalesced across threads)
        for (int x = t; x < W; x += 1024) {
            float acc = 0.f;
            for (int y = 0; y < H; y++) {
                size_t at = (size_t)y * W + x;
                acc += (un[at].x - uo[at].x) * vy[y];
            }
            vx[x] = acc;
        }
        __syncthreads();
        part = 0.f;
        for (int i = t; i < W; i += 1024) part += vx[i] * vx[i];
        float nx = sqrtf(hs_block_sum(part, sh));
        if (nx == 0.f) break;
        float inv = 1.f / nx;
        for (int i = t; i < W; i += 1024) vx[i] *= inv;
        __syncthreads();
    }
    if (t == 0) {
        state->ambiguous = 0;
        if (!reaches) state->done = 1;
    }
}

__global__ void k_hs_begin(HsState* state, float* rowsq, float* colsq, float* rowabs, float* colabs, float* stats,
                           int H, int W) {
    for (int i = threadIdx.x; i < H; i += blockDim.x) rowsq[i] = rowabs[i] = 0.f;
    for (int i = threadIdx.x; i < W; i += blockDim.x) colsq[i] = colabs[i] = 0.f;
    if (threadIdx.x == 0) {
        stats[0] = 0.f;
        state->done = state->sweeps = state->cur = state->ambiguous = 0;
    }
}

// copy-out of the buffer the state points at, with the final clip of FlowSource.post_process
// (source.py:361-362) fused in
__global__ void __launch_bounds__(256) k_hs_copy_out(const float2* __restrict__ uv0, const float2* __restrict__ uv1,
                                                     const HsState* __restrict__ state, float2* __restrict__ dst,
                                                     int H, int W, int clip) {
    int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float2* src = state->cur ? uv1 : uv0;
    float2 o = src[(size_t)y * W + x];
    if (clip) {
        o.x = fminf(fmaxf(o.x, (float)(-x)), (float)(W - 1 - x));
        o.y = fminf(fmaxf(o.y, (float)(-y)), (float)(H - 1 - y));
    }
    dst[(size_t)y * W + x] = o;
}

extern "C" int tf_hs_destroy(tf_horn_schunck* h) {
    if (!h) return TF_OK;
    cudaFree(h->Ex); cudaFree(h->Ey); cudaFree(h->Et); cudaFree(h->uv[0]); cudaFree(h->uv[1]);
    cudaFree(h->rowsq); cudaFree(h->colsq); cudaFree(h->rowabs); cudaFree(h->colabs); cudaFree(h->stats);
    cudaFree(h->vx); cudaFree(h->vy); cudaFree(h->state);
    delete h;
    return TF_OK;
}

extern "C" int tf_hs_create(tf_horn_schunck** out, int height, int width) {
    TF_REQUIRE(out, TF_ERR_INVALID_ARG, "tf_hs_create: null out");
    TF_REQUIRE(height >= 3 && width >= 3 && (size_t)height * width < (1u << 30), TF_ERR_SHAPE,
               "tf_hs_create: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    tf_horn_schunck* h = new (std::nothrow) tf_horn_schunck();
    TF_REQUIRE(h, TF_ERR_CUDA, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->H = height;
    h->W = width;
    size_t n = (size_t)height * width;
    bool ok = cudaMalloc(&h->Ex, n * 4) == cudaSuccess && cudaMalloc(&h->Ey, n * 4) == cudaSuccess &&
              cudaMalloc(&h->Et, n * 4) == cudaSuccess && cudaMalloc(&h->uv[0], n * 8) == cudaSuccess &&
              cudaMalloc(&h->uv[1], n * 8) == cudaSuccess && cudaMalloc(&h->rowsq, height * 4) == cudaSuccess &&
              cudaMalloc(&h->colsq, width * 4) == cudaSuccess && cudaMalloc(&h->rowabs, height * 4) == cudaSuccess &&
              cudaMalloc(&h->colabs, width * 4) == cudaSuccess && cudaMalloc(&h->stats, 16) == cudaSuccess &&
              cudaMalloc(&h->vx, width * 4) == cudaSuccess && cudaMalloc(&h->vy, height * 4) == cudaSuccess &&
              cudaMalloc(&h->state, sizeof(HsState)) == cudaSuccess;
    if (!ok) {
        tf_hs_destroy(h);
        return fail(TF_ERR_CUDA, "tf_hs_create: allocation failed for %dx%d", height, width);
    }
    *out = h;
    return TF_OK;
}

extern "C" int tf_hs_run(tf_horn_schunck* h, const uint8_t* left, const uint8_t* right, const float* prev_flow,
                         double alpha, int max_iters, double decay, double delta, float* flow, int clip,
                         int* sweeps_done_host, void* stream) {
    TF_REQUIRE(h && left && right && flow, TF_ERR_INVALID_ARG, "tf_hs_run: null argument");
    TF_REQUIRE(max_iters >= 0, TF_ERR_INVALID_ARG, "hs_iterations must be >= 0");
    TF_REQUIRE(((uintptr_t)flow & 7) == 0 && ((uintptr_t)prev_flow & 7) == 0, TF_ERR_INVALID_ARG,
               "tf_hs_run: flow buffers must be 8-byte aligned");
    cudaStream_t st = as_stream(stream);
    int H = h->H, W = h->W;
    size_t n = (size_t)H * W;
    k_hs_derivatives<<<dim3(ceil_div(W, HS_TX), ceil_div(H, HS_TY)), dim3(HS_TX, HS_TY), 0, st>>>(left, right, h->Ex,
                                                                                                 h->Ey, h->Et, H, W);
    TF_LAUNCHED();
    k_hs_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(prev_flow), h->uv[0],
                                                          (float)decay, n);
    TF_LAUNCHED();
    k_hs_begin<<<1, 256, 0, st>>>(h->state, h->rowsq, h->colsq, h->rowabs, h->colabs, h->stats, H, W);
    TF_LAUNCHED();
    // The sweep count is only known on the device (early exit): every launch below is conditional on
    // the device-side state, nothing here waits for the GPU.
    bool track = delta >= 0;
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8)), block(32, 8);
    for (int it = 0; it < max_iters; it++) {
        bool check = track && it + 1 < max_iters;  // the reference's test after the last sweep changes nothing
        {
            ScopedKernelTimer timer(TFK_HS_SWEEP, st);
            k_hs_sweep<<<grid, block, 0, st>>>(h->Ex, h->Ey, h->Et, h->uv[0], h->uv[1], (float)(alpha * alpha), H, W,
                                               h->state, check, h->rowsq, h->colsq, h->rowabs, h->colabs, h->stats);
        }
        TF_LAUNCHED();
        k_hs_decide<<<1, 256, 0, st>>>(h->state, h->rowsq, h->colsq, h->rowabs, h->colabs, h->stats, H, W, (float)delta,
                                       check);
        TF_LAUNCHED();
        if (check) {
            k_hs_power<<<1, 1024, 0, st>>>(h->state, h->uv[0], h->uv[1], h->vx, h->vy, H, W, (float)delta, 48);
            TF_LAUNCHED();
        }
    }
    k_hs_copy_out<<<dim3(ceil_div(W, 256), H), 256, 0, st>>>(h->uv[0], h->uv[1], h->state, reinterpret_cast<float2*>(flow),
                                                             H, W, clip);
    TF_LAUNCHED();
    if (sweeps_done_host) {  // only tests ask: this is the one place that waits for the device
        HsState hs;
        TF_CUDA(cudaMemcpyAsync(&hs, h->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
        TF_CUDA(cudaStreamSynchronize(st));
        *sweeps_done_host = hs.sweeps;
    }
    return TF_OK;
}
