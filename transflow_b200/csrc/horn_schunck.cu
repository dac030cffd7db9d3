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
    float* stats;   // [4] frobenius^2, max rowsq, max colsq, scratch
    float* vx;      // [W] power-iteration vector
    float* vy;      // [H]
    float* stats_host;  // pinned
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

// One Jacobi sweep.  Block 32 x 8.  Accumulates the squared change of u per row / column / total.
__global__ void __launch_bounds__(256) k_hs_sweep(const float* __restrict__ Ex, const float* __restrict__ Ey,
                                                  const float* __restrict__ Et, const float2* __restrict__ in,
                                                  float2* __restrict__ out, float alpha2, int H, int W, int clip,
                                                  float* __restrict__ rowsq, float* __restrict__ colsq,
                                                  float* __restrict__ stats) {
    __shared__ float scol[8][32];
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    float d2 = 0.f;
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
        float2 o = make_float2(un, vn);
        if (clip) {
            o.x = fminf(fmaxf(o.x, (float)(-x)), (float)(W - 1 - x));
            o.y = fminf(fmaxf(o.y, (float)(-y)), (float)(H - 1 - y));
        }
        out[at] = o;
    }
    if (!rowsq) return;
    // row partial: reduce over the 32 lanes of this warp (one warp == one row of the block)
    float r = d2;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) r += __shfl_xor_sync(0xffffffffu, r, s);
    if (threadIdx.x == 0 && y < H) {
        atomicAdd(rowsq + y, r);
        atomicAdd(stats, r);
    }
    scol[threadIdx.y][threadIdx.x] = d2;
    __syncthreads();
    if (threadIdx.y == 0 && x < W) {
        float c = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) c += scol[j][threadIdx.x];
        atomicAdd(colsq + x, c);
    }
}

__global__ void __launch_bounds__(256) k_hs_max(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float s[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, v[i]);
    s[threadIdx.x] = m;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] = fmaxf(s[threadIdx.x], s[threadIdx.x + k]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

// Power iteration helpers on D = u_new - u_old (H x W): y = D x ; x = D^T y.
__global__ void __launch_bounds__(256) k_hs_dx(const float2* __restrict__ un, const float2* __restrict__ uo,
                                               const float* __restrict__ vx, float* __restrict__ vy, int H, int W) {
    int y = blockIdx.x;
    __shared__ float s[256];
    float acc = 0.f;
    for (int x = threadIdx.x; x < W; x += 256) {
        size_t at = (size_t)y * W + x;
        acc += (un[at].x - uo[at].x) * vx[x];
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) vy[y] = s[0];
}

__global__ void __launch_bounds__(256) k_hs_dty(const float2* __restrict__ un, const float2* __restrict__ uo,
                                                const float* __restrict__ vy, float* __restrict__ vx, int H, int W) {
    int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    float acc = 0.f;
    for (int y = 0; y < H; y++) {
        size_t at = (size_t)y * W + x;
        acc += (un[at].x - uo[at].x) * vy[y];
    }
    vx[x] = acc;
}

// x <- x / ||x||; writes ||x|| to *norm_out
__global__ void __launch_bounds__(256) k_hs_normalize(float* __restrict__ v, int n, float* __restrict__ norm_out) {
    __shared__ float s[256];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) acc += v[i] * v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) s[threadIdx.x] += s[threadIdx.x + k];
        __syncthreads();
    }
    float nrm = sqrtf(s[0]);
    if (threadIdx.x == 0) *norm_out = nrm;
    float inv = nrm > 0.f ? 1.f / nrm : 0.f;
    for (int i = threadIdx.x; i < n; i += 256) v[i] *= inv;
}

// copy-out with the final clip of FlowSource.post_process (source.py:361-362) fused in
__global__ void __launch_bounds__(256) k_hs_copy_out(const float2* __restrict__ src, float2* __restrict__ dst, int H,
                                                     int W, int clip) {
    int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float2 o = src[(size_t)y * W + x];
    if (clip) {
        o.x = fminf(fmaxf(o.x, (float)(-x)), (float)(W - 1 - x));
        o.y = fminf(fmaxf(o.y, (float)(-y)), (float)(H - 1 - y));
    }
    dst[(size_t)y * W + x] = o;
}

__global__ void __launch_bounds__(256) k_hs_fill(float* v, int n, float val) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) v[i] = val;
}

extern "C" int tf_hs_destroy(tf_horn_schunck* h) {
    if (!h) return TF_OK;
    cudaFree(h->Ex); cudaFree(h->Ey); cudaFree(h->Et); cudaFree(h->uv[0]); cudaFree(h->uv[1]);
    cudaFree(h->rowsq); cudaFree(h->colsq); cudaFree(h->stats); cudaFree(h->vx); cudaFree(h->vy);
    if (h->stats_host) cudaFreeHost(h->stats_host);
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
              cudaMalloc(&h->colsq, width * 4) == cudaSuccess && cudaMalloc(&h->stats, 16) == cudaSuccess &&
              cudaMalloc(&h->vx, width * 4) == cudaSuccess && cudaMalloc(&h->vy, height * 4) == cudaSuccess &&
              cudaMallocHost(&h->stats_host, 16) == cudaSuccess;
    if (!ok) {
        tf_hs_destroy(h);
        return fail(TF_ERR_CUDA, "tf_hs_create: allocation failed for %dx%d", height, width);
    }
    *out = h;
    return TF_OK;
}

// sigma_max(D) >= delta ?  Power iteration on D^T D gives a monotone lower bound.
static int hs_sigma_reaches(tf_horn_schunck* h, const float2* un, const float2* uo, double delta, bool* reaches,
                            cudaStream_t st) {
    k_hs_fill<<<ceil_div(h->W, 256), 256, 0, st>>>(h->vx, h->W, 1.0f / sqrtf((float)h->W));
    TF_LAUNCHED();
    *reaches = false;
    for (int it = 0; it < 64; it++) {
        k_hs_dx<<<h->H, 256, 0, st>>>(un, uo, h->vx, h->vy, h->H, h->W);
        TF_LAUNCHED();
        k_hs_normalize<<<1, 256, 0, st>>>(h->vy, h->H, h->stats + 3);  // ||D x|| with ||x|| = 1: a lower bound
        TF_LAUNCHED();
        k_hs_dty<<<ceil_div(h->W, 256), 256, 0, st>>>(un, uo, h->vy, h->vx, h->H, h->W);
        TF_LAUNCHED();
        k_hs_normalize<<<1, 256, 0, st>>>(h->vx, h->W, h->stats + 2);
        TF_LAUNCHED();
        if ((it & 7) == 7) {
            TF_CUDA(cudaMemcpyAsync(h->stats_host, h->stats, 16, cudaMemcpyDeviceToHost, st));
            TF_CUDA(cudaStreamSynchronize(st));
            if ((double)h->stats_host[3] >= delta) {
                *reaches = true;
                return TF_OK;
            }
        }
    }
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
    float2* out = reinterpret_cast<float2*>(flow);
    // ping-pong so that the last sweep that runs writes `out`... the number of sweeps is only known
    // at run time (early exit), so sweeps alternate between two scratch planes and the result is
    // copied (or clipped) into `out` at the end.
    k_hs_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(prev_flow), h->uv[0],
                                                          (float)decay, n);
    TF_LAUNCHED();
    int cur = 0, done = 0;
    bool track = delta >= 0;
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8)), block(32, 8);
    for (int it = 0; it < max_iters; it++) {
        if (track) {
            TF_CUDA(cudaMemsetAsync(h->rowsq, 0, H * 4, st));
            TF_CUDA(cudaMemsetAsync(h->colsq, 0, W * 4, st));
            TF_CUDA(cudaMemsetAsync(h->stats, 0, 16, st));
        }
        {
            ScopedKernelTimer timer(TFK_HS_SWEEP, st);
            k_hs_sweep<<<grid, block, 0, st>>>(h->Ex, h->Ey, h->Et, h->uv[cur], h->uv[cur ^ 1], (float)(alpha * alpha), H,
                                               W, 0, track ? h->rowsq : nullptr, h->colsq, h->stats);
        }
        TF_LAUNCHED();
        cur ^= 1;
        done++;
        if (track && it + 1 < max_iters) {
            k_hs_max<<<1, 256, 0, st>>>(h->rowsq, H, h->stats + 1);
            TF_LAUNCHED();
            k_hs_max<<<1, 256, 0, st>>>(h->colsq, W, h->stats + 2);
            TF_LAUNCHED();
            TF_CUDA(cudaMemcpyAsync(h->stats_host, h->stats, 16, cudaMemcpyDeviceToHost, st));
            TF_CUDA(cudaStreamSynchronize(st));
            double upper = sqrt((double)h->stats_host[0]);
            double lower = sqrt((double)fmaxf(h->stats_host[1], h->stats_host[2]));
            if (upper < delta) break;  // sigma_max <= ||D||_F < delta: converged
            if (lower < delta) {       // undecided: ask the power iteration
                bool reaches = false;
                if (int e = hs_sigma_reaches(h, h->uv[cur], h->uv[cur ^ 1], delta, &reaches, st)) return e;
                if (!reaches) break;
            }
        }
    }
    k_hs_copy_out<<<dim3(ceil_div(W, 256), H), 256, 0, st>>>(h->uv[cur], out, H, W, clip);
    TF_LAUNCHED();
    if (sweeps_done_host) *sweeps_done_host = done;
    return TF_OK;
}
