// Probe: issue rate of FFMA / FADD vs the packed FFMA2 / FADD2 of sm_100a (same flops, half the instructions).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }
            if (MODE == 1) v[i] = __ffma2_rn(v[i], A, B);
            if (MODE == 2) { v[i].x = v[i].x + b; v[i].y = v[i].y + b; }
            if (MODE == 3) v[i] = __fadd2_rn(v[i], B);
            if (MODE == 4) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b);
                             asm volatile("" ::: "memory"); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out) {
    int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * 16;   // scalar fp32 operations
    printf("%-8s %.3f ms  %.1f Gop/s  (%.1f op/clk/SM at 1.965 GHz)\n", name, ms, ops / ms * 1e-6, ops / ms * 1e-6 / 148 / 1.965);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out);
    return 0;
}
