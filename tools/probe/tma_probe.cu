// Probe: which 2-D tensor-copy configurations work on this box (each case in its own process).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu   (built binary is git-ignored)
//   tma_probe <desc: 0 param, 1 global> <inner> <rows> <box_inner> <box_rows> <x> <y>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void load2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ bool run(const CUtensorMap* map, float* out, int x, int y, int n, uint32_t bytes) {
    extern __shared__ __align__(128) float buf[];
    __shared__ __align__(8) unsigned long long mbar;
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar), dst = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        load2d(dst, map, x, y, bar);
    }
    bool ok = false;
    for (unsigned t = 0; t < (1u << 20); t++) {
        uint32_t done;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        if (done) { ok = true; break; }
    }
    if (ok) for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
    return ok;
}
__global__ void k_param(const __grid_constant__ CUtensorMap map, float* out, int x, int y, int n, uint32_t bytes, int* flag) {
    bool ok = run(&map, out, x, y, n, bytes);
    if (threadIdx.x == 0) *flag = ok ? 1 : -1;
}
__global__ void k_global(const CUtensorMap* map, float* out, int x, int y, int n, uint32_t bytes, int* flag) {
    bool ok = run(map, out, x, y, n, bytes);
    if (threadIdx.x == 0) *flag = ok ? 1 : -1;
}
int main(int argc, char** argv) {
    if (argc < 8) return 2;
    int mode = atoi(argv[1]); long inner = atol(argv[2]), rows = atol(argv[3]); int bi = atoi(argv[4]), br = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    encode_fn encode = (encode_fn)sym;
    std::vector<float> h(inner * rows);
    for (long i = 0; i < inner * rows; i++) h[i] = (float)(i % 100003);
    float *d, *out; int* flag;
    CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    int n = bi * br;
    CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&flag, 4)); CK(cudaMemset(flag, 0, 4));
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)inner * 4};
    cuuint32_t box[2] = {(cuuint32_t)bi, (cuuint32_t)br}, es[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("FAIL encode %d\n", (int)r); return 1; }
    size_t smem = (size_t)n * 4;
    if (mode == 0) {
        CK(cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_param<<<1, 128, smem>>>(m, out, x, y, n, (uint32_t)smem, flag);
    } else {
        CUtensorMap* dm; CK(cudaMalloc(&dm, sizeof(m))); CK(cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_global<<<1, 128, smem>>>(dm, out, x, y, n, (uint32_t)smem, flag);
    }
    CK(cudaDeviceSynchronize());
    int f = 0; CK(cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost));
    if (f != 1) { printf("FAIL barrier never completed (flag %d)\n", f); return 1; }
    std::vector<float> o(n); CK(cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int rr = 0; rr < br; rr++) for (int c = 0; c < bi; c++) {
        long gx = x + c, gy = y + rr;
        float want = (gx >= 0 && gx < inner && gy >= 0 && gy < rows) ? h[gy * inner + gx] : 0.f;
        if (o[rr * bi + c] != want) bad++;
    }
    printf("%s mode %d inner %ld rows %ld box %dx%d at (%d,%d): %ld mismatches\n", bad ? "FAIL" : "OK", mode, inner, rows, bi, br, x, y, bad);
    return bad != 0;
}
