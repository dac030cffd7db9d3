// Shared helpers for the transflow_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <new>

#include "../../include/transflow_b200.h"

namespace tf {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define TF_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return tf::fail(TF_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                            cudaGetErrorString(_e));                                           \
    } while (0)

#define TF_REQUIRE(cond, code, ...)                                                            \
    do {                                                                                       \
        if (!(cond)) return tf::fail(code, __VA_ARGS__);                                       \
    } while (0)

// Count + check a kernel launch (launch errors only; execution errors surface at the next sync).
#define TF_LAUNCHED()                                                                          \
    do {                                                                                       \
        tf::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return tf::fail(TF_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__,     \
                            cudaGetErrorString(_e));                                           \
    } while (0)

// ---- optional per-kernel device timing (bench.py roofline): CUDA events around tagged launches ----
enum KernelTag {
    TFK_FB_ITER_FINEST = 0,   // fused Farneback iteration kernel at the finest pyramid level
    TFK_FB_UM_FINEST = 1,     // unfused variant: update-matrices
    TFK_FB_BOXV_FINEST = 2,   //                  vertical box sums
    TFK_FB_BOXH_FINEST = 3,   //                  horizontal box sums + solve
    TFK_FB_POLYEXP_FINEST = 4,
    TFK_COMPOSITOR_LAYER = 5, // fused move/reset/remap/composite kernel
    TFK_POST_FORWARD = 6,
    TFK_HS_SWEEP = 7,
    TFK_LK_TRACK_FINEST = 8,
    TFK_COUNT = 9
};
void timer_begin(int tag, cudaStream_t st);
void timer_end(int tag, cudaStream_t st);
struct ScopedKernelTimer {
    int tag;
    cudaStream_t st;
    ScopedKernelTimer(int t, cudaStream_t s) : tag(t), st(s) { timer_begin(tag, st); }
    ~ScopedKernelTimer() { timer_end(tag, st); }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Verifies the current device is sm_100 (B200); cached per process.
int require_sm100();
int sm_count();

template <typename T>
int dev_alloc(T** p, size_t count) {
    TF_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return TF_OK;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    // BORDER_REFLECT_101 for |overshoot| < n (gfedcb|abcdefgh|gfedcba)
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
__device__ __forceinline__ int clampi(int i, int lo, int hi) { return max(lo, min(i, hi)); }

}  // namespace tf
