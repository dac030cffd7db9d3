// Multi-GPU flow hand-off plumbing: CUDA IPC handles, release/acquire flags and a peer-store
// copy kernel.  Frame pairs are sharded across ranks (one process per GPU); a producer rank
// stores its finished flow straight into the accumulator rank's ring slot over NVLink and then
// raises a flag; the accumulator's stream waits on the flag with a driver stream-memory op, so
// no kernel ever spins on memory another rank writes.
#include "common.cuh"
#include <cuda.h>

using namespace tf;

extern "C" int tf_device_malloc(size_t bytes, void** dev_ptr_out) {
    TF_REQUIRE(dev_ptr_out && bytes > 0, TF_ERR_INVALID_ARG, "tf_device_malloc: bad argument");
    TF_CUDA(cudaMalloc(dev_ptr_out, bytes));
    TF_CUDA(cudaMemset(*dev_ptr_out, 0, bytes));
    return TF_OK;
}

extern "C" int tf_device_free(void* dev_ptr) {
    if (dev_ptr) TF_CUDA(cudaFree(dev_ptr));
    return TF_OK;
}

extern "C" int tf_ipc_get_handle(const void* dev_ptr, uint8_t handle_out[64]) {
    TF_REQUIRE(dev_ptr && handle_out, TF_ERR_INVALID_ARG, "tf_ipc_get_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    cudaIpcMemHandle_t hnd;
    TF_CUDA(cudaIpcGetMemHandle(&hnd, const_cast<void*>(dev_ptr)));
    memcpy(handle_out, &hnd, 64);
    return TF_OK;
}

extern "C" int tf_ipc_open_handle(const uint8_t handle[64], void** dev_ptr_out) {
    TF_REQUIRE(handle && dev_ptr_out, TF_ERR_INVALID_ARG, "tf_ipc_open_handle: null argument");
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle, 64);
    TF_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, hnd, cudaIpcMemLazyEnablePeerAccess));
    return TF_OK;
}

extern "C" int tf_ipc_close_handle(void* dev_ptr) {
    if (!dev_ptr) return TF_OK;
    TF_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return TF_OK;
}

__global__ void k_flag_signal(volatile uint32_t* flag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

extern "C" int tf_flag_signal(uint32_t* flag, uint32_t value, void* stream) {
    TF_REQUIRE(flag, TF_ERR_INVALID_ARG, "tf_flag_signal: null flag");
    // stream order guarantees every earlier kernel on `stream` has completed (and its peer
    // stores were issued); the system-scope fence + release store publish them to the peer.
    k_flag_signal<<<1, 1, 0, as_stream(stream)>>>(flag, value);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_flag_wait_geq(uint32_t* flag, uint32_t value, void* stream) {
    TF_REQUIRE(flag, TF_ERR_INVALID_ARG, "tf_flag_wait_geq: null flag");
    // resolved through the runtime so the library has no link-time dependency on libcuda.so
    typedef CUresult (*wait_fn_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static wait_fn_t wait_fn = nullptr;
    if (!wait_fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        TF_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &q));
        TF_REQUIRE(sym && q == cudaDriverEntryPointSuccess, TF_ERR_CUDA, "cuStreamWaitValue32 is not available");
        wait_fn = reinterpret_cast<wait_fn_t>(sym);
    }
    CUresult r = wait_fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(flag), value,
                         CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
        return fail(TF_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult %d", (int)r);
    }
    return TF_OK;
}

// 128-bit grid-stride copy; dst may be a peer (NVLink) mapping.  Grid sized to the SM count.
__global__ void __launch_bounds__(256) k_copy_to_peer(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16,
                                                      uint8_t* dst_tail, const uint8_t* src_tail, int tail) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = __ldg(src + i);
    if (blockIdx.x == 0 && threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

extern "C" int tf_copy_to_peer(void* dst, const void* src, size_t bytes, void* stream) {
    TF_REQUIRE(dst && src, TF_ERR_INVALID_ARG, "tf_copy_to_peer: null argument");
    TF_REQUIRE((((uintptr_t)dst | (uintptr_t)src) & 15) == 0, TF_ERR_INVALID_ARG, "tf_copy_to_peer: 16-byte alignment");
    if (int e = require_sm100()) return e;
    size_t n16 = bytes / 16;
    int tail = (int)(bytes - n16 * 16);
    int blocks = sm_count() * 4;
    k_copy_to_peer<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<uint4*>(dst),
                                                         reinterpret_cast<const uint4*>(src), n16,
                                                         reinterpret_cast<uint8_t*>(dst) + n16 * 16,
                                                         reinterpret_cast<const uint8_t*>(src) + n16 * 16, tail);
    TF_LAUNCHED();
    return TF_OK;
}
