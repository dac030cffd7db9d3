// Library globals, device checks, frame prep (BGR->gray) and flow post-processing kernels.
#include "common.cuh"

namespace tf {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

static int g_arch_checked_device = -1;
static int g_arch_status = TF_OK;
static int g_sm_count = 148;

int require_sm100() {
    int dev = 0;
    TF_CUDA(cudaGetDevice(&dev));
    if (dev == g_arch_checked_device) {
        if (g_arch_status != TF_OK)
            return fail(g_arch_status, "device %d is not an sm_100 (B200) part; no fallback path exists", dev);
        return TF_OK;
    }
    cudaDeviceProp prop;
    TF_CUDA(cudaGetDeviceProperties(&prop, dev));
    g_arch_checked_device = dev;
    g_sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
        g_arch_status = TF_ERR_UNSUPPORTED_ARCH;
        return fail(TF_ERR_UNSUPPORTED_ARCH,
                    "device %d (%s, sm_%d%d) is not an sm_100 (B200) part; no fallback path exists",
                    dev, prop.name, prop.major, prop.minor);
    }
    g_arch_status = TF_OK;
    return TF_OK;
}
int sm_count() { return g_sm_count; }
}  // namespace tf

// ---- kernel timers -------------------------------------------------------------------------------
#include <vector>
namespace tf {
struct TimerSlot {
    std::vector<cudaEvent_t> begin, end;
    size_t used = 0;
};
static bool g_timing = false;
static TimerSlot g_timers[TFK_COUNT];

void timer_begin(int tag, cudaStream_t st) {
    if (!g_timing || tag < 0) return;
    TimerSlot& t = g_timers[tag];
    if (t.used == t.begin.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        t.begin.push_back(a);
        t.end.push_back(b);
    }
    cudaEventRecord(t.begin[t.used], st);
}
void timer_end(int tag, cudaStream_t st) {
    if (!g_timing || tag < 0) return;
    TimerSlot& t = g_timers[tag];
    if (t.used < t.end.size()) {
        cudaEventRecord(t.end[t.used], st);
        t.used++;
    }
}
}  // namespace tf

using namespace tf;

extern "C" int tf_timer_enable(int on) {
    g_timing = on != 0;
    if (on)
        for (auto& t : g_timers) t.used = 0;
    return TF_OK;
}

extern "C" int tf_timer_read(int tag, double* total_ms, uint64_t* launches) {
    TF_REQUIRE(tag >= 0 && tag < TFK_COUNT, TF_ERR_INVALID_ARG, "tf_timer_read: bad tag %d", tag);
    TF_CUDA(cudaDeviceSynchronize());
    TimerSlot& t = g_timers[tag];
    double sum = 0;
    for (size_t i = 0; i < t.used; i++) {
        float ms = 0.f;
        TF_CUDA(cudaEventElapsedTime(&ms, t.begin[i], t.end[i]));
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (launches) *launches = t.used;
    return TF_OK;
}

extern "C" int tf_version(void) { return 100; }
extern "C" const char* tf_last_error(void) { return tf::g_err; }
extern "C" uint64_t tf_launch_count(void) { return tf::g_launches.load(); }

extern "C" int tf_device_check(int device, int* sms) {
    int count = 0;
    TF_CUDA(cudaGetDeviceCount(&count));
    TF_REQUIRE(device >= 0 && device < count, TF_ERR_INVALID_ARG, "no CUDA device %d (have %d)", device, count);
    cudaDeviceProp prop;
    TF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sms) *sms = prop.multiProcessorCount;
    TF_REQUIRE(prop.major == 10, TF_ERR_UNSUPPORTED_ARCH, "device %d (%s, sm_%d%d) is not sm_100", device,
               prop.name, prop.major, prop.minor);
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// BGR -> gray, bit-exact with cv2.cvtColor(BGR2GRAY) (flow/sources/cv.py:465):
//   (3735*B + 19235*G + 9798*R + 16384) >> 15
// 4 pixels per thread: three 32-bit loads (12 bytes) -> one 32-bit store.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gray1(uint32_t b, uint32_t g, uint32_t r) {
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) k_gray_from_bgr(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
                                                       size_t n) {
    size_t quad = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nq = n >> 2;
    if (quad < nq) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr) + quad * 3;
        uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        uint32_t g0 = gray1(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255);
        uint32_t g1 = gray1(w0 >> 24, w1 & 255, (w1 >> 8) & 255);
        uint32_t g2 = gray1((w1 >> 16) & 255, w1 >> 24, w2 & 255);
        uint32_t g3 = gray1((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24);
        reinterpret_cast<uint32_t*>(gray)[quad] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else if (quad == nq) {
        for (size_t i = nq * 4; i < n; i++) gray[i] = (uint8_t)gray1(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
    }
}

extern "C" int tf_gray_from_bgr(const uint8_t* bgr, uint8_t* gray, int height, int width, void* stream) {
    TF_REQUIRE(bgr && gray, TF_ERR_INVALID_ARG, "tf_gray_from_bgr: null buffer");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_gray_from_bgr: bad shape %dx%d", height, width);
    TF_REQUIRE(((uintptr_t)bgr & 3) == 0 && ((uintptr_t)gray & 3) == 0, TF_ERR_INVALID_ARG,
               "tf_gray_from_bgr: buffers must be 4-byte aligned");
    if (int e = require_sm100()) return e;
    size_t n = (size_t)height * width;
    size_t threads = (n >> 2) + 1;
    k_gray_from_bgr<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(bgr, gray, n);
    TF_LAUNCHED();
    return TF_OK;
}

// cv2.resize(frame, dsize, interpolation=INTER_NEAREST) of a BGR frame (flow/sources/cv.py:464): OpenCV's resizeNN takes
// source column min(floor(x * ifx), sw - 1) with ifx = 1 / (dw / sw) in double, and the same for rows.
__global__ void __launch_bounds__(256) k_resize_nearest_c3(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           int sh, int sw, int dh, int dw, double ifx, double ify) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int sx = min(__double2int_rd((double)x * ifx), sw - 1);
    const int sy = min(__double2int_rd((double)y * ify), sh - 1);
    const uint8_t* p = src + ((size_t)sy * sw + sx) * 3;
    uint8_t* q = dst + ((size_t)y * dw + x) * 3;
    q[0] = __ldg(p);
    q[1] = __ldg(p + 1);
    q[2] = __ldg(p + 2);
}

extern "C" int tf_resize_nearest_bgr(const uint8_t* src, int src_height, int src_width, uint8_t* dst, int height, int width,
                                     void* stream) {
    TF_REQUIRE(src && dst, TF_ERR_INVALID_ARG, "tf_resize_nearest_bgr: null buffer");
    TF_REQUIRE(src_height > 0 && src_width > 0 && height > 0 && width > 0, TF_ERR_SHAPE,
               "tf_resize_nearest_bgr: bad shape %dx%d -> %dx%d", src_width, src_height, width, height);
    if (int e = require_sm100()) return e;
    const double ifx = 1.0 / ((double)width / (double)src_width), ify = 1.0 / ((double)height / (double)src_height);
    k_resize_nearest_c3<<<dim3(ceil_div(width, 256), height), 256, 0, as_stream(stream)>>>(src, dst, src_height, src_width,
                                                                                          height, width, ifx, ify);
    TF_LAUNCHED();
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// FlowSource.post_process (flow/sources/source.py:337-363)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 clip_flow(float2 f, int x, int y, int w, int h) {
    // numpy.clip(flow, -x, W-1-x) with int32 bounds: float32 result
    f.x = fminf(fmaxf(f.x, (float)(-x)), (float)(w - 1 - x));
    f.y = fminf(fmaxf(f.y, (float)(-y)), (float)(h - 1 - y));
    return f;
}

// Flow filters (transflow/flow/filters.py:36-68) as a short per-pixel program; NumPy's promotion rules decide
// the arithmetic type: a Python scalar threshold is "weak" (everything stays float32), a NumPy float64 scalar
// is "strong" (compare / scale in float64, result rounded to float32).
struct FlowOps {
    int n;
    int kind[TF_MAX_FLOW_OPS];    // TF_FLOW_SCALE / TF_FLOW_THRESHOLD / TF_FLOW_CLIP
    int strong[TF_MAX_FLOW_OPS];
    float v32[TF_MAX_FLOW_OPS];
    double v64[TF_MAX_FLOW_OPS];
};

__device__ __forceinline__ float2 apply_flow_ops(float2 f, const FlowOps& ops) {
    for (int i = 0; i < ops.n; i++) {
        const int kind = ops.kind[i];
        if (kind == TF_FLOW_SCALE) {  // filters.py:41  flow *= expr(t)
            if (ops.strong[i]) {
                f.x = (float)((double)f.x * ops.v64[i]);
                f.y = (float)((double)f.y * ops.v64[i]);
            } else {
                f.x = __fmul_rn(f.x, ops.v32[i]);
                f.y = __fmul_rn(f.y, ops.v32[i]);
            }
            continue;
        }
        // numpy.linalg.norm(axis=1) on float32: sqrt(x*x + y*y), every step rounded to float32
        const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)));
        if (kind == TF_FLOW_THRESHOLD) {  // filters.py:49-53  flow[norm <= threshold] = 0
            const bool hit = ops.strong[i] ? (double)norm <= ops.v64[i] : norm <= ops.v32[i];
            if (hit) f = make_float2(0.f, 0.f);
        } else {  // filters.py:61-68  flow *= threshold / norm where norm >= threshold (factors are float64)
            if (ops.strong[i]) {
                if ((double)norm >= ops.v64[i]) {
                    const double fac = ops.v64[i] / (double)norm;
                    f.x = (float)((double)f.x * fac);
                    f.y = (float)((double)f.y * fac);
                }
            } else if (norm >= ops.v32[i]) {
                const float fac = __fdiv_rn(ops.v32[i], norm);
                f.x = __fmul_rn(f.x, fac);  // float32 x float32 in float64 is exact: one rounding, like NumPy
                f.y = __fmul_rn(f.y, fac);
            }
        }
    }
    return f;
}

static int pack_flow_ops(const tf_flow_op* ops, int n_ops, FlowOps& out) {
    TF_REQUIRE(n_ops >= 0 && n_ops <= TF_MAX_FLOW_OPS, TF_ERR_INVALID_ARG, "at most %d flow filters per call (got %d)",
               TF_MAX_FLOW_OPS, n_ops);
    TF_REQUIRE(n_ops == 0 || ops, TF_ERR_INVALID_ARG, "null flow filter list");
    memset(&out, 0, sizeof(out));
    out.n = n_ops;
    for (int i = 0; i < n_ops; i++) {
        TF_REQUIRE(ops[i].kind >= TF_FLOW_SCALE && ops[i].kind <= TF_FLOW_CLIP, TF_ERR_INVALID_ARG,
                   "unknown flow filter kind %d", ops[i].kind);
        out.kind[i] = ops[i].kind;
        out.strong[i] = ops[i].strong != 0;
        out.v32[i] = (float)ops[i].value;
        out.v64[i] = ops[i].value;
    }
    return TF_OK;
}

// filters + mask only (the stage in front of the convolution kernel, source.py:339-343)
__global__ void __launch_bounds__(256) k_flow_filters(const float2* __restrict__ flow, const float* __restrict__ mask,
                                                      float2* __restrict__ out, size_t n, const FlowOps ops) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float2 f = apply_flow_ops(flow[p], ops);
    if (mask) {
        float m = mask[p];
        f.x = __fmul_rn(m, f.x);
        f.y = __fmul_rn(m, f.y);
    }
    out[p] = f;
}

// scipy.signal.convolve2d(flow[..., c], kernel, mode="same", boundary="fill", fillvalue=0) (source.py:344-348):
// float64 accumulation in kernel-index order, 'same' window centred at (k - 1) / 2.  The reference carries the
// float64 result on; the public flow type is float32, so the value is rounded once -- except in the forward
// direction, where only round-half-even of the clipped value survives (source.py:352): there the float32 value
// is nudged to the float64 rounding result whenever the two would disagree.
__global__ void __launch_bounds__(256) k_flow_convolve(const float2* __restrict__ flow, const double* __restrict__ ker,
                                                       float2* __restrict__ out, int h, int w, int kh, int kw,
                                                       int forward) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    const int oy = (kh - 1) / 2, ox = (kw - 1) / 2;
    double sx = 0.0, sy = 0.0;
    for (int j = 0; j < kh; j++) {
        const int yy = y + oy - j;
        for (int k = 0; k < kw; k++) {
            const int xx = x + ox - k;
            float2 v = make_float2(0.f, 0.f);
            if ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w) v = __ldg(flow + (size_t)yy * w + xx);
            const double c = ker[j * kw + k];
            sx = __dadd_rn(sx, __dmul_rn((double)v.x, c));
            sy = __dadd_rn(sy, __dmul_rn((double)v.y, c));
        }
    }
    float2 f = make_float2((float)sx, (float)sy);
    if (forward) {
        const double lox = -x, hix = w - 1 - x, loy = -y, hiy = h - 1 - y;
        const double rx = rint(fmin(fmax(sx, lox), hix)), ry = rint(fmin(fmax(sy, loy), hiy));
        if (rintf(fminf(fmaxf(f.x, (float)lox), (float)hix)) != (float)rx) f.x = (float)rx;
        if (rintf(fminf(fmaxf(f.y, (float)loy), (float)hiy)) != (float)ry) f.y = (float)ry;
    }
    out[(size_t)y * w + x] = f;
}

// backward direction: optional filters + mask multiply + final clip.
__global__ void __launch_bounds__(256) k_post_backward(const float2* __restrict__ flow, const float* __restrict__ mask,
                                                       float2* __restrict__ out, int h, int w, const FlowOps ops) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t p = (size_t)y * w + x;
    float2 f = apply_flow_ops(flow[p], ops);
    if (mask) {
        float m = mask[p];
        f.x = __fmul_rn(m, f.x);
        f.y = __fmul_rn(m, f.y);
    }
    out[p] = clip_flow(f, x, y, w, h);
}

// forward pass 1: clip, round half-even, claim the target with atomicMax(source index + 1):
// numpy.put with duplicate targets keeps the LAST source in raster order (quirk Q7).
__global__ void __launch_bounds__(256) k_post_forward_scatter(const float2* __restrict__ flow,
                                                              const float* __restrict__ mask,
                                                              int* __restrict__ owner, int h, int w,
                                                              const FlowOps ops) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int p = y * w + x;
    float2 f = apply_flow_ops(flow[p], ops);
    if (mask) {
        float m = mask[p];
        f.x = __fmul_rn(m, f.x);
        f.y = __fmul_rn(m, f.y);
    }
    f = clip_flow(f, x, y, w, h);
    int fx = __float2int_rn(f.x), fy = __float2int_rn(f.y);
    int off = fy * w + fx;
    if (off != 0) {
        int q = min(max(p + off, 0), h * w - 1);  // numpy.put(mode="clip")
        atomicMax(owner + q, p + 1);
    }
}

// forward pass 2: flow := owner position - own position (0 where unclaimed); owner plane re-zeroed.
__global__ void __launch_bounds__(256) k_post_forward_gather(float2* __restrict__ out, int* __restrict__ owner,
                                                             int h, int w, float inv_w) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    int p = y * w + x;
    int o = owner[p];
    float2 f = make_float2(0.f, 0.f);
    if (o != 0) {
        int s = o - 1;
        owner[p] = 0;
        // s = sy * w + sx without an integer division: float estimate of the quotient (off by at most one: s < 2^30
        // rounds to float within 2^-24 relative), corrected exactly in integers
        int sy = __float2int_rz(__fmul_rn((float)s, inv_w));
        int sx = s - sy * w;
        if (sx < 0) { sy--; sx += w; }
        if (sx >= w) { sy++; sx -= w; }
        f.x = (float)(sx - x);
        f.y = (float)(sy - y);
    }
    out[p] = f;  // already inside the frame: the final clip is the identity here
}

extern "C" int tf_flow_postprocess_ex(float* flow, const tf_flow_op* ops, int n_ops, const float* mask, int forward,
                                      int32_t* owner, float* out, int height, int width, void* stream) {
    TF_REQUIRE(flow, TF_ERR_INVALID_ARG, "tf_flow_postprocess: null flow");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_postprocess: bad shape %dx%d", height, width);
    TF_REQUIRE(!forward || owner, TF_ERR_INVALID_ARG, "tf_flow_postprocess: forward direction needs the owner plane");
    if (int e = require_sm100()) return e;
    FlowOps packed;
    if (int e = pack_flow_ops(ops, n_ops, packed)) return e;
    if (!out) out = flow;
    dim3 grid(ceil_div(width, 256), height);
    cudaStream_t st = as_stream(stream);
    if (!forward) {
        k_post_backward<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(flow), mask, reinterpret_cast<float2*>(out),
                                              height, width, packed);
        TF_LAUNCHED();
    } else {
        ScopedKernelTimer timer(TFK_POST_FORWARD, st);
        k_post_forward_scatter<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(flow), mask, owner, height, width,
                                                     packed);
        TF_LAUNCHED();
        k_post_forward_gather<<<grid, 256, 0, st>>>(reinterpret_cast<float2*>(out), owner, height, width,
                                                    1.0f / (float)width);
        TF_LAUNCHED();
    }
    return TF_OK;
}

// The two passes of the forward direction on their own: a consumer that only needs WHICH pixel lands on each target
// (the move-reference compositor layer, tf_layer_update_claims) takes the claim plane of the scatter pass and the
// flow of the gather pass is never formed.
extern "C" int tf_flow_forward_claims(const float* flow, const tf_flow_op* ops, int n_ops, const float* mask,
                                      int32_t* claims, int height, int width, void* stream) {
    TF_REQUIRE(flow && claims, TF_ERR_INVALID_ARG, "tf_flow_forward_claims: null argument");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_forward_claims: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    FlowOps packed;
    if (int e = pack_flow_ops(ops, n_ops, packed)) return e;
    cudaStream_t st = as_stream(stream);
    ScopedKernelTimer timer(TFK_POST_FORWARD, st);
    k_post_forward_scatter<<<dim3(ceil_div(width, 256), height), 256, 0, st>>>(reinterpret_cast<const float2*>(flow), mask,
                                                                              claims, height, width, packed);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_flow_from_claims(int32_t* claims, float* flow_out, int height, int width, void* stream) {
    TF_REQUIRE(flow_out && claims, TF_ERR_INVALID_ARG, "tf_flow_from_claims: null argument");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_from_claims: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    cudaStream_t st = as_stream(stream);
    k_post_forward_gather<<<dim3(ceil_div(width, 256), height), 256, 0, st>>>(reinterpret_cast<float2*>(flow_out), claims,
                                                                             height, width, 1.0f / (float)width);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_flow_postprocess_to(float* flow, const float* mask, int forward, int32_t* owner, float* out,
                                      int height, int width, void* stream) {
    return tf_flow_postprocess_ex(flow, nullptr, 0, mask, forward, owner, out, height, width, stream);
}

extern "C" int tf_flow_filters(const float* flow, const tf_flow_op* ops, int n_ops, const float* mask, float* out,
                               int height, int width, void* stream) {
    TF_REQUIRE(flow && out, TF_ERR_INVALID_ARG, "tf_flow_filters: null buffer");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_filters: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    FlowOps packed;
    if (int e = pack_flow_ops(ops, n_ops, packed)) return e;
    size_t n = (size_t)height * width;
    k_flow_filters<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(flow), mask, reinterpret_cast<float2*>(out), n, packed);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_flow_convolve(const float* flow, const double* kernel, int kh, int kw, int forward, float* out,
                                int height, int width, void* stream) {
    TF_REQUIRE(flow && kernel && out, TF_ERR_INVALID_ARG, "tf_flow_convolve: null buffer");
    TF_REQUIRE(flow != out, TF_ERR_INVALID_ARG, "tf_flow_convolve: cannot run in place");
    TF_REQUIRE(height > 0 && width > 0 && kh > 0 && kw > 0, TF_ERR_SHAPE, "tf_flow_convolve: bad shape");
    if (int e = require_sm100()) return e;
    k_flow_convolve<<<dim3(ceil_div(width, 256), height), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(flow), kernel, reinterpret_cast<float2*>(out), height, width, kh, kw, forward);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_flow_postprocess(float* flow, const float* mask, int forward, int32_t* owner, int height, int width,
                                   void* stream) {
    return tf_flow_postprocess_to(flow, mask, forward, owner, nullptr, height, width, stream);
}

// ------------------------------------------------------------------------------------------
// Pipeline._update_flow (pipeline.py:492-507): merge of several flow sources (pipeline.py:149-158,
// utils.py:359-381) and integer upscale (utils.py:417-418), float32 like NumPy evaluates them
// ------------------------------------------------------------------------------------------
struct MergeArgs {
    const float* src[TF_MAX_MERGE_FLOWS];
    int n;
};

__global__ void __launch_bounds__(256) k_flow_merge(const MergeArgs a, int mode, float* __restrict__ out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float v = __ldg(a.src[0] + i);
    switch (mode) {
        case TF_MERGE_FIRST: break;
        case TF_MERGE_SUM:
        case TF_MERGE_AVERAGE:
            for (int k = 1; k < a.n; k++) v = __fadd_rn(v, __ldg(a.src[k] + i));
            if (mode == TF_MERGE_AVERAGE) v = __fdiv_rn(v, (float)a.n);
            break;
        case TF_MERGE_DIFFERENCE: {  // flows[0] - sum(flows[1:]), Python's sum starts from the int 0
            if (a.n > 1) {
                float s = __ldg(a.src[1] + i);
                for (int k = 2; k < a.n; k++) s = __fadd_rn(s, __ldg(a.src[k] + i));
                v = __fsub_rn(v, s);
            }
            break;
        }
        case TF_MERGE_PRODUCT:
            for (int k = 1; k < a.n; k++) v = __fmul_rn(v, __ldg(a.src[k] + i));
            break;
        case TF_MERGE_MASKBIN:  // utils.py:368-373: |f| > 0.2 -> 1 else 0 (0.2 is weak: compared as float32)
            for (int k = 1; k < a.n; k++) v = __fmul_rn(v, fabsf(__ldg(a.src[k] + i)) > 0.2f ? 1.f : 0.f);
            break;
        case TF_MERGE_MASKLIN:
            for (int k = 1; k < a.n; k++) v = __fmul_rn(v, fabsf(__ldg(a.src[k] + i)));
            break;
        case TF_MERGE_ABSMAX: {  // utils.py:376-381: two flows, argmax of |.| (first wins ties)
            float o = __ldg(a.src[1] + i);
            if (fabsf(o) > fabsf(v)) v = o;
            break;
        }
    }
    out[i] = v;
}

extern "C" int tf_flow_merge(const float* const* flows, int n, int mode, float* out, int height, int width, void* stream) {
    TF_REQUIRE(flows && out, TF_ERR_INVALID_ARG, "tf_flow_merge: null argument");
    TF_REQUIRE(n >= 1 && n <= TF_MAX_MERGE_FLOWS, TF_ERR_INVALID_ARG, "tf_flow_merge: %d flows (1..%d)", n,
               TF_MAX_MERGE_FLOWS);
    TF_REQUIRE(mode >= TF_MERGE_FIRST && mode <= TF_MERGE_ABSMAX, TF_ERR_INVALID_ARG, "tf_flow_merge: unknown mode %d", mode);
    TF_REQUIRE(mode != TF_MERGE_ABSMAX || n == 2, TF_ERR_INVALID_ARG,
               "absmax merges exactly two flows (utils.py:378 reshapes to (2, ...)), got %d", n);
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_merge: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    MergeArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    for (int k = 0; k < n; k++) {
        TF_REQUIRE(flows[k], TF_ERR_INVALID_ARG, "tf_flow_merge: flow %d is null", k);
        a.src[k] = flows[k];
    }
    size_t count = (size_t)height * width * 2;
    k_flow_merge<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(a, mode, out, count);
    TF_LAUNCHED();
    return TF_OK;
}

__global__ void __launch_bounds__(256) k_flow_upscale(const float2* __restrict__ in, float2* __restrict__ out, int h, int w,
                                                      int wf, int hf) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;  // output coordinates
    int y = blockIdx.y;
    if (x >= w * wf) return;
    float2 v = __ldg(in + (size_t)(y / hf) * w + x / wf);
    out[(size_t)y * (w * wf) + x] = make_float2(__fmul_rn(v.x, (float)wf), __fmul_rn(v.y, (float)hf));
}

extern "C" int tf_flow_upscale(const float* flow, float* out, int height, int width, int wf, int hf, void* stream) {
    TF_REQUIRE(flow && out && flow != out, TF_ERR_INVALID_ARG, "tf_flow_upscale: bad buffers");
    TF_REQUIRE(height > 0 && width > 0 && wf >= 1 && hf >= 1, TF_ERR_SHAPE, "tf_flow_upscale: bad shape / factors");
    if (int e = require_sm100()) return e;
    k_flow_upscale<<<dim3(ceil_div(width * wf, 256), height * hf), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(flow), reinterpret_cast<float2*>(out), height, width, wf, hf);
    TF_LAUNCHED();
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// Flow visualisers render2d / render1d (transflow/output/render.py:9-48, called at
// pipeline.py:512-516 when view_flow / view_flow_magnitude is set): float32 arithmetic in NumPy's
// evaluation order (no contraction), clip to [0, 255], truncating cast to uint8.
//   mode 0  render2d(flow)                       4 colours (y, b, m, g)
//   mode 1  render1d(sqrt(fx^2 + fy^2))          2 colours, the magnitude of pipeline.py:515 fused in
//   mode 2  render1d(arr) for a scalar (H, W) array
// ------------------------------------------------------------------------------------------
struct RenderArgs {
    float c[4][3];
    float scale;
    int mode, binary;
};

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__global__ void __launch_bounds__(256) k_flow_render(const float* __restrict__ in, uint8_t* __restrict__ rgb, size_t n,
                                                     const RenderArgs a) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float out[3];
    if (a.mode == 0) {
        float2 f = __ldg(reinterpret_cast<const float2*>(in) + p);
        float sx = __fmul_rn(a.scale, f.x), sy = __fmul_rn(a.scale, f.y);
        float cy = clip01(__fadd_rn(1.f, sx)), cb = clip01(__fsub_rn(1.f, sx));
        float cm = clip01(__fadd_rn(1.f, sy)), cg = clip01(__fsub_rn(1.f, sy));
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float s = __fadd_rn(__fmul_rn(cy, a.c[0][k]), __fmul_rn(cb, a.c[1][k]));
            s = __fadd_rn(s, __fmul_rn(cm, a.c[2][k]));
            s = __fadd_rn(s, __fmul_rn(cg, a.c[3][k]));
            out[k] = __fmul_rn(0.5f, s);
        }
    } else {
        float v;
        if (a.mode == 1) {
            float2 f = __ldg(reinterpret_cast<const float2*>(in) + p);
            v = __fsqrt_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)));
        } else {
            v = __ldg(in + p);
        }
        float s = __fmul_rn(a.scale, v), ca, cb;
        if (a.binary) {
            cb = clip01(rintf(s));  // numpy.round: half to even
            ca = __fsub_rn(1.f, cb);
        } else {
            ca = clip01(__fsub_rn(1.f, s));
            cb = clip01(s);
        }
#pragma unroll
        for (int k = 0; k < 3; k++) out[k] = __fadd_rn(__fmul_rn(ca, a.c[0][k]), __fmul_rn(cb, a.c[1][k]));
    }
    uint8_t* dst = rgb + 3 * p;
#pragma unroll
    for (int k = 0; k < 3; k++) dst[k] = (uint8_t)(int)fminf(fmaxf(out[k], 0.f), 255.f);
}

extern "C" int tf_flow_render(const float* in, int mode, float scale, const float* colors, int n_colors, int binary,
                              uint8_t* rgb, int height, int width, void* stream) {
    TF_REQUIRE(in && rgb && colors, TF_ERR_INVALID_ARG, "tf_flow_render: null argument");
    TF_REQUIRE(mode >= 0 && mode <= 2, TF_ERR_INVALID_ARG, "tf_flow_render: unknown mode %d", mode);
    TF_REQUIRE(n_colors == (mode == 0 ? 4 : 2), TF_ERR_INVALID_ARG, "tf_flow_render: mode %d takes %d colours, got %d",
               mode, mode == 0 ? 4 : 2, n_colors);
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_flow_render: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    RenderArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_colors; i++)
        for (int k = 0; k < 3; k++) a.c[i][k] = colors[3 * i + k];
    a.scale = scale;
    a.mode = mode;
    a.binary = binary != 0;
    size_t n = (size_t)height * width;
    k_flow_render<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(in, rgb, n, a);
    TF_LAUNCHED();
    return TF_OK;
}
