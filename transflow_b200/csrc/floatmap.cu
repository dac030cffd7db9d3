// Float displacement-map accumulation + bilinear / nearest remap: the reference's WebGL
// variant (extra/www/shaders/acc.frag:17-41, remap.frag:10-18, sampler state transflow.js:358-364), in pixel
// units, as restated in oracle/floatmap_np.py.  An opt-in extension: the Python reference composes integer maps
// with nearest remap only (compositor.cu).  Both kernels are gathers bounded by HBM bandwidth:
//   accumulate: read flow 8 + gather map 8 (x4 taps, L1/L2 hits) + write map 8 = 24 B/px (blur_size 1)
//   remap     : read map 8 + gather pixmap 3|4 (x4 taps) + write RGB 3 = 14-15 B/px
#include "common.cuh"

using namespace tf;

namespace {

// texture2D at texel-index coordinates, CLAMP_TO_EDGE; LINEAR weights in fp32 like the restatement
template <typename T, typename LoadF>
__device__ __forceinline__ T sample_tex(LoadF load, float x, float y, int w, int h, bool linear) {
    if (!linear) {
        int xi = clampi(__float2int_rd(x + 0.5f), 0, w - 1), yi = clampi(__float2int_rd(y + 0.5f), 0, h - 1);
        return load(yi, xi);
    }
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    // saturating conversions keep absurd coordinates finite before the clamp
    int xa = __float2int_rd(x0f), ya = __float2int_rd(y0f);
    int x0 = clampi(xa, 0, w - 1), x1 = clampi(xa < 0x7fffffff ? xa + 1 : xa, 0, w - 1);
    int y0 = clampi(ya, 0, h - 1), y1 = clampi(ya < 0x7fffffff ? ya + 1 : ya, 0, h - 1);
    T top = load(y0, x0) * (1.f - fx) + load(y0, x1) * fx;
    T bot = load(y1, x0) * (1.f - fx) + load(y1, x1) * fx;
    return top * (1.f - fy) + bot * fy;
}

struct F2 {
    float x, y;
    __device__ F2 operator*(float s) const { return {x * s, y * s}; }
    __device__ F2 operator+(F2 o) const { return {x + o.x, y + o.y}; }
};
struct F4 {
    float x, y, z, w;
    __device__ F4 operator*(float s) const { return {x * s, y * s, z * s, w * s}; }
    __device__ F4 operator+(F4 o) const { return {x + o.x, y + o.y, z + o.z, w + o.w}; }
};

__global__ void __launch_bounds__(256) k_floatmap_accumulate(const float2* __restrict__ map_prev,
                                                             const float2* __restrict__ flow, float2* __restrict__ out,
                                                             int h, int w, float scale, float decay, int blur,
                                                             int linear) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    auto ld_flow = [&](int yy, int xx) { float2 v = __ldg(flow + (size_t)yy * w + xx); return F2{v.x, v.y}; };
    auto ld_map = [&](int yy, int xx) { float2 v = __ldg(map_prev + (size_t)yy * w + xx); return F2{v.x, v.y}; };
    const float half = (float)(blur / 2);
    const float weight = 1.0f / ((float)blur * (float)blur);
    F2 c{0.f, 0.f};
    for (int j = 0; j < blur; j++)          // acc.frag:24-35
        for (int i = 0; i < blur; i++) {
            F2 s = sample_tex<F2>(ld_flow, (float)x + ((float)j - half), (float)y + ((float)i - half), w, h, linear);
            c.x = __fadd_rn(c.x, __fmul_rn(s.x, weight));
            c.y = __fadd_rn(c.y, __fmul_rn(s.y, weight));
        }
    F2 f{__fmul_rn(scale, c.x), __fmul_rn(scale, c.y)};
    F2 m = sample_tex<F2>(ld_map, __fadd_rn((float)x, f.x), __fadd_rn((float)y, f.y), w, h, linear);
    float ux = __fadd_rn(m.x, f.x), uy = __fadd_rn(m.y, f.y);                    // acc.frag:39
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    out[(size_t)y * w + x] = make_float2(__fsub_rn(ux, __fmul_rn(__fmul_rn(sgn(ux), decay), fabsf(ux))),   // :40
                                         __fsub_rn(uy, __fmul_rn(__fmul_rn(sgn(uy), decay), fabsf(uy))));
}

template <int CH>
__global__ void __launch_bounds__(256) k_floatmap_remap(const float2* __restrict__ map, const uint8_t* __restrict__ pix,
                                                        uint8_t* __restrict__ rgba, uint8_t* __restrict__ rgb, int h,
                                                        int w, int linear, int first_layer, uint32_t bg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t p = (size_t)y * w + x;
    float2 m = __ldg(map + p);
    auto ld = [&](int yy, int xx) {
        const uint8_t* q = pix + ((size_t)yy * w + xx) * CH;
        return F4{(float)q[0], (float)q[1], (float)q[2], CH == 4 ? (float)q[3] : 255.f};
    };
    F4 v = sample_tex<F4>(ld, __fadd_rn((float)x, m.x), __fadd_rn((float)y, m.y), w, h, linear);   // remap.frag:11-17
    auto q8 = [](float c) { return (uint8_t)fminf(fmaxf(floorf(c + 0.5f), 0.f), 255.f); };
    uint8_t r = q8(v.x), g = q8(v.y), b = q8(v.z), a = CH == 4 ? q8(v.w) : 255;
    if (rgba) {
        // layer image in the compositor's convention: alpha is a 0 / 1 flag (Layer.render, layer.py:32-34)
        rgba[4 * p] = r; rgba[4 * p + 1] = g; rgba[4 * p + 2] = b; rgba[4 * p + 3] = a ? 1 : 0;
    }
    if (rgb) {  // Compositor.render (compositor.py:31-40): opaque pixels overwrite, the first layer paints the background
        if (a) {
            rgb[3 * p] = r; rgb[3 * p + 1] = g; rgb[3 * p + 2] = b;
        } else if (first_layer) {
            rgb[3 * p] = (uint8_t)(bg >> 16); rgb[3 * p + 1] = (uint8_t)(bg >> 8); rgb[3 * p + 2] = (uint8_t)bg;
        }
    }
}

}  // namespace

extern "C" int tf_floatmap_accumulate(const float* map_prev, const float* flow, float* map_out, int height, int width,
                                      float scale, float decay, int blur_size, int linear, void* stream) {
    TF_REQUIRE(map_prev && flow && map_out, TF_ERR_INVALID_ARG, "tf_floatmap_accumulate: null buffer");
    TF_REQUIRE(map_prev != map_out, TF_ERR_INVALID_ARG, "tf_floatmap_accumulate: the map is gathered, it cannot be "
               "updated in place (ping-pong two buffers)");
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_floatmap_accumulate: bad shape %dx%d", height, width);
    TF_REQUIRE(blur_size >= 1 && blur_size <= 15, TF_ERR_INVALID_ARG, "blur size must be in [1, 15] (acc.frag:13), got %d",
               blur_size);
    if (int e = require_sm100()) return e;
    k_floatmap_accumulate<<<dim3(ceil_div(width, 256), height), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float2*>(map_prev), reinterpret_cast<const float2*>(flow),
        reinterpret_cast<float2*>(map_out), height, width, scale, decay, blur_size, linear != 0);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_floatmap_remap(const float* map, const uint8_t* pixmap, int channels, int linear, uint8_t* rgba_out,
                                 uint8_t* rgb_inout, int first_layer, uint32_t background_rgb, int height, int width,
                                 void* stream) {
    TF_REQUIRE(map && pixmap && (rgba_out || rgb_inout), TF_ERR_INVALID_ARG, "tf_floatmap_remap: null buffer");
    TF_REQUIRE(channels == 3 || channels == 4, TF_ERR_INVALID_ARG, "pixmap must have 3 or 4 channels, got %d", channels);
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_floatmap_remap: bad shape %dx%d", height, width);
    if (int e = require_sm100()) return e;
    dim3 grid(ceil_div(width, 256), height);
    cudaStream_t st = as_stream(stream);
    const float2* m = reinterpret_cast<const float2*>(map);
    if (channels == 3)
        k_floatmap_remap<3><<<grid, 256, 0, st>>>(m, pixmap, rgba_out, rgb_inout, height, width, linear != 0, first_layer,
                                                  background_rgb);
    else
        k_floatmap_remap<4><<<grid, 256, 0, st>>>(m, pixmap, rgba_out, rgb_inout, height, width, linear != 0, first_layer,
                                                  background_rgb);
    TF_LAUNCHED();
    return TF_OK;
}
