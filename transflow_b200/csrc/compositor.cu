// Compositor layers on device: map composition (moveref), sum, static, introduction, the reset
// modes, the integer nearest remap and the opaque-overwrite composite, as fused per-pixel
// gather kernels.  Reference: transflow/compositor/** (file:line cited at each stage).
//
// All of this is 32-bit integer / byte work bounded by HBM: one thread handles 4 consecutive
// pixels so that `data` moves as 128-bit records and the RGB frame as packed 32-bit words.
#include "common.cuh"

using namespace tf;

#define TF_MAX_SOURCES 8

struct tf_layer {
    tf_layer_config cfg;
    int h, w, depth;
    int4* data[2];   // ping-pong (the move is a gather from the previous state)
    int cur;
    // moveref layers also keep the records PACKED in 32 bits (i: 13, j: 13, alpha: 1, source: 5 -- frames up to
    // 8192 x 8192, 32 sources): the fast kernel gathers and writes 4 bytes per pixel instead of 16.  The int32 x 4
    // array of the reference (`layer.data`) is materialised from it on demand, and the other way round.
    uint32_t* pdata[2];
    int pcur;
    bool packed_ok, int4_valid, packed_valid;
    uchar4* rgba;    // persistent Layer.rgba (reference / static layers)
    uint8_t* mask_src;
    uint8_t* mask_dst;
    float* mask_alpha;
    float* reset_scale;
    int8_t* base_src;  // last source whose introduction mask covers the pixel, -1 = none
    uint8_t* intro[TF_MAX_SOURCES];
    int n_sources;
    uint32_t* vacated;  // frame stamp of the last time a pixel was the source of a move
    int* err;
    uint64_t frames;
    int introduced_once;
};

struct LayerParams {
    const float2* flow;
    const int4* old;
    int4* out;
    uchar4* rgba;
    const uint8_t* mask_src;
    const uint8_t* mask_dst;
    const float* mask_alpha;
    const float* reset_scale;
    const int8_t* base_src;
    uint32_t* vacated;
    const double* random;
    const uint8_t* pix[TF_MAX_SOURCES];
    const uint8_t* intro[TF_MAX_SOURCES];
    int chan[TF_MAX_SOURCES];
    int frame_no[TF_MAX_SOURCES];
    int n_src;
    uint8_t* rgb;
    int first_layer;
    uint32_t bg;
    int h, w, n;
    uint32_t stamp;
    uint64_t seed, frame;
    int* err;
    tf_layer_config cfg;
    int do_introduce;
};

// ---- Philox4x32-7 (counter = pixel pair, key = seed; frame in the counter): throughput-mode reset
// draws (7 rounds is the smallest variant that passes BigCrush; this is a visual effect, not
// cryptography).  One call yields 128 bits = two 53-bit uniforms, so a pixel pair shares one evaluation.
__device__ __forceinline__ uint4 philox4x32_7(uint64_t seed, uint64_t frame, uint32_t ctr) {
    uint32_t c0 = ctr, c1 = (uint32_t)frame, c2 = (uint32_t)(frame >> 32), c3 = 0x7f4a7c15u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 7; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double philox_uniform53(uint64_t seed, uint64_t frame, uint32_t pixel) {
    uint4 r = philox4x32_7(seed, frame, pixel >> 1);
    uint32_t a = (pixel & 1) ? r.z : r.x, b = (pixel & 1) ? r.w : r.y;
    // same 53-bit construction as numpy's random_sample: (a >> 5) * 2^26 + (b >> 6)
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ int wrap_index(int q, int n, int p, int* err) {
    // NumPy .flat[]: negative indices wrap once; anything else outside [0, n) raises IndexError.
    if (q < 0) q += n;
    if (q < 0 || q >= n) {
        *err = 1;
        q = p;
    }
    return q;
}

// Predicate of MovementLayer._update_move (movement.py:25-50): is p a target, and from where.
struct MoveDecision {
    int off, q;
    bool target, filled_target;
};

template <int ALPHA_IN_W>
__device__ __forceinline__ MoveDecision decide_move(const LayerParams& P, int p, int stride) {
    MoveDecision d;
    float2 f = __ldg(P.flow + p);
    d.off = __float2int_rn(f.y) * P.w + __float2int_rn(f.x);  // numpy.round = half-even
    d.q = p;
    d.target = d.filled_target = false;
    if (d.off == 0) return d;
    d.q = wrap_index(p + d.off, P.n, p, P.err);
    int4 rq = __ldg(P.old + (size_t)d.q * stride);
    int4 rp = __ldg(P.old + (size_t)p * stride);
    int aq = ALPHA_IN_W ? rq.w : rq.z, ap = ALPHA_IN_W ? rp.w : rp.z;
    bool src_alpha = aq != 0;
    bool ok = (P.mask_src == nullptr || P.mask_src[d.q] != 0) && (P.mask_dst == nullptr || P.mask_dst[p] != 0);
    if (!P.cfg.transparent_pixels_can_move) ok = ok && src_alpha;
    if (!P.cfg.pixels_can_move_to_empty_spot) ok = ok && ap != 0;
    if (!P.cfg.pixels_can_move_to_filled_spot) ok = ok && ap == 0;
    d.target = ok;
    d.filled_target = ok && (!P.cfg.transparent_pixels_can_move || src_alpha);
    return d;
}

// Pass A (only with moving_pixels_leave_empty_spot, movement.py:53-54): stamp the sources.
template <int ALPHA_IN_W>
__global__ void __launch_bounds__(256) k_mark_vacated(LayerParams P, int stride) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    MoveDecision d = decide_move<ALPHA_IN_W>(P, p, stride);
    if (d.target) P.vacated[d.q] = P.stamp;
}

// ReferenceLayer._update_reset* (reference.py:58-91) on one record.
__device__ __forceinline__ void apply_reset(const LayerParams& P, int p, int y, int x, int4& rec) {
    int mode = P.cfg.reset_mode;
    if (mode == TF_RESET_RANDOM) {
        double r = P.random ? __ldg(P.random + p) : philox_uniform53(P.seed, P.frame, (uint32_t)p);
        float thr = P.reset_scale ? __ldg(P.reset_scale + p) : P.cfg.reset_random_factor;
        if (r < (double)thr) {
            rec.x = y;
            rec.y = x;
            rec.z = 1;
            if (P.cfg.reset_source && P.base_src) {
                int s = P.base_src[p];
                if (s >= 0) rec.w = s;
            }
        }
    } else if (mode == TF_RESET_CONSTANT) {
        float d0y = (float)(y - rec.x), d0x = (float)(x - rec.y);
        float n0 = fmaxf(fabsf(d0y), fabsf(d0x));
        float dy = d0y, dx = d0x;
        if (n0 != 0.f) {
            dy = __fdiv_rn(dy, n0);
            dx = __fdiv_rn(dx, n0);
        }
        float s = P.reset_scale ? __ldg(P.reset_scale + p) : P.cfg.reset_constant_step;
        dy = __fmul_rn(dy, s);
        dx = __fmul_rn(dx, s);
        float n1 = fmaxf(fabsf(dy), fabsf(dx));
        if (n1 > n0) {
            dy = d0y;
            dx = d0x;
        }
        rec.x += __float2int_rn(dy);
        rec.y += __float2int_rn(dx);
    } else if (mode == TF_RESET_LINEAR) {
        double m = P.reset_scale ? (double)__ldg(P.reset_scale + p) : 1.0;
        double dy = __dmul_rn(P.cfg.reset_linear_factor, (double)(y - rec.x));
        double dx = __dmul_rn(P.cfg.reset_linear_factor, (double)(x - rec.y));
        rec.x += __double2int_rn(__dmul_rn(m, dy));
        rec.y += __double2int_rn(__dmul_rn(m, dx));
    }
}

// ReferenceLayer._update_rgba (reference.py:93-105): integer nearest remap through `data`.
__device__ __forceinline__ uchar4 remap_pixel(const LayerParams& P, const int4& rec, uchar4 px) {
    for (int s = 0; s < P.n_src; s++) {
        bool sel = rec.w == s && rec.z != 0;
        int c = P.chan[s];
        if (sel) {
            int ii = clampi(rec.x, 0, P.h - 1), jj = clampi(rec.y, 0, P.w - 1);
            size_t at = (size_t)ii * P.w + jj;
            if (c == 4) {
                px = __ldg(reinterpret_cast<const uchar4*>(P.pix[s]) + at);
            } else {
                const uint8_t* src = P.pix[s] + at * 3;
                px.x = __ldg(src);
                px.y = __ldg(src + 1);
                px.z = __ldg(src + 2);
            }
        }
        if (c == 3) px.w = sel ? 1 : 0;
    }
    return px;
}

// Layer.render alpha step (layer.py:33) for uint8 rgba: uint8 * float32 -> float32 -> uint8.
__device__ __forceinline__ uint8_t alpha_mask_u8(const float* mask_alpha, int p, uint8_t a) {
    if (!mask_alpha) return a;
    return (uint8_t)(int)__fmul_rn(__ldg(mask_alpha + p), (float)a);
}

// Fused Compositor.render (compositor.py:31-40) for 4 consecutive pixels: packed 12-byte RMW.
__device__ __forceinline__ void composite4(const LayerParams& P, int p0, int count, const uchar4* px) {
    if (!P.rgb) return;
    uint8_t out[12];
    if (count == 4) {
        uint32_t* wp = reinterpret_cast<uint32_t*>(P.rgb + (size_t)p0 * 3);
        bool need_under = false;
#pragma unroll
        for (int k = 0; k < 4; k++) need_under |= (px[k].w == 0);
        if (P.first_layer || !need_under) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                out[3 * k] = (uint8_t)(P.bg >> 16);
                out[3 * k + 1] = (uint8_t)(P.bg >> 8);
                out[3 * k + 2] = (uint8_t)P.bg;
            }
        } else {
            uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
            memcpy(out, &w0, 4);
            memcpy(out + 4, &w1, 4);
            memcpy(out + 8, &w2, 4);
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (px[k].w != 0) {
                out[3 * k] = px[k].x;
                out[3 * k + 1] = px[k].y;
                out[3 * k + 2] = px[k].z;
            }
        uint32_t w0, w1, w2;
        memcpy(&w0, out, 4);
        memcpy(&w1, out + 4, 4);
        memcpy(&w2, out + 8, 4);
        wp[0] = w0;
        wp[1] = w1;
        wp[2] = w2;
    } else {
        for (int k = 0; k < count; k++) {
            uint8_t* o = P.rgb + (size_t)(p0 + k) * 3;
            if (px[k].w != 0) {
                o[0] = px[k].x; o[1] = px[k].y; o[2] = px[k].z;
            } else if (P.first_layer) {
                o[0] = (uint8_t)(P.bg >> 16); o[1] = (uint8_t)(P.bg >> 8); o[2] = (uint8_t)P.bg;
            }
        }
    }
}

// ---- moveref / sum: move (or add) -> reset -> remap -> [render + composite] -----------------
// The four pixels of a thread are processed stage by stage (all flow loads, then all record
// gathers, then all pixmap gathers), so a thread exposes three memory round trips instead of twelve.
template <int KIND>
__global__ void __launch_bounds__(256) k_reference_layer(LayerParams P) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= P.n) return;
    const int count = min(4, P.n - p0);
    int pk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) pk[k] = min(p0 + k, P.n - 1);  // tail lanes recompute the last pixel, stores are guarded

    // stage 1: flow
    float2 f[4];
    if (count == 4 && (reinterpret_cast<uintptr_t>(P.flow) & 15) == 0) {
        float4 a = __ldg(reinterpret_cast<const float4*>(P.flow + p0));
        float4 b = __ldg(reinterpret_cast<const float4*>(P.flow + p0) + 1);
        f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w);
        f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) f[k] = __ldg(P.flow + pk[k]);
    }
    // stage 2: records (own + source) and per-pixel planes
    int4 rec[4];
    uchar4 old_px[4];
    if (KIND == TF_LAYER_MOVEREF) {
        int q[4], off[4];
        int4 rp[4], rq[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            off[k] = __float2int_rn(f[k].y) * P.w + __float2int_rn(f[k].x);  // numpy.round = half-even
            q[k] = off[k] != 0 ? wrap_index(pk[k] + off[k], P.n, pk[k], P.err) : pk[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) rp[k] = __ldg(P.old + pk[k]);
#pragma unroll
        for (int k = 0; k < 4; k++) rq[k] = off[k] != 0 ? __ldg(P.old + q[k]) : rp[k];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // MovementLayer._update_move (movement.py:25-60)
            bool src_alpha = rq[k].z != 0;
            bool ok = off[k] != 0 && (P.mask_src == nullptr || P.mask_src[q[k]] != 0) &&
                      (P.mask_dst == nullptr || P.mask_dst[pk[k]] != 0);
            if (!P.cfg.transparent_pixels_can_move) ok = ok && src_alpha;
            if (!P.cfg.pixels_can_move_to_empty_spot) ok = ok && rp[k].z != 0;
            if (!P.cfg.pixels_can_move_to_filled_spot) ok = ok && rp[k].z == 0;
            rec[k] = ok ? rq[k] : rp[k];
            if (P.cfg.moving_pixels_leave_empty_spot && P.vacated[pk[k]] == P.stamp) rec[k].z = 0;
            if (ok && (!P.cfg.transparent_pixels_can_move || src_alpha)) rec[k].z = 1;
        }
    } else {
        // SumLayer._update_sum (sum.py:9-10): x is added to the ROW index (quirk Q8); in place, so
        // plain loads (not the read-only path)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            rec[k] = P.out[pk[k]];
            rec[k].x += (int)floorf(f[k].x);
            rec[k].y += (int)floorf(f[k].y);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) old_px[k] = P.rgba[pk[k]];
    // stage 3: reset + remap
    uchar4 px[4];
    {
        int y = p0 / P.w, x = p0 - y * P.w;   // one division per thread; the other pixels step from it
#pragma unroll
        for (int k = 0; k < 4; k++) {
            apply_reset(P, pk[k], y, x, rec[k]);
            if (k + 1 < count && ++x == P.w) {
                x = 0;
                y++;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        px[k] = remap_pixel(P, rec[k], old_px[k]);
        if (P.rgb) px[k].w = alpha_mask_u8(P.mask_alpha, pk[k], px[k].w);
    }
    // stage 4: stores
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < count) {
            P.out[p0 + k] = rec[k];
            P.rgba[p0 + k] = px[k];
        }
    composite4(P, p0, count, px);
}

// ---- moveref fast path --------------------------------------------------------------------------
// The configuration every README example and the benchmark use: one source, default movement flags,
// no masks except the reset mask, reset off or random with device draws, single fused layer.  All
// configuration branches are compile-time, the random test runs in fp32 with an exact fallback, and a
// thread's four pixels move as 128-bit words end to end.
struct FastParams {
    const float2* flow;
    int* owner;             // OWNER kernels: the forward scatter's claim plane instead of the flow (consumed and re-zeroed)
    const int4* old;
    int4* out;
    const uint32_t* oldp;   // packed records (PACKED kernels)
    uint32_t* outp;
    uchar4* rgba;
    const float* reset_scale;
    float reset_factor;
    const uint8_t* pix;
    uint8_t* rgb;
    uint32_t bg;
    int h, w, n;
    uint64_t seed, frame;
    int* err;
};

// SUM: the same kernel for the `sum` layer (sum.py:9-10) -- no gather, the record of the pixel itself moves by
// floor(flow) (x to the ROW index, quirk Q8), in place (old == out); the reset, remap and composite stages are shared.
__device__ __forceinline__ int4 rec_unpack(uint32_t r) {
    return make_int4((int)(r & 0x1FFFu), (int)((r >> 13) & 0x1FFFu), (int)((r >> 26) & 1u), (int)(r >> 27));
}
__device__ __forceinline__ uint32_t rec_pack(int4 r) {
    return ((uint32_t)r.x & 0x1FFFu) | (((uint32_t)r.y & 0x1FFFu) << 13) | ((r.z != 0 ? 1u : 0u) << 26) |
           (((uint32_t)r.w & 31u) << 27);
}
// int32 x 4 records <-> packed records; a record that does not fit sets bit 1 of the layer's error word
__global__ void __launch_bounds__(256) k_pack_records(const int4* __restrict__ src, uint32_t* __restrict__ dst, int n,
                                                      int* __restrict__ err) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int4 r = src[p];
    if ((unsigned)r.x > 8191u || (unsigned)r.y > 8191u || (unsigned)r.z > 1u || (unsigned)r.w > 31u) atomicOr(err, 2);
    dst[p] = rec_pack(r);
}
__global__ void __launch_bounds__(256) k_unpack_records(const uint32_t* __restrict__ src, int4* __restrict__ dst, int n) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    dst[p] = rec_unpack(src[p]);
}

// OWNER: the flow of the forward direction is "position of the pixel that claimed me - my position" (core.cu,
// k_post_forward_gather), so the record to fetch is the claimant's: the kernel reads the claim plane of the scatter pass
// (4 bytes per pixel) directly, the flow is never written or read, and the claims are cleared for the next frame here.
template <int RESET, int CHAN, bool SUM = false, bool PACKED = false, bool OWNER = false>
__global__ void __launch_bounds__(256) k_moveref_fast(FastParams P) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= P.n) return;
    float2 f[4];
    int own[4];
    if (OWNER) {
        int4 o = *reinterpret_cast<const int4*>(P.owner + p0);
        own[0] = o.x; own[1] = o.y; own[2] = o.z; own[3] = o.w;
        if ((o.x | o.y | o.z | o.w) != 0) *reinterpret_cast<int4*>(P.owner + p0) = make_int4(0, 0, 0, 0);
    } else {
        float4 a = __ldg(reinterpret_cast<const float4*>(P.flow + p0));
        float4 b = __ldg(reinterpret_cast<const float4*>(P.flow + p0) + 1);
        f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w);
        f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
    }
    int4 rec[4];
    if (SUM) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            rec[k] = P.out[p0 + k];          // in place: plain loads, not the read-only path
            rec[k].x += (int)floorf(f[k].x);
            rec[k].y += (int)floorf(f[k].y);
        }
    } else {
        int q[4];
        bool moved[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (OWNER) {
                moved[k] = own[k] != 0 && own[k] - 1 != p0 + k;   // a claim by the pixel itself is a zero flow
                q[k] = moved[k] ? own[k] - 1 : p0 + k;
            } else {
                int off = __float2int_rn(f[k].y) * P.w + __float2int_rn(f[k].x);
                moved[k] = off != 0;
                q[k] = moved[k] ? wrap_index(p0 + k + off, P.n, p0 + k, P.err) : p0 + k;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) rec[k] = PACKED ? rec_unpack(__ldg(P.oldp + q[k])) : __ldg(P.old + q[k]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // a transparent source does not move (movement.py:34-35): keep the pixel's own record
            if (moved[k] && rec[k].z == 0) rec[k] = PACKED ? rec_unpack(__ldg(P.oldp + p0 + k)) : __ldg(P.old + p0 + k);
            else if (moved[k]) rec[k].z = 1;
        }
    }
    if (RESET == TF_RESET_RANDOM) {
        int y = p0 / P.w, x = p0 - y * P.w;
        float4 thr4 = P.reset_scale ? __ldg(reinterpret_cast<const float4*>(P.reset_scale + p0))
                                    : make_float4(P.reset_factor, P.reset_factor, P.reset_factor, P.reset_factor);
        const float thr[4] = {thr4.x, thr4.y, thr4.z, thr4.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint4 r = philox4x32_7(P.seed, P.frame, (uint32_t)(p0 + k) >> 1);
            uint32_t a = (k & 1) ? r.z : r.x, b = (k & 1) ? r.w : r.y;
            // r = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 lies in [lo, lo + 2^-23): decide in fp32, exact otherwise
            float lo = __uint_as_float(0x3f800000u | (a >> 9)) - 1.0f;
            bool hit;
            if (lo + 1.1920929e-07f <= thr[k]) hit = true;
            else if (lo >= thr[k]) hit = false;
            else hit = ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0) < (double)thr[k];
            if (hit) {
                rec[k].x = y;
                rec[k].y = x;
                rec[k].z = 1;
            }
            if (++x == P.w) {
                x = 0;
                y++;
            }
        }
    }
    uchar4 px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        bool sel = rec[k].w == 0 && rec[k].z != 0;
        if (sel) {
            size_t at = (size_t)clampi(rec[k].x, 0, P.h - 1) * P.w + clampi(rec[k].y, 0, P.w - 1);
            if (CHAN == 4) {
                px[k] = __ldg(reinterpret_cast<const uchar4*>(P.pix) + at);
            } else {
                const uint8_t* src = P.pix + at * 3;
                px[k] = make_uchar4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 1);
            }
        } else {
            px[k] = P.rgba[p0 + k];      // keeps the stale colour (quirk Q10); RGB sources clear the alpha
            if (CHAN == 3) px[k].w = 0;
        }
    }
    if (PACKED) {
        *reinterpret_cast<uint4*>(P.outp + p0) = make_uint4(rec_pack(rec[0]), rec_pack(rec[1]), rec_pack(rec[2]), rec_pack(rec[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) P.out[p0 + k] = rec[k];
    }
    *reinterpret_cast<uint4*>(P.rgba + p0) = make_uint4(*reinterpret_cast<uint32_t*>(&px[0]), *reinterpret_cast<uint32_t*>(&px[1]),
                                                        *reinterpret_cast<uint32_t*>(&px[2]), *reinterpret_cast<uint32_t*>(&px[3]));
    // composite over the constant background, 12 packed bytes
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        c[k] = px[k].w != 0 ? ((uint32_t)px[k].x | ((uint32_t)px[k].y << 8) | ((uint32_t)px[k].z << 16))
                            : (((P.bg >> 16) & 255u) | (P.bg & 0xff00u) | ((P.bg & 255u) << 16));
    uint32_t* wp = reinterpret_cast<uint32_t*>(P.rgb + (size_t)p0 * 3);
    wp[0] = c[0] | (c[1] << 24);
    wp[1] = (c[1] >> 8) | (c[2] << 16);
    wp[2] = (c[2] >> 16) | (c[3] << 8);
}

// ---- static (static.py:13-17) -----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_static_layer(LayerParams P) {
    int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= P.n) return;
    int count = min(4, P.n - p0);
    uchar4 px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k >= count) break;
        int p = p0 + k;
        uchar4 c = P.rgba[p];
        for (int s = 0; s < P.n_src; s++) {
            if (P.intro[s] != nullptr && P.intro[s][p] == 0) continue;
            if (P.chan[s] == 4) {
                c = __ldg(reinterpret_cast<const uchar4*>(P.pix[s]) + p);
            } else {
                const uint8_t* src = P.pix[s] + (size_t)p * 3;
                c.x = __ldg(src); c.y = __ldg(src + 1); c.z = __ldg(src + 2);
            }
        }
        if (P.rgb) c.w = alpha_mask_u8(P.mask_alpha, p, c.w);
        P.rgba[p] = c;
        px[k] = c;
    }
    composite4(P, p0, count, px);
}

// ---- introduction (introduction.py:20-67): records are two int4 ------------------------------
__device__ __forceinline__ uint8_t clip_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

__global__ void __launch_bounds__(256) k_introduction_layer(LayerParams P) {
    int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= P.n) return;
    int count = min(4, P.n - p0);
    uchar4 px[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k >= count) break;
        int p = p0 + k;
        MoveDecision d = decide_move<1>(P, p, 2);
        size_t from = (size_t)(d.target ? d.q : p) * 2;
        int4 a = __ldg(P.old + from), b = __ldg(P.old + from + 1);
        if (P.cfg.moving_pixels_leave_empty_spot && P.vacated[p] == P.stamp) a.w = 0;
        if (d.filled_target) a.w = 1;
        if (P.do_introduce) {
            // mask construction of _update_introduction; the empty-spot / unmoving switches are
            // no-ops in the reference (quirk Q13)
            bool ok = true;
            if (!P.cfg.introduce_pixels_on_filled_spots) ok = ok && a.w == 0;
            if (!P.cfg.introduce_moving_pixels) ok = ok && d.off == 0;
            if (P.cfg.introduce_on_all_filled_spots) ok = ok || a.w != 0;
            bool use_flow = !(P.cfg.introduce_on_all_filled_spots || P.cfg.introduce_on_all_empty_spots);
            if (ok) {
                int src = use_flow ? wrap_index(p + d.off, P.n, p, P.err) : p;
                int sy = src / P.w, sx = src - sy * P.w;
                for (int s = 0; s < P.n_src; s++) {
                    if (P.intro[s] != nullptr && P.intro[s][p] == 0) continue;
                    if (P.chan[s] == 4) {
                        uchar4 c = __ldg(reinterpret_cast<const uchar4*>(P.pix[s]) + src);
                        a = make_int4(c.x, c.y, c.z, c.w);
                    } else {
                        const uint8_t* sp = P.pix[s] + (size_t)src * 3;
                        a = make_int4(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2), 1);
                    }
                    b = make_int4(s, sy, sx, P.frame_no[s]);
                }
            }
        }
        if (P.rgb && P.mask_alpha) {
            // Layer.render on the int32 view: float32 * int32 promotes to float64 in NumPy
            a.w = (int)((double)__ldg(P.mask_alpha + p) * (double)a.w);
        }
        P.out[(size_t)p * 2] = a;
        P.out[(size_t)p * 2 + 1] = b;
        px[k] = make_uchar4(clip_u8(a.x), clip_u8(a.y), clip_u8(a.z), clip_u8(a.w));
    }
    composite4(P, p0, count, px);
}

// ---- unfused Layer.render and Compositor.render ------------------------------------------------
__global__ void __launch_bounds__(256) k_render_rgba(uchar4* rgba, const float* mask_alpha, uchar4* out, int n) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uchar4 c = rgba[p];
    if (mask_alpha) {
        c.w = alpha_mask_u8(mask_alpha, p, c.w);
        rgba[p] = c;
    }
    out[p] = c;
}

__global__ void __launch_bounds__(256) k_render_introduction(int4* data, const float* mask_alpha, uchar4* out, int n) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int4 a = data[(size_t)p * 2];
    if (mask_alpha) {
        a.w = (int)((double)__ldg(mask_alpha + p) * (double)a.w);
        data[(size_t)p * 2] = a;
    }
    out[p] = make_uchar4(clip_u8(a.x), clip_u8(a.y), clip_u8(a.z), clip_u8(a.w));
}

struct CompositeParams {
    const uchar4* layers[TF_MAX_SOURCES * 2];
    int n_layers;
    uint32_t bg;
    uint8_t* rgb;
    int n;
};

__global__ void __launch_bounds__(256) k_composite(CompositeParams C) {
    int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= C.n) return;
    int count = min(4, C.n - p0);
    uint8_t out[12];
    for (int k = 0; k < count; k++) {
        uint8_t r = (uint8_t)(C.bg >> 16), g = (uint8_t)(C.bg >> 8), b = (uint8_t)C.bg;
        for (int l = 0; l < C.n_layers; l++) {
            uchar4 c = __ldg(C.layers[l] + p0 + k);
            if (c.w != 0) {
                r = c.x; g = c.y; b = c.z;
            }
        }
        out[3 * k] = r; out[3 * k + 1] = g; out[3 * k + 2] = b;
    }
    if (count == 4) {
        uint32_t w0, w1, w2;
        memcpy(&w0, out, 4); memcpy(&w1, out + 4, 4); memcpy(&w2, out + 8, 4);
        uint32_t* wp = reinterpret_cast<uint32_t*>(C.rgb + (size_t)p0 * 3);
        wp[0] = w0; wp[1] = w1; wp[2] = w2;
    } else {
        for (int k = 0; k < count * 3; k++) C.rgb[(size_t)p0 * 3 + k] = out[k];
    }
}

__global__ void __launch_bounds__(256) k_init_reference_data(int4* data, const int8_t* base_src, int h, int w,
                                                             int keep_existing) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    int y = p / w, x = p - y * w;
    int4 rec = keep_existing ? data[p] : make_int4(y, x, 1, 0);
    if (base_src && base_src[p] >= 0) rec.w = base_src[p];
    data[p] = rec;
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* p, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256) k_base_source(int8_t* base_src, LayerParams P) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    int v = -1;
    for (int s = 0; s < P.n_src; s++)
        if (P.intro[s] == nullptr || P.intro[s][p] != 0) v = s;
    base_src[p] = (int8_t)v;
}

// ------------------------------------------------------------------------------------------------
static int copy_plane(void** dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!src) {
        if (*dst) cudaFree(*dst);
        *dst = nullptr;
        return TF_OK;
    }
    if (!*dst) TF_CUDA(cudaMalloc(dst, bytes));
    TF_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    return TF_OK;
}

extern "C" int tf_layer_create(tf_layer** out, int height, int width, const tf_layer_config* cfg) {
    TF_REQUIRE(out && cfg, TF_ERR_INVALID_ARG, "tf_layer_create: null argument");
    TF_REQUIRE(height > 0 && width > 0 && (size_t)height * width < (1u << 30), TF_ERR_SHAPE,
               "tf_layer_create: bad shape %dx%d", height, width);
    TF_REQUIRE(cfg->kind >= TF_LAYER_MOVEREF && cfg->kind <= TF_LAYER_INTRODUCTION, TF_ERR_INVALID_ARG,
               "Unknown layer kind %d", cfg->kind);
    TF_REQUIRE(cfg->reset_mode >= TF_RESET_OFF && cfg->reset_mode <= TF_RESET_LINEAR, TF_ERR_INVALID_ARG,
               "Unknown reset mode %d", cfg->reset_mode);
    if (int e = require_sm100()) return e;
    tf_layer* l = new (std::nothrow) tf_layer();
    TF_REQUIRE(l, TF_ERR_CUDA, "out of host memory");
    memset(l, 0, sizeof(*l));
    l->cfg = *cfg;
    l->h = height;
    l->w = width;
    l->depth = cfg->kind == TF_LAYER_INTRODUCTION ? 8 : 4;
    size_t n = (size_t)height * width;
    int blocks = ceil_div((int)n, 256);
    auto bail = [&](int e) { tf_layer_destroy(l); return e; };
    if (cudaMalloc(&l->err, sizeof(int)) != cudaSuccess) return bail(fail(TF_ERR_CUDA, "cudaMalloc failed"));
    cudaMemset(l->err, 0, sizeof(int));
    if (cfg->kind != TF_LAYER_STATIC) {
        for (int i = 0; i < 2; i++)
            if (cudaMalloc(&l->data[i], n * l->depth * sizeof(int)) != cudaSuccess)
                return bail(fail(TF_ERR_CUDA, "cudaMalloc(data) failed for %dx%d", height, width));
        if (cfg->kind == TF_LAYER_INTRODUCTION) {
            cudaMemset(l->data[0], 0, n * 32);
        } else {
            k_init_reference_data<<<blocks, 256>>>(l->data[0], nullptr, height, width, 0);
            TF_LAUNCHED();
        }
        if (cfg->kind != TF_LAYER_SUM && cfg->moving_pixels_leave_empty_spot) {
            if (cudaMalloc(&l->vacated, n * 4) != cudaSuccess) return bail(fail(TF_ERR_CUDA, "cudaMalloc failed"));
            cudaMemset(l->vacated, 0, n * 4);
        }
        l->int4_valid = true;
        const char* pk = getenv("TFB200_PACKED_RECORDS");
        if (cfg->kind == TF_LAYER_MOVEREF && height <= 8192 && width <= 8192 && !(pk && pk[0] == '0')) {
            for (int i = 0; i < 2; i++)
                if (cudaMalloc(&l->pdata[i], n * sizeof(uint32_t)) != cudaSuccess)
                    return bail(fail(TF_ERR_CUDA, "cudaMalloc(packed data) failed for %dx%d", height, width));
            l->packed_ok = true;
        }
    }
    if (cfg->kind != TF_LAYER_INTRODUCTION) {
        if (cudaMalloc(&l->rgba, n * 4) != cudaSuccess) return bail(fail(TF_ERR_CUDA, "cudaMalloc(rgba) failed"));
        // Layer.__init__: zeros; StaticLayer presets alpha = 1 (static.py:11)
        k_fill_u32<<<blocks, 256>>>(reinterpret_cast<uint32_t*>(l->rgba),
                                    cfg->kind == TF_LAYER_STATIC ? 0x01000000u : 0u, n);
        TF_LAUNCHED();
    }
    TF_CUDA(cudaDeviceSynchronize());
    *out = l;
    return TF_OK;
}

// the int32 x 4 array / the packed array is brought up to date from the other one when a consumer needs it
static int ensure_int4(tf_layer* l, cudaStream_t st) {
    if (l->int4_valid || !l->data[l->cur]) return TF_OK;
    int n = l->h * l->w;
    k_unpack_records<<<ceil_div(n, 256), 256, 0, st>>>(l->pdata[l->pcur], l->data[l->cur], n);
    TF_LAUNCHED();
    l->int4_valid = true;
    return TF_OK;
}
static int ensure_packed(tf_layer* l, cudaStream_t st) {
    if (l->packed_valid) return TF_OK;
    int n = l->h * l->w;
    k_pack_records<<<ceil_div(n, 256), 256, 0, st>>>(l->data[l->cur], l->pdata[l->pcur], n, l->err);
    TF_LAUNCHED();
    l->packed_valid = true;
    return TF_OK;
}

extern "C" int tf_layer_destroy(tf_layer* l) {
    if (!l) return TF_OK;
    cudaFree(l->data[0]); cudaFree(l->data[1]); cudaFree(l->rgba);
    cudaFree(l->pdata[0]); cudaFree(l->pdata[1]);
    cudaFree(l->mask_src); cudaFree(l->mask_dst); cudaFree(l->mask_alpha); cudaFree(l->reset_scale);
    cudaFree(l->base_src); cudaFree(l->vacated); cudaFree(l->err);
    for (int s = 0; s < TF_MAX_SOURCES; s++) cudaFree(l->intro[s]);
    delete l;
    return TF_OK;
}

extern "C" int tf_layer_set_masks(tf_layer* l, const uint8_t* mask_src, const uint8_t* mask_dst,
                                  const float* mask_alpha, const float* reset_scale, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_set_masks: null layer");
    size_t n = (size_t)l->h * l->w;
    cudaStream_t st = as_stream(stream);
    if (int e = copy_plane((void**)&l->mask_src, mask_src, n, st)) return e;
    if (int e = copy_plane((void**)&l->mask_dst, mask_dst, n, st)) return e;
    if (int e = copy_plane((void**)&l->mask_alpha, mask_alpha, n * 4, st)) return e;
    if (int e = copy_plane((void**)&l->reset_scale, reset_scale, n * 4, st)) return e;
    return TF_OK;
}

static void fill_common(const tf_layer* l, LayerParams& P) {
    memset(&P, 0, sizeof(P));
    P.h = l->h; P.w = l->w; P.n = l->h * l->w;
    P.cfg = l->cfg;
    P.mask_src = l->mask_src; P.mask_dst = l->mask_dst; P.mask_alpha = l->mask_alpha;
    P.reset_scale = l->reset_scale;
    P.base_src = l->base_src;
    P.vacated = l->vacated;
    P.err = l->err;
    P.rgba = l->rgba;
    P.n_src = l->n_sources;
    for (int s = 0; s < l->n_sources; s++) P.intro[s] = l->intro[s];
}

extern "C" int tf_layer_set_sources(tf_layer* l, int n_sources, const uint8_t* const* intro_masks, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_set_sources: null layer");
    TF_REQUIRE(n_sources >= 0 && n_sources <= TF_MAX_SOURCES, TF_ERR_INVALID_ARG,
               "tf_layer_set_sources: %d sources (max %d)", n_sources, TF_MAX_SOURCES);
    size_t n = (size_t)l->h * l->w;
    cudaStream_t st = as_stream(stream);
    for (int s = 0; s < TF_MAX_SOURCES; s++) {
        const uint8_t* m = (s < n_sources && intro_masks) ? intro_masks[s] : nullptr;
        if (int e = copy_plane((void**)&l->intro[s], m, n, st)) return e;
    }
    l->n_sources = n_sources;
    if (l->cfg.kind == TF_LAYER_MOVEREF || l->cfg.kind == TF_LAYER_SUM) {
        // ReferenceLayer._set_base_source_indices (reference.py:46-52)
        if (n_sources == 0) {
            cudaFree(l->base_src);
            l->base_src = nullptr;
            return TF_OK;
        }
        if (!l->base_src) TF_CUDA(cudaMalloc(&l->base_src, n));
        LayerParams P;
        fill_common(l, P);
        int blocks = ceil_div((int)n, 256);
        k_base_source<<<blocks, 256, 0, st>>>(l->base_src, P);
        TF_LAUNCHED();
        if (int e = ensure_int4(l, st)) return e;
        k_init_reference_data<<<blocks, 256, 0, st>>>(l->data[l->cur], l->base_src, l->h, l->w, 1);
        TF_LAUNCHED();
        l->packed_valid = false;
    }
    return TF_OK;
}

// does the single-source move-reference layer run on the fast kernel with packed records?  (shared by tf_layer_update
// and tf_layer_takes_claims; `aligned` = the caller's per-call pointers are 16-byte aligned)
static bool moveref_fast_ok(const tf_layer* l, int n_pixmaps, bool random_given, bool fused_first, bool aligned) {
    const tf_layer_config& c = l->cfg;
    const bool leave = c.moving_pixels_leave_empty_spot;
    const int n = l->h * l->w;
    return n_pixmaps == 1 && !leave && !c.transparent_pixels_can_move && c.pixels_can_move_to_empty_spot &&
           c.pixels_can_move_to_filled_spot && !l->mask_src && !l->mask_dst && !l->mask_alpha &&
           (c.reset_mode == TF_RESET_OFF || (c.reset_mode == TF_RESET_RANDOM && !random_given && !c.reset_source)) &&
           fused_first && (n & 3) == 0 && aligned && ((uintptr_t)l->reset_scale & 15) == 0;
}

extern "C" int tf_layer_takes_claims(const tf_layer* l, int n_pixmaps) {
    if (!l || l->cfg.kind != TF_LAYER_MOVEREF) return 0;
    return moveref_fast_ok(l, n_pixmaps, false, true, true) && l->packed_ok ? 1 : 0;
}

static int layer_update(tf_layer* l, const float* flow, int* owner, const tf_pixmap* pixmaps, int n_pixmaps,
                        const double* random, uint64_t rng_seed, uint8_t* rgb_inout, int first_layer,
                        uint32_t background_rgb, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_update: null layer");
    TF_REQUIRE(flow || owner || l->cfg.kind == TF_LAYER_STATIC, TF_ERR_INVALID_ARG, "tf_layer_update: null flow");
    TF_REQUIRE(!owner || (l->cfg.kind == TF_LAYER_MOVEREF && l->packed_ok &&
                          moveref_fast_ok(l, n_pixmaps, random != nullptr, rgb_inout && first_layer, ((uintptr_t)owner & 15) == 0)),
               TF_ERR_INVALID_ARG, "tf_layer_update_claims: this layer configuration needs the flow (see tf_layer_takes_claims)");
    TF_REQUIRE(n_pixmaps == l->n_sources, TF_ERR_INVALID_ARG, "tf_layer_update: %d pixmaps for %d sources", n_pixmaps,
               l->n_sources);
    TF_REQUIRE(n_pixmaps == 0 || pixmaps, TF_ERR_INVALID_ARG, "tf_layer_update: null pixmap array");
    TF_REQUIRE(((uintptr_t)rgb_inout & 3) == 0, TF_ERR_INVALID_ARG, "tf_layer_update: rgb buffer must be 4-byte aligned");
    cudaStream_t st = as_stream(stream);
    LayerParams P;
    fill_common(l, P);
    for (int s = 0; s < n_pixmaps; s++) {
        TF_REQUIRE(pixmaps[s].pixels && (pixmaps[s].channels == 3 || pixmaps[s].channels == 4), TF_ERR_SHAPE,
                   "tf_layer_update: pixmap %d must be (H, W, 3|4) uint8", s);
        TF_REQUIRE(pixmaps[s].channels == 3 || ((uintptr_t)pixmaps[s].pixels & 3) == 0, TF_ERR_INVALID_ARG,
                   "tf_layer_update: RGBA pixmap must be 4-byte aligned");
        P.pix[s] = pixmaps[s].pixels;
        P.chan[s] = pixmaps[s].channels;
        P.frame_no[s] = pixmaps[s].frame_number;
    }
    P.flow = reinterpret_cast<const float2*>(flow);
    P.random = random;
    P.seed = rng_seed;
    P.frame = l->frames;
    P.stamp = (uint32_t)(l->frames + 1);
    P.rgb = rgb_inout;
    P.first_layer = first_layer;
    P.bg = background_rgb;
    int groups = ceil_div(P.n, 4);
    int blocks4 = ceil_div(groups, 256), blocks1 = ceil_div(P.n, 256);
    int kind = l->cfg.kind;
    if (kind == TF_LAYER_STATIC) {
        k_static_layer<<<blocks4, 256, 0, st>>>(P);
        TF_LAUNCHED();
    } else {
        P.old = l->data[l->cur];
        // the sum layer has no gather: update in place; the others ping-pong
        P.out = kind == TF_LAYER_SUM ? l->data[l->cur] : l->data[l->cur ^ 1];
        bool leave = kind != TF_LAYER_SUM && l->cfg.moving_pixels_leave_empty_spot;
        if (kind == TF_LAYER_MOVEREF) {
            if (leave) {
                if (int e = ensure_int4(l, st)) return e;
                k_mark_vacated<0><<<blocks1, 256, 0, st>>>(P, 1);
                TF_LAUNCHED();
            }
            const tf_layer_config& c = l->cfg;
            bool fast = moveref_fast_ok(l, n_pixmaps, random != nullptr, rgb_inout && first_layer,
                                        (((uintptr_t)flow | (uintptr_t)owner) & 15) == 0);
            const bool packed = fast && l->packed_ok;
            if (packed) {
                if (int e = ensure_packed(l, st)) return e;
            } else {
                if (int e = ensure_int4(l, st)) return e;
            }
            {
                ScopedKernelTimer timer(TFK_COMPOSITOR_LAYER, st);
                if (fast) {
                    FastParams F;
                    F.flow = P.flow; F.old = P.old; F.out = P.out; F.rgba = P.rgba; F.reset_scale = l->reset_scale;
                    F.oldp = l->pdata[l->pcur]; F.outp = l->pdata[l->pcur ^ 1];
                    F.reset_factor = c.reset_random_factor; F.pix = P.pix[0]; F.rgb = rgb_inout; F.bg = background_rgb;
                    F.h = P.h; F.w = P.w; F.n = P.n; F.seed = P.seed; F.frame = P.frame; F.err = P.err;
                    bool rnd = c.reset_mode == TF_RESET_RANDOM;
                    F.owner = owner;
                    if (packed && owner) {
                        if (P.chan[0] == 4) {
                            if (rnd) k_moveref_fast<TF_RESET_RANDOM, 4, false, true, true><<<blocks4, 256, 0, st>>>(F);
                            else k_moveref_fast<TF_RESET_OFF, 4, false, true, true><<<blocks4, 256, 0, st>>>(F);
                        } else {
                            if (rnd) k_moveref_fast<TF_RESET_RANDOM, 3, false, true, true><<<blocks4, 256, 0, st>>>(F);
                            else k_moveref_fast<TF_RESET_OFF, 3, false, true, true><<<blocks4, 256, 0, st>>>(F);
                        }
                    } else if (packed) {
                        if (P.chan[0] == 4) {
                            if (rnd) k_moveref_fast<TF_RESET_RANDOM, 4, false, true><<<blocks4, 256, 0, st>>>(F);
                            else k_moveref_fast<TF_RESET_OFF, 4, false, true><<<blocks4, 256, 0, st>>>(F);
                        } else {
                            if (rnd) k_moveref_fast<TF_RESET_RANDOM, 3, false, true><<<blocks4, 256, 0, st>>>(F);
                            else k_moveref_fast<TF_RESET_OFF, 3, false, true><<<blocks4, 256, 0, st>>>(F);
                        }
                    } else if (P.chan[0] == 4) {
                        if (rnd) k_moveref_fast<TF_RESET_RANDOM, 4><<<blocks4, 256, 0, st>>>(F);
                        else k_moveref_fast<TF_RESET_OFF, 4><<<blocks4, 256, 0, st>>>(F);
                    } else {
                        if (rnd) k_moveref_fast<TF_RESET_RANDOM, 3><<<blocks4, 256, 0, st>>>(F);
                        else k_moveref_fast<TF_RESET_OFF, 3><<<blocks4, 256, 0, st>>>(F);
                    }
                } else {
                    k_reference_layer<TF_LAYER_MOVEREF><<<blocks4, 256, 0, st>>>(P);
                }
            }
            TF_LAUNCHED();
            if (packed) {
                l->pcur ^= 1;
                l->int4_valid = false;
            } else {
                l->cur ^= 1;
                l->packed_valid = false;
            }
        } else if (kind == TF_LAYER_SUM) {
            const tf_layer_config& c = l->cfg;
            // single source, no alpha mask, reset off or random with device draws, fused single layer: the fast kernel
            bool fast = n_pixmaps == 1 && !l->mask_alpha &&
                        (c.reset_mode == TF_RESET_OFF || (c.reset_mode == TF_RESET_RANDOM && !random && !c.reset_source)) &&
                        rgb_inout && first_layer && (P.n & 3) == 0 && ((uintptr_t)flow & 15) == 0 &&
                        ((uintptr_t)l->reset_scale & 15) == 0;
            {
                ScopedKernelTimer timer(TFK_COMPOSITOR_LAYER, st);
                if (fast) {
                    FastParams F;
                    F.flow = P.flow; F.old = P.old; F.out = P.out; F.rgba = P.rgba; F.reset_scale = l->reset_scale;
                    F.reset_factor = c.reset_random_factor; F.pix = P.pix[0]; F.rgb = rgb_inout; F.bg = background_rgb;
                    F.h = P.h; F.w = P.w; F.n = P.n; F.seed = P.seed; F.frame = P.frame; F.err = P.err;
                    bool rnd = c.reset_mode == TF_RESET_RANDOM;
                    if (P.chan[0] == 4) {
                        if (rnd) k_moveref_fast<TF_RESET_RANDOM, 4, true><<<blocks4, 256, 0, st>>>(F);
                        else k_moveref_fast<TF_RESET_OFF, 4, true><<<blocks4, 256, 0, st>>>(F);
                    } else {
                        if (rnd) k_moveref_fast<TF_RESET_RANDOM, 3, true><<<blocks4, 256, 0, st>>>(F);
                        else k_moveref_fast<TF_RESET_OFF, 3, true><<<blocks4, 256, 0, st>>>(F);
                    }
                } else {
                    k_reference_layer<TF_LAYER_SUM><<<blocks4, 256, 0, st>>>(P);
                }
            }
            TF_LAUNCHED();
        } else {
            P.do_introduce = !(l->cfg.introduce_once && l->introduced_once);
            if (leave) {
                k_mark_vacated<1><<<blocks1, 256, 0, st>>>(P, 2);
                TF_LAUNCHED();
            }
            k_introduction_layer<<<blocks4, 256, 0, st>>>(P);
            TF_LAUNCHED();
            l->cur ^= 1;
            l->introduced_once = 1;
        }
    }
    l->frames++;
    return TF_OK;
}

extern "C" int tf_layer_update(tf_layer* l, const float* flow, const tf_pixmap* pixmaps, int n_pixmaps,
                               const double* random, uint64_t rng_seed, uint8_t* rgb_inout, int first_layer,
                               uint32_t background_rgb, void* stream) {
    return layer_update(l, flow, nullptr, pixmaps, n_pixmaps, random, rng_seed, rgb_inout, first_layer, background_rgb,
                        stream);
}

extern "C" int tf_layer_update_claims(tf_layer* l, int32_t* claims, const tf_pixmap* pixmaps, int n_pixmaps,
                                      uint64_t rng_seed, uint8_t* rgb_inout, uint32_t background_rgb, void* stream) {
    TF_REQUIRE(claims, TF_ERR_INVALID_ARG, "tf_layer_update_claims: null claim plane");
    return layer_update(l, nullptr, claims, pixmaps, n_pixmaps, nullptr, rng_seed, rgb_inout, 1, background_rgb, stream);
}

extern "C" int tf_layer_render(tf_layer* l, uint8_t* rgba_out, void* stream) {
    TF_REQUIRE(l && rgba_out, TF_ERR_INVALID_ARG, "tf_layer_render: null argument");
    TF_REQUIRE(((uintptr_t)rgba_out & 3) == 0, TF_ERR_INVALID_ARG, "tf_layer_render: output must be 4-byte aligned");
    int n = l->h * l->w;
    cudaStream_t st = as_stream(stream);
    if (l->cfg.kind == TF_LAYER_INTRODUCTION) {
        k_render_introduction<<<ceil_div(n, 256), 256, 0, st>>>(l->data[l->cur], l->mask_alpha,
                                                                reinterpret_cast<uchar4*>(rgba_out), n);
    } else {
        k_render_rgba<<<ceil_div(n, 256), 256, 0, st>>>(l->rgba, l->mask_alpha, reinterpret_cast<uchar4*>(rgba_out), n);
    }
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_composite(const uint8_t* const* layer_rgba, int n_layers, uint32_t background_rgb, uint8_t* rgb_out,
                            int height, int width, void* stream) {
    TF_REQUIRE(rgb_out && (n_layers == 0 || layer_rgba), TF_ERR_INVALID_ARG, "tf_composite: null argument");
    TF_REQUIRE(n_layers >= 0 && n_layers <= TF_MAX_SOURCES * 2, TF_ERR_INVALID_ARG, "tf_composite: %d layers (max %d)",
               n_layers, TF_MAX_SOURCES * 2);
    TF_REQUIRE(height > 0 && width > 0, TF_ERR_SHAPE, "tf_composite: bad shape %dx%d", height, width);
    TF_REQUIRE(((uintptr_t)rgb_out & 3) == 0, TF_ERR_INVALID_ARG, "tf_composite: output must be 4-byte aligned");
    if (int e = require_sm100()) return e;
    CompositeParams C;
    memset(&C, 0, sizeof(C));
    for (int i = 0; i < n_layers; i++) C.layers[i] = reinterpret_cast<const uchar4*>(layer_rgba[i]);
    C.n_layers = n_layers;
    C.bg = background_rgb;
    C.rgb = rgb_out;
    C.n = height * width;
    k_composite<<<ceil_div(ceil_div(C.n, 4), 256), 256, 0, as_stream(stream)>>>(C);
    TF_LAUNCHED();
    return TF_OK;
}

extern "C" int tf_layer_depth(const tf_layer* l) { return l ? l->depth : 0; }

extern "C" int tf_layer_get_state(tf_layer* l, int32_t* data, uint8_t* rgba, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_get_state: null layer");
    size_t n = (size_t)l->h * l->w;
    cudaStream_t st = as_stream(stream);
    if (data && l->data[l->cur]) {
        if (int e = ensure_int4(l, st)) return e;
        TF_CUDA(cudaMemcpyAsync(data, l->data[l->cur], n * l->depth * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (rgba && l->rgba) TF_CUDA(cudaMemcpyAsync(rgba, l->rgba, n * 4, cudaMemcpyDeviceToDevice, st));
    return TF_OK;
}

extern "C" int tf_layer_set_state(tf_layer* l, const int32_t* data, const uint8_t* rgba, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_set_state: null layer");
    size_t n = (size_t)l->h * l->w;
    cudaStream_t st = as_stream(stream);
    if (data && l->data[l->cur]) {
        TF_CUDA(cudaMemcpyAsync(l->data[l->cur], data, n * l->depth * 4, cudaMemcpyDeviceToDevice, st));
        l->int4_valid = true;
        l->packed_valid = false;
    }
    if (rgba && l->rgba) TF_CUDA(cudaMemcpyAsync(l->rgba, rgba, n * 4, cudaMemcpyDeviceToDevice, st));
    return TF_OK;
}

extern "C" int tf_layer_poll_error(tf_layer* l, void* stream) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "tf_layer_poll_error: null layer");
    int v = 0;
    TF_CUDA(cudaMemcpyAsync(&v, l->err, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
    TF_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (v & 2) return fail(TF_ERR_CUDA, "internal: a layer record does not fit the packed 13/13/1/5-bit form");
    if (v) return fail(TF_ERR_INDEX, "index out of bounds: a flow vector points outside the frame "
                                     "(post-process the flow or clip it first)");
    return TF_OK;
}

extern "C" int tf_layer_get_counters(const tf_layer* l, uint64_t* frames, int* introduced_once) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "null layer");
    if (frames) *frames = l->frames;
    if (introduced_once) *introduced_once = l->introduced_once;
    return TF_OK;
}

extern "C" int tf_layer_set_counters(tf_layer* l, uint64_t frames, int introduced_once) {
    TF_REQUIRE(l, TF_ERR_INVALID_ARG, "null layer");
    l->frames = frames;
    l->introduced_once = introduced_once;
    return TF_OK;
}
