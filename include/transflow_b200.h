/*
 * transflow_b200 -- C ABI of the B200-native (sm_100a) transflow hot path.
 *
 * Dense optical flow per frame pair -> per-pixel flow accumulation -> pixmap remap.
 * Every entry point is what a binding of the reference (ychalier/transflow, pure Python)
 * would call in place of its cv2 / NumPy code; the reference interface each one replaces is
 * cited as  transflow/<file>:<line>.
 *
 * Conventions
 *   - every function returns 0 (TF_OK) or a negative tf_status; tf_last_error() gives text;
 *   - all array arguments are DEVICE pointers to C-contiguous buffers owned by the caller
 *     unless the name ends in _host; the library owns only per-handle scratch/state;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - handles are not thread-safe: one handle <-> one stream at a time;
 *   - there is no CPU fallback: a device that is not sm_100 yields TF_ERR_UNSUPPORTED_ARCH.
 *
 * Layouts (transflow/types.py:8-14, compositor/layers/data.py:8-12):
 *   gray   uint8  (H, W)          bgr/rgb uint8 (H, W, 3)     rgba uint8 (H, W, 4)
 *   flow   float32 (H, W, 2)      [...,0] = dx (columns), [...,1] = dy (rows)
 *   data   int32  (H, W, 4) = (i, j, alpha, source)           (reference layers)
 *          int32  (H, W, 8) = (r, g, b, alpha, source, i, j, frame)   (introduction layer)
 */
#ifndef TRANSFLOW_B200_H
#define TRANSFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TF_API __attribute__((visibility("default")))
#else
#define TF_API
#endif

typedef enum {
    TF_OK = 0,
    TF_ERR_INVALID_ARG = -1,      /* -> ValueError   */
    TF_ERR_CUDA = -2,             /* -> RuntimeError */
    TF_ERR_UNSUPPORTED_ARCH = -3, /* -> RuntimeError */
    TF_ERR_SHAPE = -4,            /* -> ValueError   */
    TF_ERR_INDEX = -5             /* -> IndexError (gather index outside the frame) */
} tf_status;

TF_API int tf_version(void);
TF_API const char* tf_last_error(void);
/* 0 when `device` is an sm_100 part; fills sm count when sm_count != NULL. */
TF_API int tf_device_check(int device, int* sm_count);
/* Number of kernels this library has launched in the calling process (bench.py gpu_launches). */
TF_API uint64_t tf_launch_count(void);

/* Optional device timing of the dominant kernels (CUDA events on the launching stream), used by
 * bench.py for the roofline figure.  Tags: 0 fused Farneback iteration (finest level), 1/2/3
 * unfused update-matrices / vertical sums / horizontal sums+solve (finest level), 4 polynomial
 * expansion (finest level), 5 compositor layer kernel, 6 forward post-process, 7 Horn-Schunck
 * sweep, 8 Lucas-Kanade tracker (level 0).  tf_timer_read synchronises the device. */
TF_API int tf_timer_enable(int on);
TF_API int tf_timer_read(int tag, double* total_ms, uint64_t* launches);

/* ---- frame prep: cv2.cvtColor(BGR2GRAY) at transflow/flow/sources/cv.py:465 ------------- */
TF_API int tf_gray_from_bgr(const uint8_t* bgr, uint8_t* gray, int height, int width, void* stream);
/* cv2.resize(frame, (width, height), interpolation=cv2.INTER_NEAREST) of a BGR u8 frame
 * (transflow/flow/sources/cv.py:464, :455-457), bit-exact: source index min(floor(x / (dw / sw)), sw - 1). */
TF_API int tf_resize_nearest_bgr(const uint8_t* src, int src_height, int src_width, uint8_t* dst, int height, int width,
                                 void* stream);

/* ---- Farneback: cv2.calcOpticalFlowFarneback at transflow/flow/sources/cv.py:477-490 ----- */
typedef struct tf_farneback tf_farneback;
/* Parameters = CvFlowConfig.fb_* (transflow/flow/sources/cv.py:275-281).  Only flags == 0 is
 * supported (the reference default).  r_fp16 != 0 stores the polynomial-expansion tensors in
 * half precision (compute stays fp32). */
TF_API int tf_farneback_create(tf_farneback** out, int height, int width, double pyr_scale, int levels,
                        int winsize, int iterations, int poly_n, double poly_sigma, int flags,
                        int r_fp16);
TF_API int tf_farneback_destroy(tf_farneback* h);
/* Gaussian pyramid + polynomial expansion of one frame into slot 0, 1 or 2 (cached across pairs,
 * since a frame is the right image of one pair and the left image of the next; slot 2 is allocated
 * on first use and only needed by tf_farneback_step_lane). */
TF_API int tf_farneback_prepare(tf_farneback* h, int slot, const uint8_t* gray, void* stream);
/* Coarse-to-fine displacement solve between two prepared slots -> flow (H, W, 2).
 * variant: 9-16 = experimental staged kernels, all bit-identical to 8 and slower on B200:
 * 9 / 10 = R1 operand staged in shared memory by per-row bulk copies, 64- / 32-column strips (fb_stage.cuh);
 * 11 = 10 with the half's own R0 / flow rows staged too; 12 / 13 / 14 = every operand by 2-D bulk tensor copies,
 * box margin 3 / 2 / 1 px (fb_tma.cuh), 15 / 16 = 14 with two sets of operand buffers (192 / 384 compute
 * threads); 8 = default: fused half-buffer iteration kernel (fb_half.cuh) on pyramid levels of >= 0.4 Mpx, fused
 * rolling-tile kernel (fb_tile.cuh) on the smaller ones; 4-7 = half-buffer kernel everywhere (scalar or 128-bit
 * shared accesses, 4 or 3 CTAs per SM); 3 = rolling-tile kernel everywhere; 1 = unfused reference kernels
 * (M and double vertical sums materialised in HBM); 0 / 2 = fused column-streaming kernel with
 * float / double sums in shared memory (kept for comparison); 17-24 = configurations of the half-buffer kernel and
 * the TMA-fed ring kernels (fb_ring.cuh); 25 = packed half-buffer kernel (channel pairs in shared memory, FADD2 window
 * sums, fb_pack.cuh), 27 / 31 = the same with pairs of rows sharing their middle tap row (31: one wait per pair, 80
 * registers, 3 CTAs per SM) -- all bit-identical to 8 and not faster;
 * 28 / 29 / 30 = timing experiments (phase A only / phases B + C only / B + C with 14 instead of 23 phase-C reads of
 * variant 25: NOT a flow), winsize 15 only.
 * If clip != 0 the final clip of FlowSource.post_process (source.py:361-362) is fused into
 * the last store. */
TF_API int tf_farneback_solve(tf_farneback* h, int slot_left, int slot_right, float* flow, int variant,
                       int clip, void* stream);
/* prepare(new_slot, gray) + solve(slot_left, slot_right) for a streaming source, overlapped: the new
 * frame's levels are built coarse -> fine on an internal auxiliary stream while `stream` already
 * solves the coarser levels (the small pyramid levels cannot fill 148 SMs on their own).  The other
 * slot must hold the previously prepared frame.  Everything is ordered after prior work on `stream`
 * and complete, from `stream`'s point of view, when the call's last kernel finishes. */
TF_API int tf_farneback_step(tf_farneback* h, int new_slot, const uint8_t* gray, int slot_left, int slot_right,
                             float* flow, int variant, int clip, void* stream);
/* tf_farneback_step on one of two solve LANES (0 or 1; lane 1's per-level flow buffers are allocated on
 * first use).  Pair t's flow does not depend on pair t-1's (fb_flags == 0), so a streaming source can keep two
 * pairs in flight: frame t lives in slot t % 3, pair (t, t+1) runs on lane t % 2 with its own `stream`, and
 * every frame is still prepared exactly once.  The second pair fills the SMs the first leaves idle (the tails
 * of its launches and the small pyramid levels): +9 % pairs/s at 4K on B200.  The handle orders, with events,
 * (a) each solve after the expansion of BOTH its frames, whichever call built them, and (b) the overwrite of
 * new_slot after every solve that read its previous frame, on either lane.  tf_farneback_step == lane 0. */
TF_API int tf_farneback_step_lane(tf_farneback* h, int lane, int new_slot, const uint8_t* gray, int slot_left,
                                  int slot_right, float* flow, int variant, int clip, void* stream);
/* prepare(0, left) + prepare(1, right) + solve(0, 1). */
TF_API int tf_farneback_run(tf_farneback* h, const uint8_t* left, const uint8_t* right, float* flow,
                     int variant, void* stream);
TF_API int tf_farneback_num_levels(const tf_farneback* h);
TF_API int tf_farneback_level_size(const tf_farneback* h, int level_index, int* width, int* height);
/* Allocate frame slots [0, slots) and solve lanes [0, lanes) now (at most 3 / 2).  tf_farneback_create allocates two
 * slots and one lane; a caller that keeps two pairs in flight (tf_farneback_step_lane) reserves 3 / 2 up front so that
 * no cudaMalloc -- a device-wide synchronisation -- happens inside the frame loop. */
TF_API int tf_farneback_reserve(tf_farneback* h, int slots, int lanes);
/* Debug mode: prepare() also writes the finest level's blurred image, which otherwise only exists inside the
 * fused polynomial-expansion kernel (needed by tf_farneback_debug_read(what = 0) on that level). */
TF_API int tf_farneback_set_debug(tf_farneback* h, int on);
/* Test hooks: copy stage outputs (as float32) to a device buffer.  what: 0 = pyramid image
 * (h, w); 1 = polynomial expansion (5, h, w) of `slot`; 2 = flow (h, w, 2) at the end of the
 * level during the last solve.  level_index 0 = coarsest. */
TF_API int tf_farneback_debug_read(tf_farneback* h, int slot, int level_index, int what, float* out,
                            void* stream);
/* Tuning knob for experiments (process-wide): key 0 = rows per CTA of the variant 4-7 kernels (0 = heuristic);
 * key 1 = 1 selects the separate horizontal / vertical pyramid blur passes instead of the fused kernel;
 * key 2 = smallest level (pixels) the key-0 override applies to; key 3 / 4 = rows per CTA / smallest level of the ring
 * kernels. */
TF_API int tf_farneback_tune(int key, int value);
/* Algorithmic bytes moved per solved pair (SURVEY.md 8d model), for roofline reporting. */
TF_API double tf_farneback_algorithmic_bytes(const tf_farneback* h, int reuse_r);

/* ---- Horn-Schunck: transflow/flow/methods/horn_schunck.py:9-45 ---------------------------- */
typedef struct tf_horn_schunck tf_horn_schunck;
TF_API int tf_hs_create(tf_horn_schunck** out, int height, int width);
TF_API int tf_hs_destroy(tf_horn_schunck* h);
/* prev_flow may be NULL (first pair: u = v = 0).  delta < 0 disables the early exit
 * (reference: delta is None).  sweeps_done_host (may be NULL) receives the sweeps executed. */
TF_API int tf_hs_run(tf_horn_schunck* h, const uint8_t* left, const uint8_t* right,
              const float* prev_flow, double alpha, int max_iters, double decay, double delta,
              float* flow, int clip, int* sweeps_done_host, void* stream);

/* ---- pyramidal Lucas-Kanade: transflow/flow/methods/lukas_kanade.py:9-36 ------------------ */
typedef struct tf_lucas_kanade tf_lucas_kanade;
TF_API int tf_lk_create(tf_lucas_kanade** out, int height, int width, int win_size, int max_level,
                 int step);
TF_API int tf_lk_destroy(tf_lucas_kanade* h);
TF_API int tf_lk_run(tf_lucas_kanade* h, const uint8_t* left, const uint8_t* right, float* flow,
              int clip, void* stream);

/* ---- FlowSource.post_process: transflow/flow/sources/source.py:337-363 --------------------- */
/* In place.  mask (float32 HxW) may be NULL.  forward != 0 runs the clip / round / scatter
 * (last source in raster order wins) conversion of source.py:349-360; `owner` is an int32
 * (H, W) scratch plane that must be all zero on entry and is left all zero on exit.
 * The final clip (source.py:361-362) always runs. */
TF_API int tf_flow_postprocess(float* flow, const float* mask, int forward, int32_t* owner, int height,
                        int width, void* stream);

/* Same, writing the post-processed flow to `out` instead of in place (out may be PEER memory: the
 * last kernel of the producer stores straight into the accumulator rank's ring slot over NVLink).
 * out == NULL or out == flow -> in place. */
TF_API int tf_flow_postprocess_to(float* flow, const float* mask, int forward, int32_t* owner, float* out,
                                  int height, int width, void* stream);

/* ---- flow filters, mask, convolution kernel: transflow/flow/filters.py:36-68, source.py:339-348 ---------- */
enum { TF_FLOW_SCALE = 0, TF_FLOW_THRESHOLD = 1, TF_FLOW_CLIP = 2 };
#define TF_MAX_FLOW_OPS 8
/* One elementwise filter with its per-frame scalar (the reference evaluates a lambda of t).  strong != 0:
 * the scalar is a NumPy float64 (comparisons / products in float64, rounded to float32); strong == 0: a Python
 * number, which NumPy treats as float32 next to the float32 flow. */
typedef struct tf_flow_op {
    int kind;
    int strong;
    double value;
} tf_flow_op;
/* post_process with the filters fused in front of the mask multiply (everything in one pass over the flow).
 * ops is a HOST array (copied into the launch).  out == NULL -> in place. */
TF_API int tf_flow_postprocess_ex(float* flow, const tf_flow_op* ops, int n_ops, const float* mask, int forward,
                                  int32_t* owner, float* out, int height, int width, void* stream);
/* The forward direction's two passes on their own (source.py:349-360).  tf_flow_forward_claims: filters + mask + clip +
 * round half-even, then every moved pixel p claims its target with max(p + 1) into `claims` (int32 (H, W), all zero on
 * entry: numpy.put keeps the LAST source in raster order).  tf_flow_from_claims turns the claims into the flow the
 * reference returns (claimant - own position, 0 where unclaimed) and zeroes them again.  tf_layer_update_claims consumes
 * the claim plane directly, so a pipeline whose only consumer is the compositor never forms that flow. */
TF_API int tf_flow_forward_claims(const float* flow, const tf_flow_op* ops, int n_ops, const float* mask,
                                  int32_t* claims, int height, int width, void* stream);
TF_API int tf_flow_from_claims(int32_t* claims, float* flow_out, int height, int width, void* stream);
/* filters + mask only -> out (the stage in front of the convolution kernel). */
TF_API int tf_flow_filters(const float* flow, const tf_flow_op* ops, int n_ops, const float* mask, float* out,
                           int height, int width, void* stream);
/* scipy.signal.convolve2d(mode="same", boundary="fill") of both flow channels with a float64 kernel (kh, kw)
 * in DEVICE memory; float64 accumulation, float32 result.  forward != 0: the float32 value is chosen so that the
 * round-half-even of the forward conversion equals the float64 one.  Not in place. */
TF_API int tf_flow_convolve(const float* flow, const double* kernel, int kh, int kw, int forward, float* out,
                            int height, int width, void* stream);

/* ---- Pipeline._update_flow: transflow/pipeline.py:492-507 ---------------------------------------------------
 * Merge of several flow sources (FLOW_MERGING_FUNCTIONS, pipeline.py:149-158 + utils.py:359-381) and the integer
 * upscale of a smaller flow (utils.upscale_array, utils.py:417-418), in NumPy's float32 arithmetic. */
enum { TF_MERGE_FIRST = 0, TF_MERGE_SUM, TF_MERGE_AVERAGE, TF_MERGE_DIFFERENCE, TF_MERGE_PRODUCT, TF_MERGE_MASKBIN,
       TF_MERGE_MASKLIN, TF_MERGE_ABSMAX };
#define TF_MAX_MERGE_FLOWS 8
/* flows: HOST array of n device pointers (H, W, 2); out may alias flows[0]. */
TF_API int tf_flow_merge(const float* const* flows, int n, int mode, float* out, int height, int width, void* stream);
/* (H, W, 2) -> (H * hf, W * wf, 2): vectors scaled by (wf, hf), then block-replicated. */
TF_API int tf_flow_upscale(const float* flow, float* out, int height, int width, int wf, int hf, void* stream);
/* Flow visualisers of transflow/output/render.py:9-48 (pipeline.py:512-516, view_flow / view_flow_magnitude):
 * mode 0 = render2d(flow (H, W, 2)) with 4 colours; mode 1 = render1d(|flow|) with the magnitude of
 * pipeline.py:515 fused in, 2 colours; mode 2 = render1d(in (H, W)) for a scalar array.  `colors` is a HOST array
 * of n_colors x 3 floats (R, G, B in 0..255).  float32 arithmetic in NumPy's evaluation order -> rgb u8 (H, W, 3),
 * bit-exact with the reference. */
TF_API int tf_flow_render(const float* in, int mode, float scale, const float* colors, int n_colors, int binary,
                          uint8_t* rgb, int height, int width, void* stream);

/* ---- compositor: transflow/compositor/** ---------------------------------------------------- */
enum { TF_LAYER_MOVEREF = 0, TF_LAYER_SUM = 1, TF_LAYER_STATIC = 2, TF_LAYER_INTRODUCTION = 3 };
enum { TF_RESET_OFF = 0, TF_RESET_RANDOM = 1, TF_RESET_CONSTANT = 2, TF_RESET_LINEAR = 3 };

/* Mirrors LayerConfig (transflow/config.py:57-104); booleans are 0/1 ints. */
typedef struct {
    int32_t kind;
    int32_t transparent_pixels_can_move;
    int32_t pixels_can_move_to_empty_spot;
    int32_t pixels_can_move_to_filled_spot;
    int32_t moving_pixels_leave_empty_spot;
    int32_t reset_mode;
    int32_t reset_source;
    int32_t introduce_pixels_on_empty_spots;   /* no effect in the reference (quirk Q13) */
    int32_t introduce_pixels_on_filled_spots;
    int32_t introduce_moving_pixels;
    int32_t introduce_unmoving_pixels;         /* no effect in the reference (quirk Q13) */
    int32_t introduce_once;
    int32_t introduce_on_all_filled_spots;
    int32_t introduce_on_all_empty_spots;
    float reset_constant_step;
    float reset_random_factor; /* float32(reset_random_factor): threshold when reset_scale == NULL */
    double reset_linear_factor;
} tf_layer_config;

typedef struct {
    const uint8_t* pixels; /* device, (H, W, channels) */
    int32_t channels;      /* 3 or 4 */
    int32_t frame_number;  /* PixmapSourceInterface.frame_number (introduction layer) */
} tf_pixmap;

typedef struct tf_layer tf_layer;
TF_API int tf_layer_create(tf_layer** out, int height, int width, const tf_layer_config* cfg);
TF_API int tf_layer_destroy(tf_layer* l);
/* Any pointer may be NULL = the reference default (all true / all 1).  mask_src, mask_dst:
 * uint8 0/1 (H, W); mask_alpha: float32 (H, W); reset_scale: float32 (H, W) holding
 * reset_random_factor*reset_mask (random), reset_constant_step*reset_mask (constant) or
 * reset_mask (linear), evaluated by the host exactly as NumPy does. */
TF_API int tf_layer_set_masks(tf_layer* l, const uint8_t* mask_src, const uint8_t* mask_dst,
                       const float* mask_alpha, const float* reset_scale, void* stream);
/* Layer.set_sources (compositor/layers/reference.py:54-56): introduction masks uint8 0/1
 * (H, W) per source; re-applies the base source indices. */
TF_API int tf_layer_set_sources(tf_layer* l, int n_sources, const uint8_t* const* intro_masks_host_array,
                         void* stream);
/* Layer.update(flow).  pixmaps: host array of n descriptors (device pixel pointers).
 * random: float64 (H, W) draws of numpy.random.random (parity mode) or NULL -> counter-based
 * Philox keyed by (rng_seed, frame counter, pixel).
 * If rgb_inout != NULL the Layer.render + Compositor.render step for this layer is fused:
 * first_layer != 0 starts from the constant background colour, otherwise from rgb_inout. */
TF_API int tf_layer_update(tf_layer* l, const float* flow, const tf_pixmap* pixmaps_host, int n_pixmaps,
                    const double* random, uint64_t rng_seed, uint8_t* rgb_inout, int first_layer,
                    uint32_t background_rgb, void* stream);
/* Layer.update for a FORWARD flow given as the claim plane of tf_flow_forward_claims (consumed and zeroed): the record a
 * pixel fetches is its claimant's (movement.py:25-50 with flow = claimant - position), bit-identical to tf_layer_update
 * on the flow of tf_flow_from_claims.  Only for layers tf_layer_takes_claims() accepts (returns 1): a single-source
 * move-reference layer in its default movement configuration, rendered fused as the first layer; others take the flow. */
TF_API int tf_layer_takes_claims(const tf_layer* l, int n_pixmaps);
TF_API int tf_layer_update_claims(tf_layer* l, int32_t* claims, const tf_pixmap* pixmaps_host, int n_pixmaps,
                                  uint64_t rng_seed, uint8_t* rgb_inout, uint32_t background_rgb, void* stream);
/* Layer.render (compositor/layers/layer.py:32-34): alpha *= mask_alpha in place; -> uint8 rgba. */
TF_API int tf_layer_render(tf_layer* l, uint8_t* rgba_out, void* stream);
/* Compositor.render (compositor/compositor.py:31-40) over n rendered layers (device rgba). */
TF_API int tf_composite(const uint8_t* const* layer_rgba_host_array, int n_layers, uint32_t background_rgb,
                 uint8_t* rgb_out, int height, int width, void* stream);
/* Checkpoint interchange (pickled compositor, transflow/pipeline.py:225-242): state as the
 * reference's NumPy arrays.  data: int32 (H, W, depth); rgba: uint8 (H, W, 4) (for the
 * introduction layer rgba aliases data[..., :4] and may be NULL). */
TF_API int tf_layer_depth(const tf_layer* l);
TF_API int tf_layer_get_state(tf_layer* l, int32_t* data, uint8_t* rgba, void* stream);
TF_API int tf_layer_set_state(tf_layer* l, const int32_t* data, const uint8_t* rgba, void* stream);
/* Non-zero once a gather index left the frame (NumPy would have raised IndexError). */
TF_API int tf_layer_poll_error(tf_layer* l, void* stream);
/* Frames applied so far / introduced_once flag (for checkpoints). */
TF_API int tf_layer_get_counters(const tf_layer* l, uint64_t* frames, int* introduced_once);
TF_API int tf_layer_set_counters(tf_layer* l, uint64_t frames, int introduced_once);

/* ---- float displacement map + bilinear remap (extension) ----------------------------
 * The reference's WebGL variant in pixel units: extra/www/shaders/acc.frag:17-41 (map(p) = map_prev(p + f) + f
 * with f = scale * boxblur(flow), then decay), extra/www/shaders/remap.frag:10-18 (out(p) = pixmap(p + map(p))),
 * CLAMP_TO_EDGE, linear != 0 selects LINEAR sampling, else NEAREST (transflow.js:358-364).  The Python reference
 * has no such mode; the kernels are checked against oracle/floatmap_np.py. */
TF_API int tf_floatmap_accumulate(const float* map_prev, const float* flow, float* map_out, int height, int width,
                                  float scale, float decay, int blur_size, int linear, void* stream);
/* rgba_out (H, W, 4; alpha as the compositor's 0 / 1 flag) and / or rgb_inout (H, W, 3; composited like
 * Compositor.render: opaque pixels overwrite, the first layer paints background_rgb = 0xRRGGBB elsewhere). */
TF_API int tf_floatmap_remap(const float* map, const uint8_t* pixmap, int channels, int linear, uint8_t* rgba_out,
                             uint8_t* rgb_inout, int first_layer, uint32_t background_rgb, int height, int width,
                             void* stream);

/* ---- multi-GPU flow hand-off over NVLink (frame pairs sharded across ranks) ---------------- */
/* CUDA IPC plumbing so a producer rank's last flow kernel stores straight into the
 * accumulator rank's ring slot (peer memory), followed by a release flag. */
/* Raw device allocations (cudaMalloc, 256-byte aligned) for buffers that are shared through CUDA IPC:
 * a framework caching allocator hands out interior pointers, IPC handles name whole allocations. */
TF_API int tf_device_malloc(size_t bytes, void** dev_ptr_out);
TF_API int tf_device_free(void* dev_ptr);
TF_API int tf_ipc_get_handle(const void* dev_ptr, uint8_t handle_out_host[64]);
TF_API int tf_ipc_open_handle(const uint8_t handle_host[64], void** dev_ptr_out);
TF_API int tf_ipc_close_handle(void* dev_ptr);
/* Store `value` to a (possibly peer) 32-bit flag after all prior work on `stream`, with
 * system-scope release semantics. */
TF_API int tf_flag_signal(uint32_t* flag, uint32_t value, void* stream);
/* Make `stream` wait until *flag >= value (driver stream memory op; no spinning kernel). */
TF_API int tf_flag_wait_geq(uint32_t* flag, uint32_t value, void* stream);
/* Peer copy of `bytes` with 128-bit stores issued by a kernel (dst may be peer memory). */
TF_API int tf_copy_to_peer(void* dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRANSFLOW_B200_H */
