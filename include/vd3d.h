/* vd3d.h -- C ABI of libvd3d.so, the B200 (sm_100a) depth->stereo engine.
 *
 * The reference (VisionDepth3D) has no FFI: its boundary is the Python function
 * surface of core/render_3d.py and the `pipe` callable of core/render_depth.py
 * (SURVEY.md section 8(b)).  Each entry point below names the reference function it
 * replaces; visiondepth3d_b200/render_3d.py and render_depth.py bind them with
 * ctypes and re-expose the reference's names and signatures.
 *
 * Conventions: every function returns 0 on success or a negative vd3d_status;
 * nothing throws; a ctx is not re-entrant, distinct ctxs are independent.
 * Image pointers may be host or device memory as stated by `mem`
 * (VD3D_MEM_HOST / VD3D_MEM_DEVICE); outputs are caller-owned buffers.
 * There is no CPU fallback: without a CUDA device vd3d_create fails.
 */
#ifndef VD3D_H
#define VD3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vd3d_ctx vd3d_ctx;
typedef struct vd3d_depth vd3d_depth; /* depth-forward engine, see the end of this header */

typedef enum {
  VD3D_OK = 0,
  VD3D_ERR_CUDA = -1,        /* CUDA runtime error; see vd3d_last_error */
  VD3D_ERR_ARG = -2,         /* bad argument */
  VD3D_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented here */
  VD3D_ERR_NOMEM = -4,
  VD3D_ERR_STATE = -5        /* call sequence error (e.g. weights not loaded) */
} vd3d_status;

enum { VD3D_MEM_HOST = 0, VD3D_MEM_DEVICE = 1 };

/* output_format of render_sbs_3d / format_3d_output (core/render_3d.py:837-860) */
enum {
  VD3D_FMT_HALF_SBS = 0,
  VD3D_FMT_FULL_SBS = 1,
  VD3D_FMT_ANAGLYPH = 2,   /* "Red-Cyan Anaglyph" */
  VD3D_FMT_INTERLACED = 3, /* "Passive Interlaced" */
  VD3D_FMT_VR = 4          /* 2 x 1440x1600: eyes fitted by pad_to_aspect_ratio (INTER_AREA shrink only) */
};

/* which temporal state vd3d_reset_state clears */
enum {
  VD3D_STATE_GLOBAL = 1, /* module singletons: depth_ema_norm, conv_ema,
                            floating_window_tracker, bar_easer
                            (core/render_3d.py:284-285,500,511) */
  VD3D_STATE_CLIP = 2    /* per-render objects created at core/render_3d.py:1174-1182 */
};

/* kwargs of pixel_shift_cuda (core/render_3d.py:561-590).  Shifts are Python
 * floats (double) in the reference and are rounded to fp32 where a tensor op
 * consumes them, which the kernels reproduce. */
typedef struct {
  double fg_shift, mg_shift, bg_shift;
  int32_t blur_ksize;
  double feather_strength;
  double max_pixel_shift_percent;
  double parallax_balance;
  double zero_parallax_strength;
  int32_t use_subject_tracking;
  int32_t enable_floating_window;
  int32_t enable_feathering;
  int32_t enable_edge_masking;
  double convergence_strength;
  int32_t enable_dynamic_convergence;
  double depth_pop_gamma, depth_pop_mid;
  double depth_stretch_lo, depth_stretch_hi;
  double fg_pop_multiplier, bg_push_multiplier;
  double subject_lock_strength;
} vd3d_shift_params;

/* arguments of render_sbs_3d (core/render_3d.py:933-985) that reach the frame loop */
typedef struct {
  int32_t output_width, output_height;
  double fg_shift, mg_shift, bg_shift;
  double sharpness_factor;
  int32_t output_format;      /* VD3D_FMT_* */
  double aspect_ratio;        /* aspect_ratios[selected_aspect_ratio.get()] */
  double dof_strength;
  double feather_strength;
  int32_t blur_ksize;
  int32_t use_subject_tracking, use_floating_window;
  double max_pixel_shift_percent;
  int32_t preserve_original_aspect;
  double zero_parallax_strength;
  int32_t enable_edge_masking, enable_feathering;
  int32_t original_video_width, original_video_height; /* 0 = None */
  double convergence_strength;
  int32_t enable_dynamic_convergence;
  double ipd_factor;
  double color_saturation, color_contrast, color_brightness;
} vd3d_render_params;

/* sizes derived at core/render_3d.py:1074-1138,1250-1259 */
typedef struct {
  int32_t crop_x0, crop_y0, crop_w, crop_h;
  int32_t target_eye_w, target_eye_h;
  int32_t resized_width, resized_height;
  int32_t per_eye_w, per_eye_h;
  int32_t out_width, out_height;
} vd3d_size_plan;

/* per-frame scalars, for parity tests and progress reporting */
typedef struct {
  float pct_lo, pct_hi;       /* DepthPercentileEMA state after this frame */
  float subj_raw, stretch_lo, stretch_hi, subj_shaped;
  float subj_norm;            /* estimate_subject_depth(depth_tensor) (1334/1390) */
  double dyn_scale, fg, mg, bg;
  double zero_parallax_offset;
  double focal_depth, motion_metric;
  double stable_zero;
  int32_t bar_width, bar_side; /* side: 0 none, 1 right, 2 left */
} vd3d_frame_info;

/* sizeof() of the ABI structs, so bindings can verify their layout:
 * 0 vd3d_shift_params, 1 vd3d_render_params, 2 vd3d_size_plan, 3 vd3d_frame_info */
int vd3d_struct_size(int which);

/* ---- lifecycle ------------------------------------------------------- */
int vd3d_create(int device, vd3d_ctx** out);
void vd3d_destroy(vd3d_ctx* ctx);
const char* vd3d_last_error(vd3d_ctx* ctx); /* ctx may be NULL: last create error */
int vd3d_reset_state(vd3d_ctx* ctx, uint32_t which);
/* pinned host memory for the end-to-end path */
void* vd3d_host_alloc(size_t bytes);
void vd3d_host_free(void* p);
/* the stream all work of this ctx is enqueued on (cudaStream_t) */
void* vd3d_stream(vd3d_ctx* ctx);
int vd3d_sync(vd3d_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench: gpu_launches) */
uint64_t vd3d_launch_count(vd3d_ctx* ctx);
/* device-side stage timing for the roofline report (CUDA events on the ctx stream):
 * stage 0 = whole DIBR frame (ingest..pack), 1 = compose kernel, 2 = depth forward.
 * vd3d_profile_collect synchronises, returns the summed time and sample count since the
 * last collect, and resets. */
int vd3d_profile(vd3d_ctx* ctx, int enable);
int vd3d_profile_collect(vd3d_ctx* ctx, int stage, double* total_ms, int* count);
/* replay the per-frame kernel sequence from a captured CUDA graph (default 1) */
int vd3d_set_graphs(vd3d_ctx* ctx, int enable);
/* DIBR arithmetic mode.  0 (default): persistent statistics kernel + fused warp/feather/compose/pack kernel, fp32
 * hardware pow/exp, separable box sums -- inside the 1e-3 / 1-LSB tolerances of the reference's outputs.
 * 1: one kernel per reference op, correctly rounded transcendentals, the reference's row-major summation order
 * (bit-for-bit with oracle/dibr.py; what the exactness tests drive).  Env VD3D_EXACT=1 makes 1 the default. */
int vd3d_set_exact(vd3d_ctx* ctx, int enable);
int vd3d_get_exact(vd3d_ctx* ctx);
/* 1 while CUDA-graph replay is on; 0 after vd3d_set_graphs(ctx, 0) or after a failed capture fell back to eager launches */
int vd3d_graphs_active(vd3d_ctx* ctx);

/* ---- DIBR ------------------------------------------------------------ */
/* pixel_shift_cuda (core/render_3d.py:561-712).
 * rgb: f32 planar RGB [3,in_h,in_w] in 0..1; depth: f32 [in_h,in_w];
 * left/right: u8 BGR interleaved [height,width,3]; shift: f32 [height,width] or NULL.
 * Mutates the floating-window tracker held by ctx (the module singleton). */
int vd3d_pixel_shift(vd3d_ctx* ctx, const float* rgb, const float* depth, int in_h, int in_w,
                     int width, int height, const vd3d_shift_params* p, uint8_t* left_bgr,
                     uint8_t* right_bgr, float* shift, int mem, vd3d_frame_info* info);

/* sizing rules of render_sbs_3d (core/render_3d.py:1074-1138,1236-1259) */
int vd3d_plan_sizes(int src_w, int src_h, const vd3d_render_params* rp, vd3d_size_plan* out);

/* one iteration of the render_sbs_3d frame loop (core/render_3d.py:1227-1419):
 * frame_to_tensor/depth_to_tensor, aspect crop, resize, TemporalDepthFilter,
 * DepthPercentileEMA, ShiftSmoother, dynamic parallax scale, pixel_shift_cuda,
 * FocalDepthTracker, DOF, colour grade, floating-window bars, sharpen, eye fit,
 * format_3d_output.  frame/depth: u8 BGR [src_h,src_w,3] (depth may also be one
 * channel: depth_channels = 1); out: u8 BGR [out_height,out_width,3]. */
int vd3d_render_frame(vd3d_ctx* ctx, const uint8_t* frame_bgr, const uint8_t* depth, int depth_channels,
                      int src_h, int src_w, const vd3d_render_params* rp, uint8_t* out_bgr, int mem,
                      vd3d_frame_info* info);

/* Throughput form of the same loop: n frames, host or device arrays of
 * frame pointers; H2D of frame i+1 and D2H of frame i-1 overlap the kernels of
 * frame i on separate streams.  State carries across calls exactly as across
 * loop iterations.  infos may be NULL. */
int vd3d_render_clip(vd3d_ctx* ctx, int n, const uint8_t* const* frames, const uint8_t* const* depths,
                     int depth_channels, int src_h, int src_w, const vd3d_render_params* rp,
                     uint8_t* const* outs, int mem, vd3d_frame_info* infos);

/* depth + stereo in one pipelined loop: depth of frame i is inferred on the GPU by `depth`
 * (vd3d_depth_infer_device) and handed to the DIBR loop in HBM as a 1-channel u8 map -- the
 * in-memory replacement of the reference's XVID depth-video round trip. */
int vd3d_render_clip_depth(vd3d_ctx* ctx, vd3d_depth* depth, int n, const uint8_t* const* frames, int src_h,
                           int src_w, const vd3d_render_params* rp, uint8_t* const* outs, int mem);

/* ---- exact frame sharding --------------------------------------------------
 * The loop is stateful (7 EMAs / trackers + the temporal depth plane).  A rank that renders frames
 * [a, b) exactly needs the state after frame a-1: vd3d_advance_state runs one loop iteration
 * WITHOUT rendering (all state updates, ~1/3 of the DIBR cost), vd3d_export_state /
 * vd3d_import_state move the state (struct + two f32 planes of target_eye size) between contexts
 * or GPUs (NCCL send/recv of the blob). */
int vd3d_advance_state(vd3d_ctx* ctx, const uint8_t* frame_bgr, const uint8_t* depth, int depth_channels,
                       int src_h, int src_w, const vd3d_render_params* rp, int mem);
size_t vd3d_state_bytes(vd3d_ctx* ctx);
int vd3d_export_state(vd3d_ctx* ctx, void* dst, size_t capacity, int mem);
int vd3d_import_state(vd3d_ctx* ctx, const void* src, size_t bytes, int mem);

/* drop the per-ctx clones / graphs built for `depth`; call before vd3d_depth_destroy(depth) */
int vd3d_release_depth(vd3d_ctx* ctx, vd3d_depth* depth);

/* Validate a render configuration without launching anything (sizes, eye-fit mode, DOF kernel bank): what
 * render_sbs_3d (core/render_3d.py:1086-1138) decides before it opens its writer.  0 or a negative error + message. */
int vd3d_check_config(vd3d_ctx* ctx, int src_h, int src_w, const vd3d_render_params* rp);

/* stage entry points (same kernels, exposed for stage-isolated parity tests) */
/* apply_color_grade (core/render_3d.py:734-767): f32 RGB planes [3,h,w] in 0..1 -> same (saturation around Rec.709
 * luma, contrast around 0.5, additive brightness, clamp) */
int vd3d_color_grade(vd3d_ctx* ctx, const float* rgb, int h, int w, double saturation, double contrast,
                     double brightness, float* out, int mem);
/* cv2.resize(u8 plane [h,w], (ow,oh), interpolation=cv2.INTER_CUBIC): the resize the depth writer applies to the u8
 * depth (core/render_depth.py:1917, 193): float32 bicubic (A = -0.75), round half to even */
int vd3d_resize_cubic_u8(vd3d_ctx* ctx, const uint8_t* src, int h, int w, uint8_t* dst, int oh, int ow, int mem);
/* the same on interleaved u8 [h,w,ch] (ch <= 4): run_esrgan's INTER_CUBIC chain on BGR frames
 * (core/merged_pipeline.py:262-266) */
int vd3d_resize_cubic(vd3d_ctx* ctx, const uint8_t* src, int h, int w, int ch, uint8_t* dst, int oh, int ow, int mem);
/* cv2.addWeighted(a, alpha, b, beta, 0) on n u8 values: blend_images (core/merged_pipeline.py:233-238) */
int vd3d_add_weighted(vd3d_ctx* ctx, const uint8_t* a, double alpha, const uint8_t* b, double beta, size_t n, uint8_t* dst,
                      int mem);
/* apply_sharpening (717-732) on u8 BGR [h,w,3] */
int vd3d_sharpen(vd3d_ctx* ctx, const uint8_t* src, int h, int w, double factor, uint8_t* dst, int mem);
/* heal_missing_pixels (431-459; the reference's "gradient-blend occlusion fill", defined but not called by
 * its render loop): f32 RGB planes [3,h,w] warped + original, optional edge mask [h,w] -> f32 [3,h,w] */
int vd3d_heal(vd3d_ctx* ctx, const float* warped, const float* original, const float* edge_mask_or_null, int h, int w,
              double heal_strength, float* out, int mem);
/* eye fit on one u8 BGR image [h,w,3] -> [target_h,target_w,3]: keep_aspect != 0 = pad_to_aspect_ratio (101-131, black
 * canvas), 0 = cv2.resize(..., INTER_AREA) as in the Half-SBS branch (1413-1414).  Shrinking only: identity, integer
 * factors, or cv2's general (fractional) area tables; enlarging returns VD3D_ERR_UNSUPPORTED */
int vd3d_fit_eye(vd3d_ctx* ctx, const uint8_t* src, int h, int w, int target_w, int target_h, int keep_aspect,
                 uint8_t* dst, int mem);
/* host-only test hook (no GPU needed): the cv2 area-resize tables vd3d_fit_eye / the frame path build for a
 * ssize -> dsize shrink; ofs/cnt [dsize], alpha [dsize*cap]; returns the largest tap count or a negative error */
int vd3d_area_table(int ssize, int dsize, int* ofs, int* cnt, float* alpha, int cap);
/* same for the enlarging case (cv2 emulates INTER_AREA with fixed-point bilinear weights): ofs [dsize], a01 [2*dsize] */
int vd3d_area_linear_table(int ssize, int dsize, int* ofs, int* a01);
/* format_3d_output / generate_anaglyph_3d (837-883) on two same-size u8 BGR eyes [h,w,3]:
 * SBS -> [h,2w,3]; anaglyph / interlaced -> [h,w,3] */
int vd3d_pack(vd3d_ctx* ctx, const uint8_t* left, const uint8_t* right, int h, int w, int fmt, uint8_t* dst, int mem);
/* apply_dof_cuda (769-834) + apply_color_grade (734-767) + tensor_to_frame on a u8 BGR eye;
 * depth01: f32 [dh,dw] resized bilinearly to [h,w] as at 1347-1350; max_sigma<=0 skips DOF */
int vd3d_dof_grade(vd3d_ctx* ctx, const uint8_t* eye_bgr, int h, int w, const float* depth01, int dh, int dw,
                   double focal, double max_sigma, double sat, double con, double bri, uint8_t* dst, int mem);

/* ---- depth forward (Depth-Anything-V2: DINOv2 ViT + DPT neck/head) ----------
 * Replaces model.forward inside transformers' depth-estimation pipeline as called by
 * hf_batch_safe_pipe (core/render_depth.py:1106-1119).  GEMMs / 3x3 convs run on tcgen05
 * tensor cores (f16 operands, fp32 accumulation in TMEM) fed by TMA. */
typedef struct {
  int32_t hidden, layers, heads; /* 384/12/6, 768/12/12, 1024/24/16 */
  int32_t taps[4];               /* out_indices of the backbone (1-based layer numbers) */
  int32_t neck[4];               /* neck_hidden_sizes */
  int32_t fusion;                /* fusion_hidden_size */
  int32_t image_h, image_w;      /* processed size, multiples of 14 (518 x 924 for 16:9) */
} vd3d_depth_config;
int vd3d_depth_create(const vd3d_depth_config* cfg, void* cuda_stream, vd3d_depth** out);
void vd3d_depth_destroy(vd3d_depth* e);
/* CUDA-event timing of the fc1 GEMM launches (k_umma_gemm<128,3>; M = tokens, N = 4*hidden, K = hidden) */
int vd3d_depth_profile(vd3d_depth* e, int enable);
int vd3d_depth_profile_collect(vd3d_depth* e, double* total_ms, int* count, double* gflop_per_launch);
/* tuning aid: after vd3d_depth_profile(e, 2), every launch class of an eager forward is bracketed by CUDA events; this
   returns "tag total_ms spans" lines (NUL-terminated, truncated to cap) and clears the record */
int vd3d_depth_profile_spans(vd3d_depth* e, char* out, size_t cap);
/* another instance on `cuda_stream` sharing e's weights (own activations); destroy it before e */
int vd3d_depth_clone(vd3d_depth* e, void* cuda_stream, vd3d_depth** out);
const char* vd3d_depth_last_error(vd3d_depth* e);
uint64_t vd3d_depth_launch_count(vd3d_depth* e);
void vd3d_depth_add_launches(vd3d_depth* e, uint64_t n); /* bookkeeping for CUDA-graph replays */
/* upload one prepared weight tensor (names / layouts: visiondepth3d_b200/depth_weights.py) */
int vd3d_depth_set_tensor(vd3d_depth* e, const char* name, const void* host_data, size_t bytes);
/* pixel_values f32 [3,image_h,image_w] -> predicted_depth f32 [image_h,image_w] */
int vd3d_depth_forward(vd3d_depth* e, const float* pixel_values, float* depth_out, int mem);
/* the whole depth stage of the reference for one frame: DPT image processor (antialiased
 * bicubic to image_h x image_w, /255, mean/std) -> forward -> post_process_depth_estimation
 * (bicubic back to h x w) -> convert_depth_to_grayscale min-max u8 (core/render_depth.py:
 * 1113-1119, 605-611).  _device: all pointers on the GPU, enqueued without synchronising
 * (this is what vd3d_render_clip uses for the in-memory depth -> stereo handoff). */
int vd3d_depth_infer(vd3d_depth* e, const uint8_t* frame_bgr, int h, int w, float* depth_f32, uint8_t* depth_u8,
                     int invert);
int vd3d_depth_infer_device(vd3d_depth* e, const uint8_t* frame_bgr_dev, int h, int w, uint8_t* depth_u8_dev,
                            float* depth_f32_dev_or_null, int invert);
/* The same for B (1..8) frames of one size with ONE batched forward: the token-wise GEMMs see the stacked token
 * matrix of all frames (the reference hands the whole list to the HF pipeline too, core/render_depth.py:1113-1119).
 * Arrays of B pointers; depth_f32 / depth_u8 (or single entries) may be NULL. */
int vd3d_depth_infer_batch(vd3d_depth* e, int B, const uint8_t* const* frames_bgr, int h, int w, float* const* depth_f32,
                           uint8_t* const* depth_u8, int invert);
int vd3d_depth_infer_batch_device(vd3d_depth* e, int B, const uint8_t* const* frames_bgr_dev, int h, int w,
                                  uint8_t* const* depth_u8_dev, float* const* depth_f32_dev, int invert);
/* frames per depth forward inside vd3d_render_clip_depth (1..8, default 4; env VD3D_DEPTH_BATCH) */
int vd3d_set_depth_batch(vd3d_ctx* ctx, int frames);
int vd3d_get_depth_batch(vd3d_ctx* ctx);
/* ---- Real-ESRGAN upscale stage (core/merged_pipeline.py:221-267) --------------------------------
 * The reference runs an ONNX export of xinntao/Real-ESRGAN's SRVGGNetCompact (realesr-general-x4v3: 32 body convs;
 * realesr-animevideov3: 16) through ONNXRuntime.  vd3d_sr_create returns an engine container (weights by name through
 * vd3d_depth_set_tensor: "sr.c{i}.w" f16 [Cout, 9*64], "sr.c{i}.b" f32, "sr.a{i}" f32 PReLU slopes; destroy with
 * vd3d_depth_destroy).  vd3d_sr_forward: BGR u8 [h,w,3] -> BGR u8 [4h,4w,3] = preprocess_esr -> network (tcgen05
 * implicit-GEMM convs, f16 activations, fp32 accumulate, fp32 last layer) -> PixelShuffle + input -> postprocess_esr. */
int vd3d_sr_create(void* cuda_stream, vd3d_depth** out);
int vd3d_sr_forward(vd3d_depth* e, const uint8_t* frame_bgr, int h, int w, int num_conv, uint8_t* out_bgr, int mem);
/* copy an internal activation buffer to the host (parity triage: "x", "tap0.0".., "f0".., "fused3") */
int vd3d_depth_get_buffer(vd3d_depth* e, const char* name, void* host_out, size_t bytes);
/* unit-test hooks for the tensor-core kernels */
int vd3d_gemm_f16(vd3d_depth* e, const void* A_f16, const void* B_f16, int M, int N, int K, float* C_host, int bn);
/* tuning hook: average launch time (ms) of one M x N x K GEMM on device-resident operands. variant: 0 <128,3> two
   CTAs/SM, 1 <128,6>, 2 <128,3> one CTA/SM, 10-12 CTA pair 256x256 (6/4/2 stages), 20-22 CTA pair 256x128;
   dbg: 0 normal, 1 no MMAs (operand-feed rate), 2 no TMA loads (MMA rate); act: 0 none, 1 GELU */
int vd3d_gemm_bench(vd3d_depth* e, int M, int N, int K, int variant, int dbg, int act, int iters, float* ms_out);
int vd3d_conv_f16(vd3d_depth* e, const void* in_nhwc_f16, int H, int W, int cin, const void* w_f16, int cout,
                  int k3, const float* bias, int relu, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* VD3D_H */
