/* bd_b200.h -- C ABI of the B200 (sm_100a) implementation of the building-detection hot path.
 *
 * The reference (A511-1103/building-detection) has no FFI of its own: its hot path is the Python
 * call chain predict.py:run_model -> detection -> model.predict (TensorFlow runtime) ->
 * model_fuse.py:model_confuse (OpenCV) -> edge_3.py:_detection (OpenCV).  This header is the
 * boundary a binding for that path targets (ctypes stub in INTEGRATION.md); every entry point
 * names the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; bd_last_error() gives the message
 *     (thread-local, valid until the next call on the same thread);
 *   - pointers named *_dev are CUDA device pointers on the context's device, *_host are host
 *     pointers; the caller owns all of them.  The library owns only weights and workspace;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise;
 *   - feature maps are NHWC; a "tensor ref" is a channel slice [c0, c0+c) of a plan buffer;
 *   - one bd_ctx per GPU, not thread-safe.
 */
#ifndef BD_B200_H
#define BD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bd_ctx bd_ctx;
typedef struct bd_plan bd_plan;

enum bd_dtype { BD_F16 = 0, BD_F32 = 1 };
enum bd_buf_kind { BD_MAP = 0 /* (N,H,W,C) */, BD_VEC = 1 /* (N,C) fp32 */ };
enum bd_act { BD_ACT_NONE = 0, BD_ACT_RELU = 1, BD_ACT_SIGMOID = 2 };
enum bd_gate_mode { BD_GATE_SE = 0, BD_GATE_SCSE = 1, BD_GATE_BAM = 2 };
enum bd_conv_path { BD_CONV_DIRECT = 0 /* CUDA cores, generic */, BD_CONV_UMMA = 1 /* tcgen05 implicit GEMM */,
                    BD_CONV_SMALL = 2 /* CUDA cores, <= 16 output channels, HBM-bound (heads, BAM gate convs) */ };

typedef struct bd_tref { int32_t buf, c0, c; } bd_tref;

#define BD_MAX_TAPS 18

/* Fused convolution: y = act_post( act_pre( sum_t W[t] . x[h*stride+dy[t], w*stride+dx[t]] + bias ) + res )
 * written to y[(h*out_scale+out_oy), (w*out_scale+out_ox)].  Out-of-range taps read zero ('same'
 * padding).  Replaces Conv2D / Conv2DTranspose (+BatchNormalization +Activation +Add) layer groups of
 * predict_model/{res34,hrnet,v3plus,scse,bam}.py executed by tf.keras.Model.predict (predict.py:109).
 * w: fp16 bit patterns, layout [ntaps][cout][cin]; bias: fp32 [cout] (BatchNorm already folded). */
typedef struct bd_conv_desc {
  bd_tref x, y, res;            /* res.buf < 0: no residual */
  int32_t ntaps;
  int32_t dy[BD_MAX_TAPS], dx[BD_MAX_TAPS];
  int32_t stride, ho, wo;       /* logical output grid (before out_scale) */
  int32_t act_pre, act_post;    /* bd_act (NONE or RELU) */
  int32_t out_scale, out_oy, out_ox;
  int32_t path;                 /* bd_conv_path */
  const uint16_t* w_host;
  const float* bias_host;
  /* Fused SeparableConv2D (v3plus.py / bam.py Xception blocks): when dw_w_host != NULL this 1x1 stride-1 UMMA
   * convolution is the pointwise stage and x is the input of the DEPTHWISE 3x3 ('same', stride 1) stage, whose
   * weights are dw_w_host [9][x.c] fp32 (fp16-representable) and whose input optionally passes a ReLU.  The
   * depthwise result never goes to HBM; it is bit-identical to bd_plan_add_dwconv + bd_plan_add_conv. */
  const float* dw_w_host;
  int32_t dw_relu_in;
} bd_conv_desc;

/* ---- context ---------------------------------------------------------------------------- */
int bd_create(int device, bd_ctx** out);
void bd_destroy(bd_ctx* ctx);
const char* bd_last_error(void);
const char* bd_version(void);
/* number of kernel launches issued by this library on this context so far */
int64_t bd_launch_count(bd_ctx* ctx);

/* debug: with BD_UMMA_TRACE=1 in the environment when a plan is finalized, CTA 0 of the most recently built
 * tcgen05 conv records (tile, k-block, clock64) events per role; host_dst receives 4 x 4096 int64 (per role:
 * count, then triples) */
int bd_debug_read_trace(bd_ctx* ctx, long long* host_dst, int max_events);

/* ---- network plans: replace tf.keras.Model construction + Model.predict (predict.py:17-54,109) */
int bd_plan_create(bd_ctx* ctx, int batch, bd_plan** out);
void bd_plan_destroy(bd_plan* plan);
/* returns the buffer id (>= 0) or a negative error */
int bd_plan_add_buffer(bd_plan* plan, int h, int w, int c, int dtype, int kind);
int bd_plan_add_conv(bd_plan* plan, const bd_conv_desc* d);
/* depthwise 3x3 (first half of SeparableConv2D, v3plus.py:187-278); w fp32 [9][c], optional ReLU on load */
int bd_plan_add_dwconv(bd_plan* plan, bd_tref x, bd_tref y, int stride, int pad_t, int pad_l, int relu_in,
                       const float* w_host);
/* MaxPool2D k x k (scse.py:54, res34.py:152-154, v3plus.py:192); padding reads -inf */
int bd_plan_add_maxpool(bd_plan* plan, bd_tref x, bd_tref y, int k, int stride, int pad_t, int pad_l);
/* y = act(sum_i nearest_upsample(x_i, f_i)), n <= 4 (UpSampling2D / tf.add / concat slices, hrnet.py:99-162) */
int bd_plan_add_addn(bd_plan* plan, int n, const bd_tref* xs, const int32_t* fs, bd_tref y, int act);
/* GlobalAveragePooling2D -> fp32 vector buffer */
int bd_plan_add_gap(bd_plan* plan, bd_tref x, int y_vec);
/* y = act(W . sum_i x_i + b); w fp32 [cout][cin] (Dense / 1x1 conv on pooled vectors, BN folded) */
int bd_plan_add_dense(bd_plan* plan, int n_in, const int32_t* x_vecs, int y_vec, int cin, int cout, int act,
                      const float* w_host, const float* b_host);
/* attention gates: SE x*v (res34.py:102-104); scSE x*(sigmoid(w.x+b) + v) (scse.py:20-46);
 * BAM x*(1+sigmoid(v + s)) (bam.py:57-71) */
int bd_plan_add_gate(bd_plan* plan, int mode, bd_tref x, bd_tref y, int v_vec, bd_tref s, const float* w_host,
                     float b);
/* selective-kernel fusion (v3plus.py:114-136): softmax over 5 logit vectors, weighted sum of 4 maps and
 * the pooled vector g, then y = relu(acc*scale + shift) */
int bd_plan_add_skfuse(bd_plan* plan, const bd_tref* xs4, int g_vec, const int32_t* logit_vecs5, bd_tref y,
                       const float* scale_host, const float* shift_host);
/* broadcast a pooled vector over a map slice (UpSampling2D of a 1x1 map, v3plus.py:302-304) */
int bd_plan_add_bcast(bd_plan* plan, int v_vec, bd_tref y);
/* allocates the arena, uploads weights, encodes TMA tensor maps.  input_buf: the network input im2col'ed for its
 * 3x3 stem convolution of stride s (1 or 2): fp16 (N, 512/s, 512/s, 32), channel (kh*3+kw)*3+c = 255 * x at
 * input pixel (oy*s+kh-pad, ox*s+kw-pad) with TF 'same' padding, zero outside the tile, 27 channels padded to 32
 * (what bd_tiles_gather writes; the stem weights carry the 1/255); logits_buf: fp32 2-channel map at
 * 512/logits_up resolution. */
int bd_plan_finalize(bd_plan* plan, int input_buf, int logits_buf, int logits_up);
/* One forward of the whole network.  x_dev: fp32 (N,512,512,3) NHWC in [-1,1] as handed to model.predict
 * (converted on the device into the input buffer), or NULL to use what is already in the input buffer (e.g.
 * written by bd_tiles_gather).  probs_dev: fp32 (N,512,512,2) or NULL;
 * mask_dev: u8 (N,512,512), 1 where class 1 wins (argmax, ties -> class 0, predict.py:110) or NULL.
 * probs_dev must be 16-byte aligned and mask_dev 4-byte aligned (the head kernel stores four pixels per thread; any
 * cudaMalloc'ed or torch-allocated buffer is); with arena reuse (the default) the input buffer does not survive a
 * forward -- pass x_dev, or refill it (bd_tiles_gather_at), before every run. */
int bd_plan_run(bd_plan* plan, const float* x_dev, float* probs_dev, uint8_t* mask_dev, void* stream);
/* only the softmax / argmax head (last op) on whatever the logits buffer currently holds */
int bd_plan_run_head(bd_plan* plan, float* probs_dev, uint8_t* mask_dev, void* stream);
/* host-buffer convenience (synchronous): H2D, run, D2H.  Replaces model.predict(ndarray). */
int bd_plan_run_host(bd_plan* plan, const float* x_host, float* probs_host, uint8_t* mask_host);
/* device pointer / byte size of a plan buffer (tests, and writing the input buffer in place) */
void* bd_plan_buffer_ptr(bd_plan* plan, int buf);
size_t bd_plan_buffer_bytes(bd_plan* plan, int buf);
int bd_plan_read_buffer(bd_plan* plan, int buf, void* host_dst, size_t bytes);
int bd_plan_write_buffer(bd_plan* plan, int buf, const void* host_src, size_t bytes);
size_t bd_plan_arena_bytes(bd_plan* plan);
/* Arena reuse (default on; BD_ARENA_REUSE=0 in the environment turns the default off): map buffers whose lifetimes --
 * first to last plan step touching them -- are disjoint share address ranges, which cuts the activation arena of the
 * five networks 5-10x (what makes batch 64 fit next to a 20 000^2 scene).  With reuse on only the input, the logits and
 * the pooled vectors can be read back (bd_plan_read_buffer refuses the others); call with on = 0 BEFORE
 * bd_plan_finalize for debugging plans whose intermediates are inspected.  bd_plan_arena_bytes_flat = the arena size
 * without reuse. */
int bd_plan_set_arena_reuse(bd_plan* plan, int on);
size_t bd_plan_arena_bytes_flat(bd_plan* plan);
/* first / last plan step (index of the bd_plan_add_* call) touching a buffer (INT32_MAX / -1: never; the input starts at
 * -1, the logits end at INT32_MAX) and whether its range is shared -- what the arena allocator worked from (tests) */
int bd_plan_buffer_lifetime(bd_plan* plan, int buf, int* first, int* last, int* shared);
int bd_plan_num_launches(bd_plan* plan);
/* per-op device times of one run (CUDA events around every op; for profiling only): ms_out[num_ops] */
int bd_plan_num_ops(bd_plan* plan);
int bd_plan_time_ops(bd_plan* plan, float* ms_out, void* stream);
/* kind (0=conv umma, 1=conv direct, 2=other), and algorithmic flops of op i */
int bd_plan_op_info(bd_plan* plan, int i, int* kind, double* flops);

/* Plan files: a finalized plan as the sequence of bd_plan_add_* calls that built it (buffers, fused ops, BN-folded fp16
 * weights).  The graph lowering of the five predict_model modules lives in the Python builder; tools/export_plans.py writes one
 * file per (model, batch), and a host in ANY language loads them and runs the whole path -- bd_create, bd_plan_load x 5,
 * bd_scene_run, bd_fuse, bd_contours (examples/host_scene.c).  Replaces predict.py:17-54 (load_model) for such hosts. */
int bd_plan_save(bd_plan* plan, const char* path);
int bd_plan_load(bd_ctx* ctx, const char* path, bd_plan** out);
/* stride (1 or 2) of the 3x3 stem the plan's input buffer is laid out for: the stem_stride of bd_tiles_gather */
int bd_plan_input_stride(bd_plan* plan);

/* ---- tiler / stitcher: replace predict.py:detection (90-116) ------------------------------- */
/* Gather n 512x512 tiles whose top-left corners are (ys[i], xs[i]) from a BGR u8 scene (h,w,3) into the plan
 * input layout for a stem of stride stem_stride (see bd_plan_finalize): values 2*pixel-255 = 255 * (pixel/127.5 - 1)
 * exactly (predict.py:93), zero outside the scene (predict.py:102-104 pads the *normalised* image with zeros) and
 * outside the tile ('same' padding of the stem). */
int bd_tiles_gather(bd_ctx* ctx, const uint8_t* scene_bgr_dev, int h, int w, const int32_t* ys_host,
                    const int32_t* xs_host, int n, void* x_dev, int stem_stride, void* stream);
/* OR tile masks (n,512,512) into the scene mask (h,w): scene |= 255 where the tile says class 1
 * (predict.py:113-114: int8 accumulate then >=1 -> 255). */
int bd_stitch_or(bd_ctx* ctx, const uint8_t* tile_masks_dev, const int32_t* ys_host, const int32_t* xs_host,
                 int n, uint8_t* scene_mask_dev, int h, int w, void* stream);
/* The same two steps for a whole scene without per-batch host->device copies: upload the origins of ALL tiles
 * once (predict.py:105-107 enumerates them), then address batches by their first tile. */
int bd_tiles_set_origins(bd_ctx* ctx, const int32_t* ys_host, const int32_t* xs_host, int n, void* stream);
int bd_tiles_gather_at(bd_ctx* ctx, const uint8_t* scene_bgr_dev, int h, int w, int first, int n, void* x_dev,
                       int stem_stride, void* stream);
int bd_stitch_or_at(bd_ctx* ctx, const uint8_t* tile_masks_dev, int first, int n, uint8_t* scene_mask_dev, int h, int w,
                    void* stream);

/* The whole tiled forward of a scene for n_plans models in ONE call (the loop of predict.py:98-114 per model; the
 * reference's run_model, predict.py:75-87, calls it five times): tile origins (ys, xs) in the reference's order, plans
 * of one common batch size, masks_dev = n_plans planes of (h, w) u8 that must be zeroed by the caller and receive
 * 255 where any covering tile says "building".  A ragged last batch runs through the same plans.  Each plan's forward
 * is replayed from a CUDA graph captured on first use (BD_GRAPHS=0: direct launches). */
int bd_scene_run(bd_ctx* ctx, bd_plan* const* plans, int n_plans, const uint8_t* scene_bgr_dev, int h, int w,
                 const int32_t* ys_host, const int32_t* xs_host, int n_tiles, uint8_t* masks_dev, void* stream);
/* 1 when bd_scene_run replays this plan from a captured CUDA graph, 0 when capture was refused or is turned off */
int bd_plan_uses_graph(bd_plan* plan);
/* Opt-in fusion by PROBABILITY AVERAGING instead of the reference's 3-of-5 vote on argmax masks (the north star's "per-pixel
 * fusion of their probability maps"): mask_dev = ONE (h, w) u8 plane, zeroed by the caller, set to 255 where, in any
 * covering tile, the mean over the n_plans models of P(building) exceeds 0.5.  Follow with bd_mask_cleanup (the
 * reference's final clean-up, model_fuse.py:339-346) and bd_contours. */
int bd_scene_run_average(bd_ctx* ctx, bd_plan* const* plans, int n_plans, const uint8_t* scene_bgr_dev, int h, int w,
                         const int32_t* ys_host, const int32_t* xs_host, int n_tiles, uint8_t* mask_dev, void* stream);
/* device bytes the context currently holds for the scene-level stages (fusion / contour scratch, tile masks) */
size_t bd_workspace_bytes(bd_ctx* ctx);

/* ---- fusion: replaces model_fuse.py:model_confuse (271-350) --------------------------------- */
/* The literals of model_fuse.py / edge_3.py the library was built with (reference line in the comment). */
typedef struct bd_post_constants_t {
  int32_t fuse_min_area;      /* model_fuse.py:22   polygons of area <= 1000 are erased */
  int32_t fuse_min_fragment;  /* model_fuse.py:57   fragments of area <= 500 are erased */
  int32_t fuse_split_width;   /* model_fuse.py:180-181,67   1x5 kernel x 5 iterations = 21 */
  int32_t fuse_votes;         /* model_fuse.py:323  >= 3 of 5 */
  int32_t edge_min_area;      /* edge_3.py:326      <= 100 erased */
  int32_t edge_min_fragment;  /* edge_3.py:131      < 50 erased */
  int32_t edge_split_width;   /* edge_3.py:128      7 */
  double edge_iou;            /* edge_3.py:42       0.5 */
  double edge_min_moment;     /* edge_3.py:331      m00 <= 10 skipped */
  double tier_small, tier_mid, tier_big0, tier_big1, tier_big2;          /* edge_3.py:360-375: 150, 300, 3000, 8000, 15000 */
  double eps_default, eps_mid_mult, eps_big0, eps_big1, eps_big2;        /* 0.01, x5, 0.005, 0.004, 0.002 (x perimeter) */
} bd_post_constants_t;
int bd_post_constants(bd_ctx* ctx, bd_post_constants_t* out);

/* masks5_dev: 5 u8 masks (5,h,w), nonzero = building; fused_dev: u8 (h,w) {0,255}.
 * Per mask: hole fill + drop polygon-area <= 1000, directional 1x21 / 21x1 erosion split with fragment
 * filter <= 500; vote >= 3 of 5; same clean-up again.  Internally the masks are packed to 1 bit per pixel and
 * labelled by horizontal runs (csrc/rle.cuh).  Synchronises the stream (run counts size the scratch). */
int bd_fuse(bd_ctx* ctx, const uint8_t* masks5_dev, int h, int w, uint8_t* fused_dev, void* stream);
/* second half of bd_fuse: vote >= 3 of five ALREADY cleaned masks, then the final clean-up (model_fuse.py:315-346) */
int bd_fuse_cleaned(bd_ctx* ctx, const uint8_t* cleaned5_dev, int h, int w, uint8_t* fused_dev, void* stream);
/* the per-mask clean-up alone (fill_and_delete + eroede_dilate_process + only_plt, model_fuse.py:9-218) */
int bd_mask_cleanup(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, uint8_t* out_dev, void* stream);
/* Bit-packed masks ("planes": bit j of word wd of row y = pixel (32 wd + j, y), bd_plane_words_per_row(w) words per
 * row, bits beyond w zero) -- what the multi-GPU path ships between ranks (1/8 of the u8 bytes) and what bd_fuse works
 * on.  Rows are independent: a band of rows packs to the same words as the corresponding rows of the whole mask. */
size_t bd_plane_words_per_row(int w);
int bd_mask_pack(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, uint32_t* plane_dev, void* stream);
int bd_mask_unpack(bd_ctx* ctx, const uint32_t* plane_dev, int h, int w, uint8_t* mask_dev, void* stream);
/* bd_fuse on five planes stored back to back (cleaned != 0: they are already cleaned, like bd_fuse_cleaned); writes
 * the fused mask as u8 (fused_dev, may be NULL) and / or as a plane (fused_plane_dev, may be NULL) */
int bd_fuse_planes(bd_ctx* ctx, const uint32_t* planes5_dev, int cleaned, int h, int w, uint8_t* fused_dev,
                   uint32_t* fused_plane_dev, void* stream);
/* test hooks: an intermediate plane of one clean-up pass as a u8 mask (stage 0 holes filled, 1 area filter, 2 / 3
 * horizontal / vertical erosion, 4 objects kept whole, 5 / 6 surviving fragments of the two splits, 7 result), and the
 * run-based labelling as per-pixel labels (raster index of the component's first pixel, -1 outside the set; fg: label
 * the set or the clear pixels; conn8: 8- or 4-connected) */
int bd_debug_cleanup_stage(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, int stage, uint8_t* out_dev, void* stream);
int bd_debug_labels(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, int fg, int conn8, int32_t* labels_dev, void* stream);

/* ---- contours: replaces edge_3.py:_detection (310-387) -------------------------------------- */
typedef struct bd_polys {
  int32_t n_polys;
  int32_t n_points;        /* total points over all polygons (each polygon closed: first point repeated) */
  int32_t* offsets;        /* n_polys+1 entries into xs/ys */
  float* xs;               /* integer-valued except minAreaRect fallbacks (edge_3.py:281-285) */
  float* ys;
  uint8_t* is_float;       /* per polygon: 0 = closed integer polygon; 2 = the 4-vertex search of small_target
                              (edge_3.py:265-286) failed and xs/ys hold the RAW contour: the caller applies
                              cv::boxPoints(cv::minAreaRect(.)), whose float32 libm trigonometry stays on the host */
} bd_polys;
int bd_contours(bd_ctx* ctx, const uint8_t* mask_dev, int h, int w, bd_polys* out, void* stream);
void bd_polys_free(bd_polys* p);

/* ---- level-0 PNG of the stage hand-offs (host code: no GPU needed) ------------------------------------------- */
/* The reference writes its masks with cv.imwrite(..., [IMWRITE_PNG_COMPRESSION, 0]) (predict.py:115, model_fuse.py:350)
 * and base64-encodes the result file (buildAPI.py:122-123).  bd_png0_encode produces such a file for an (h, w) u8 mask
 * in host memory: 8-bit grey, stored deflate blocks, filter 0 -- multi-threaded copy + CRC-32 / Adler-32 with
 * checksum combination (threads <= 0: all cores, at most 32).  out_host needs bd_png0_size(h, w) bytes. */
size_t bd_png0_size(int h, int w);
int bd_png0_encode(const uint8_t* mask_host, int h, int w, uint8_t* out_host, size_t cap, size_t* out_len, int threads);

/* host-side restatements of the OpenCV primitives edge_3.py applies per contour (integer (x,y) pairs, closed
 * curves), exposed so that they can be checked against cv2 without a GPU: cv::contourArea, cv::arcLength,
 * cv::approxPolyDP (returns the vertex count, writes out_xy[2*count]) and the area-tiered choice of
 * edge_3.py:351-378 (returns 0 = skipped, 1 = polygon written, 2 = needs minAreaRect). */
double bd_host_contour_area(const int32_t* xy, int n);
double bd_host_arc_length(const int32_t* xy, int n);
int bd_host_approx_poly(const int32_t* xy, int n, double eps, int32_t* out_xy);
int bd_host_simplify(const int32_t* xy, int n, int32_t* out_xy, int* out_n);

#ifdef __cplusplus
}
#endif
#endif /* BD_B200_H */
