/* kpd_b200.h -- C-ABI of the B200-native keypoint-detection inference hot path.
 *
 * Drop-in boundary: the reference has no FFI layer; its hot path sits behind the Python
 * classes of `dll.models`.  This library is the native layer the re-implemented `dll` package calls
 * through ctypes.  Each entry point names the reference code it replaces (paths relative to the
 * upstream repository root; `tv:` = torchvision 0.26).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All `d_*` pointers are DEVICE pointers owned by
 *     the caller; `h_*` are HOST pointers.  `stream` is a cudaStream_t passed as void*.
 *   - no allocation, no synchronisation and no stream creation inside forward calls: work is enqueued
 *     on the caller's stream and scratch memory comes from the caller-supplied workspace.
 *   - return 0 on success, <0 = kpd_status argument/state error, >0 = cudaError_t.
 *     kpd_last_error() returns a thread-local description of the last failure.
 *   - sm_100a only.  There is no CPU fallback and no second backend.
 */
#ifndef KPD_B200_H
#define KPD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPD_ABI_VERSION 1
#define KPD_NUM_STATE_TENSORS 318   /* dll/models/keypoint_model.py:55-63 state_dict, reference order */
#define KPD_HEATMAP 56              /* dll/models/keypoint_model.py:213 output_size=(56,56) */
#define KPD_TOPK 64                 /* dll/models/keypoint_model.py:90  k=64 */
#define KPD_FPN_CH 128              /* dll/models/backbone.py:255 */

typedef enum {
    KPD_OK = 0,
    KPD_ERR_ARG = -1,        /* bad argument (null pointer, unsupported size) */
    KPD_ERR_STATE = -2,      /* plan not usable on this device / wrong ABI */
    KPD_ERR_WORKSPACE = -3,  /* workspace too small */
    KPD_ERR_UNSUPPORTED = -4 /* configuration outside the hot path (e.g. device is not sm_100) */
} kpd_status;

/* arithmetic / storage policy of the dense (tensor-core) convolutions */
typedef enum {
    KPD_PREC_TF32 = 0,   /* "fp32" mode: fp32 storage, tcgen05 kind::tf32, fp32 accumulate            */
    KPD_PREC_BF16 = 1,   /* bf16 head (bf16 storage + kind::f16), tf32 trunk (FPN-0), fp32 score path */
    KPD_PREC_FP32_SIMT = 2 /* strict fp32 FMA on CUDA cores: parity debugging only (slow)             */
} kpd_precision;

typedef struct kpd_plan kpd_plan;    /* packed weights on one device; immutable after creation */

typedef struct {
    int32_t abi_version;     /* KPD_ABI_VERSION */
    int32_t in_channels;     /* 1 or 3  (dll/configs/model_config.py:11) */
    int32_t num_keypoints;   /* 17 */
    int32_t reserved;
} kpd_model_desc;

/* Byte offsets of every stage tensor inside the caller's workspace (for stage-wise parity tests and
 * for hosts that want intermediate results).  All activations are NHWC. */
typedef struct {
    size_t total_bytes;
    int32_t batch, size, max_rois, precision;   /* what the layout was computed for */
    size_t feat[4];          /* fp32 [B,S/2,S/2,16] [B,S/8,S/8,24] [B,S/16,S/16,48] [B,S/32,S/32,576] */
    size_t lateral0;         /* fp32 [B,S/2,S/2,128] top-down sum feeding fpn_convs[0]; written only in strict-fp32 mode (the default modes fold it into the FPN-0 conv) */
    size_t fpn0;             /* fp32 [B,S/2,S/2,128]                                                  */
    size_t pool_avg;         /* fp32 [B,128]  */
    size_t pool_max;         /* fp32 [B,128]  */
    size_t scores;           /* fp32 [B,128]  */
    size_t topk;             /* int32 [B,64] channel indices in descending-score order */
    size_t roi_count;        /* int32 [1]   number of non-zero boxes N                               */
    size_t roi_box;          /* fp32 [Nmax,4] cx,cy,w,h of compacted boxes                            */
    size_t roi_batch;        /* int32 [Nmax]                                                          */
    size_t roi_slot;         /* int32 [Nmax] output slot b*P+rank                                     */
    size_t img_count;        /* int32 [B] valid boxes per image                                       */
    size_t roi_feat;         /* fp32 [Nmax,56,56,64] ROI-aligned selected channels                    */
    size_t head_cw;          /* fp32 [Nmax,64]   channel attention weights                            */
    size_t head_sw;          /* fp32 [Nmax,56,56] spatial attention weights                           */
    size_t head_in;          /* elt  [Nmax,56,56,64]  attention-scaled input of deconv_layers[0]      */
    size_t head_a;           /* elt  [Nmax,56,56,256]                                                 */
    size_t head_b;           /* elt  [Nmax,56,56,256]                                                 */
    size_t head_c;           /* elt  [Nmax,56,56,64]                                                  */
    size_t scratch;          /* kernel-private partial buffers                                        */
    size_t elt_bytes;        /* 2 (bf16) or 4 (fp32/tf32) for the `elt` tensors                       */
} kpd_workspace_layout;

/* ---- lifecycle -------------------------------------------------------------------------------
 * kpd_plan_create: `h_state[i]` is a host pointer to the i-th tensor of the reference state_dict in
 * the reference's own order (float32, or int64 for num_batches_tracked / ignored), i.e. what
 * `scripts/predict.py:47-57` passes to load_state_dict.  BatchNorm is folded, weights are repacked
 * (NHWC / K-major, bf16 + tf32 copies) and uploaded to `device`.  Replaces nn.Module construction +
 * load_state_dict for the hot path (keypoint_model.py:49-71). */
int  kpd_plan_create(kpd_plan** out, const kpd_model_desc* desc, const void* const* h_state,
                     int n_state, int device);
void kpd_plan_destroy(kpd_plan* plan);
int  kpd_plan_workspace_layout(const kpd_plan* plan, int batch, int size, int max_rois, int precision,
                               kpd_workspace_layout* out);

/* ---- whole forward: keypoint_model.py:73-206 with bboxes supplied -------------------------------
 * d_images  fp32 [B,Cin,S,S] (NCHW as the reference receives it)
 * d_boxes   fp32 [B,P,4] normalised cx,cy,w,h; all-zero rows are padding (keypoint_model.py:149)
 * d_topk_override  optional int32 [B,64] (tests: teacher-forced channel order), else NULL
 * outputs (padded to P persons, zero-filled; image without a valid box -> dummy person vis=[1,0,0]):
 *   d_heatmap fp32 [B,P,17,56,56], d_keypoints fp32 [B,P,1,17,2], d_visibilities fp32 [B,P,1,17,3] */
int  kpd_forward(const kpd_plan* plan, int precision, const float* d_images, int batch, int size,
                 const float* d_boxes, int persons, const int32_t* d_topk_override,
                 float* d_heatmap, float* d_keypoints, float* d_visibilities,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- forward pieces (stage-wise entry points; `layout` is what kpd_plan_workspace_layout returned for the
 *      (batch, size, max_rois, precision) the workspace was sized for) ---------------------------------------
 * backbone.py:258-264 (body + FPN level 0) + keypoint_model.py:37,40 global avg/max pools */
int  kpd_backbone_fpn0(const kpd_plan* plan, int precision, const float* d_images, int batch, int size,
                       void* d_workspace, const kpd_workspace_layout* layout, void* stream);
/* backbone.py:39 dead levels 1..3, computed only when MobileNetV3Wrapper is called standalone.
 * d_out[i] fp32 NHWC [B,h_i,w_i,128] for i=1..3 (entries may be NULL). Requires kpd_backbone_fpn0 first. */
int  kpd_fpn_levels123(const kpd_plan* plan, int batch, int size, void* d_workspace,
                       const kpd_workspace_layout* layout, float* const* d_out, void* stream);
/* keypoint_model.py:27-44 MLP + sigmoid, :653-661 topk(64) sorted descending (ties: lower channel first) */
int  kpd_channel_topk(const kpd_plan* plan, int batch, void* d_workspace, const kpd_workspace_layout* layout,
                      void* stream);
/* keypoint_model.py:33-44 ChannelAttention.forward and :653-661 select_top_k_channels as a leaf op on a caller's tensor:
 * d_feat_nchw fp32 [B,128,H,W] (hw = H*W) -> d_scores fp32 [B,128], d_topk int32 [B,64] (descending score, ties: lower
 * channel first), d_selected fp32 NCHW [B,64,H,W] = the selected planes in rank order.  Any output may be NULL.
 * d_scratch: B*512 floats. */
int  kpd_channel_attention(const kpd_plan* plan, const float* d_feat_nchw, int batch, int hw, float* d_scores,
                           int32_t* d_topk, float* d_selected, void* d_scratch, void* stream);
/* keypoint_model.py:137-206: box compaction, roi_align (:212-228), HeatmapHead.forward
 * (heatmap_head.py:81-113), decode_heatmap (:250-313), convert_to_original_coords (:230-248), padding.
 * The workspace must have been laid out for this (batch, size, precision) with max_rois >= batch*persons: every stage
 * entry point checks its arguments against `layout` and returns KPD_ERR_ARG on a mismatch.  Channel indices in
 * d_topk_override are clamped to [0,128) on the device. */
int  kpd_roi_head(const kpd_plan* plan, int precision, int batch, int size, const float* d_boxes, int persons,
                  const int32_t* d_topk_override, float* d_heatmap, float* d_keypoints, float* d_visibilities,
                  void* d_workspace, const kpd_workspace_layout* layout, void* stream);
/* keypoint_model.py:212-228 extract_roi_features (tv roi_align, aligned=False, adaptive sampling) as a leaf op:
 * d_feat fp32 NHWC [B,H,W,128], d_channels int32 [B,64] (channels to gather per image), d_boxes fp32 [n,4] normalised
 * cx,cy,w,h, d_roi_batch int32 [n] -> d_out fp32 NHWC [n,56,56,64].  d_partials: scratch of n*28*2*64 floats. */
int  kpd_roi_align(const float* d_feat, const int32_t* d_channels, const float* d_boxes, const int32_t* d_roi_batch,
                   int n, int height, int width, float* d_out, float* d_partials, void* stream);
/* heatmap_head.py:81-113 on caller-supplied ROI features: d_x fp32 NCHW [n,64,56,56] ->
 * d_heatmaps fp32 NCHW [n,17,56,56], d_cw fp32 [n,64], d_sw fp32 [n,56,56] (either may be NULL). */
int  kpd_heatmap_head(const kpd_plan* plan, int precision, const float* d_x, int n, float* d_heatmaps,
                      float* d_cw, float* d_sw, void* d_workspace, const kpd_workspace_layout* layout, void* stream);

/* Leaf access to the dense 3x3 convolutions (fpn_convs[0]: conv 0; head convs: 1..3), bias + ReLU fused.
 * impl 0: fp32 CUDA cores, 1: tcgen05 tf32, 2: tcgen05 bf16 (all NHWC), 3 / 4: tap-reuse tcgen05 tf32 / bf16 with the
 * input in chunk-major layout [n][C/E][h][w][E] (E = 16 bytes of channels) and NHWC output. */
int  kpd_dense_conv(const kpd_plan* plan, int conv, int impl, const void* d_in, void* d_out, int n, int h, int w, void* stream);

/* ---- standalone decoders: heatmap_head.py:265-413 and keypoint_model.py:250-313 ------------------
 * d_heatmaps [N,K,H,W], dtype 0 = fp32, 1 = bf16.  Any output pointer may be NULL.
 *   d_argmax int32 [N,K] flat index of the first maximum;  d_maxval fp32 [N,K]
 *   d_kp_plain/d_kp_subpix/d_kp_soft fp32 [N,K,2] (x/(W-1), y/(H-1));  d_score_subpix fp32 [N,K]
 *   d_vis fp32 [N,K,3] one-hot of sigmoid(max) against 0.3 / 0.7 */
int  kpd_decode(const void* d_heatmaps, int dtype, int64_t n, int k, int h, int w, float temperature,
                int32_t* d_argmax, float* d_maxval, float* d_kp_plain, float* d_kp_subpix,
                float* d_score_subpix, float* d_kp_soft, float* d_vis, void* stream);

/* ---- person head: person_head.py:27-34 1x1 heads, :54-94 box_iou, :96-139 greedy NMS -------------- */
int  kpd_person_heads(const kpd_plan* plan, int level, const float* d_feat_nhwc, int64_t pixels,
                      float* d_box_out /*[pixels,36]*/, float* d_cls_out /*[pixels,9]*/, void* stream);
int  kpd_box_iou(const float* d_boxes1, int n1, const float* d_boxes2, int n2, float* d_iou, void* stream);
/* d_boxes [B,M,4] cxcywh, d_scores [B,M]; d_keep int32 [B,M] (kept indices in descending-score order,
 * -1 padded), d_keep_count int32 [B].  M <= 1024.  max_out <= 0 means unlimited. */
int  kpd_nms(const float* d_boxes, const float* d_scores, int batch, int m, float iou_thr, int max_out,
             int32_t* d_keep, int32_t* d_keep_count, void* stream);

/* NEW BEHAVIOUR, no reference semantics (the reference's bare-tensor forward crashes: keypoint_model.py:114-119 ->
 * person_head.py:141-166 returns raw head outputs): person boxes from the level-0 heads.
 * d_box [B,H*W,36] / d_cls [B,H*W,9] = kpd_person_heads outputs on the FPN-0 map of `size` x `size` images.  Per image:
 * 9 anchors per pixel in the order of person_head.py:45-51 (sizes anchor_sizes3[0..2] x ratios 0.5,1,2; normalised by `size`),
 * standard delta decode, score = sigmoid(logit); the top_m (<= 128) candidates by score, those >= conf_threshold, greedy NMS
 * (kpd_nms) at iou_threshold; the first max_persons survivors in descending score order -> d_boxes_out [B,max_persons,4]
 * normalised cx,cy,w,h, zero-padded (all-zero box = padding, keypoint_model.py:149); d_count int32 [B] (may be NULL).
 * d_scratch: kpd_detect_scratch_bytes(batch, top_m) bytes, 16-byte aligned. */
int  kpd_detect_persons(const float* d_box, const float* d_cls, int batch, int height, int width, int size,
                        const float* anchor_sizes3, float conf_threshold, float iou_threshold, int top_m, int max_persons,
                        float* d_boxes_out, int32_t* d_count, void* d_scratch, void* stream);
size_t kpd_detect_scratch_bytes(int batch, int top_m);

/* ---- image preparation -------------------------------------------------------------- */
/* The reference's `ITransform` (dll/data/transforms.py:9-114: OpenCV + Pillow + torchvision on the host, one image at a time)
 * for a batch of equally sized uint8 images on the device, bit-exact with those libraries (OpenCV 4.13, Pillow 12.2).
 * d_src [batch,height,width,channels] uint8, channels 1 (gray), 3 (RGB) or 4 (RGBA, alpha ignored).
 * grayscale != 0 (transforms.py:43-78): RGB2GRAY -> CLAHE(clip_limit, tiles) -> {GaussianBlur 5x5 s1.5 -> medianBlur 5 ->
 *   Canny(100,200) -> dilate,erode,dilate,erode 3x3 -> GaussianBlur 3x3 -> rescale to the maximum} -> addWeighted(0.7, 0.3)
 *   -> Image.resize(out_size, BILINEAR) -> /255 -> (x-0.5)/0.5; d_out float32 [batch,1,out_size,out_size].
 * grayscale == 0 (transforms.py:80-109): per-channel CLAHE -> GaussianBlur 3x3 s0.5 -> resize -> ImageNet mean/std;
 *   d_out float32 [batch,3,out_size,out_size].
 * d_scratch: kpd_itransform_scratch_bytes(...) bytes, 256-byte aligned; no allocation and no host synchronisation inside.
 * Limits per call: images 16..16384 pixels per side, batch x planes <= 65535, batch x planes x height x width < 2^31.
 * kpd_itransform_stage_offsets: byte offsets of the stage planes inside the scratch buffer (tests / visualisation):
 *   [0] gray or colour planes u8 [B,planes,H,W], [1] CLAHE LUTs u8 [B,planes,tiles,256], [2] CLAHE (rgb: blurred) planes,
 *   [3] Canny map after hysteresis u8 (2 = edge), [4] blurred edge plane before the rescale, [5] horizontal-pass rows [B,planes,H,S]. */
size_t kpd_itransform_scratch_bytes(int batch, int height, int width, int grayscale, int tiles_x, int tiles_y, int out_size);
int  kpd_itransform_stage_offsets(int batch, int height, int width, int grayscale, int tiles_x, int tiles_y, int out_size,
                                  size_t* offsets /*[6]*/);
int  kpd_itransform(const uint8_t* d_src, int batch, int height, int width, int channels, int grayscale, double clip_limit,
                    int tiles_x, int tiles_y, int out_size, float* d_out, void* d_scratch, void* stream);

/* ---- generic fp32 layers ------------------------------------------------------------- */
/* Shape-generic CUDA-core kernels behind the classes the reference exports but never instantiates (`BACKBONE`,
 * dll/models/backbone.py:41-245; `KEYPOINT_HEAD`, dll/models/keypoint_head.py:9-90).  NHWC float32, exact fp32 FMA, fixed
 * summation orders.  Activation codes: 0 none, 1 ReLU, 2 Hardswish, 3 Hardsigmoid, 4 sigmoid, 5 ReLU6.
 * kpd_conv2d: d_out[n][oy][ox][co] = act2( act( conv(d_in * d_in_scale[n][ci]) + d_bias ) + d_residual ); d_w [k][k][cin/groups][cout];
 *   d_bias, d_in_scale [n][cin], d_residual (shape of d_out) may be NULL.  BatchNorm is folded into d_w / d_bias by the caller.
 * kpd_adaptive_avgpool: torch adaptive_avg_pool2d bins.  kpd_linear: d_y[m][n] = act(d_x[m][:] . d_w[n][:] + d_b[n]).
 * kpd_layernorm: over the last dimension, biased variance.  kpd_mul_map: d_out[p][c] = d_x[p][c] * d_map[p].
 * kpd_upsample_add: d_dst += nearest resize of d_src (F.interpolate(mode='nearest', size=...), backbone.py:35-37). */
int  kpd_conv2d(const float* d_in, const float* d_w, const float* d_bias, const float* d_in_scale, const float* d_residual,
                float* d_out, int n, int h, int w, int cin, int cout, int k, int stride, int pad, int groups, int act, int act2,
                void* stream);
int  kpd_adaptive_avgpool(const float* d_in, float* d_out, int n, int h, int w, int c, int oh, int ow, void* stream);
int  kpd_linear(const float* d_x, const float* d_w, const float* d_b, float* d_y, int m, int k, int n, int act, void* stream);
int  kpd_layernorm(const float* d_x, const float* d_gamma, const float* d_beta, float eps, float* d_y, int m, int n, int act, void* stream);
int  kpd_mul_map(const float* d_x, const float* d_map, float* d_out, int64_t pixels, int c, void* stream);
int  kpd_upsample_add(float* d_dst, const float* d_src, int n, int h, int w, int hs, int ws, int c, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------- */
const char* kpd_last_error(void);
int  kpd_abi_version(void);
/* number of kernel launches enqueued by this thread since the last call with reset!=0 (bench claim) */
int64_t kpd_launch_count(int reset);
/* CUDA-event timing of tagged regions inside the library (bench roofline): enable, then read back.
   on = 1: one event pair per stage ("body", "fpn0_conv", "head_conv2", ...); on = 2: also the nested "stage.part" tags
   (their events sit between short launches and inflate the enclosing stage's total); on = 0: off */
int  kpd_profile_enable(int on);
int  kpd_profile_read(const char* tag, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* KPD_B200_H */
