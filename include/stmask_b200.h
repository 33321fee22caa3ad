/*
 * stmask_b200.h — C ABI of libstmask_b200.so
 *
 * B200-native (sm_100a) implementation of STMask's feature-calibration and
 * temporal-fusion hot path.  Every entry point below replaces one third-party
 * native operator that the reference imports (the reference ships no native
 * code of its own; citations are into the reference tree):
 *
 *   stm_deform_conv2d_fwd   <- dcn_v2.DCN / dcn_v2_conv        (backbone.py:5,21-26,45)
 *                           <- mmcv.ops.DeformConv2d           (layers/modules/Featurealign.py:3,27-31,72)
 *                           <- mmcv.ops.ModulatedDeformConv2d  (named by the north star; same math as dcn_v2)
 *   stm_fcb_ali_offsets     <- the closed-form box->offset map (layers/modules/Featurealign.py:46-69)
 *   stm_fcb_ada_offsets     <- the 1x1 conv_offset on box deltas    (layers/modules/Featurealign.py:20-25,44)
 *   stm_roi_align_fwd       <- mmcv.ops.roi_align as bbox_feat_extractor calls it
 *                              (layers/modules/track_to_segment_head.py:65-88)
 *   stm_detect_fast_nms_fwd <- generate_candidate + Detect_TF.cc_fast_nms, batched and sync-free
 *                              (layers/functions/TF_utils.py:54-82, layers/functions/detection_TF.py:85-134)
 *   stm_mask_assembly_fwd   <- generate_mask + crop           (layers/mask_utils.py:111-128, layers/box_utils.py:341-364)
 *   stm_mask_iou_fwd        <- mask_iou                       (layers/box_utils.py:435-447)
 *   stm_track_update_fwd    <- the matching state machine of Track_TF.track + compute_comp_scores, for a batch of clips
 *                              (layers/functions/track_TF.py:52-181, layers/functions/TF_utils.py:98-123)
 *   stm_pool_fc_fwd         <- TemporalNet's AvgPool2d(7x7) + fc + fc_coeff tail; its three 3x3 convs are
 *                              stm_deform_conv2d_fwd with STM_DCN_ZERO_OFFSET
 *                              (layers/modules/track_to_segment_head.py:10-37)
 *   stm_correlation_fwd     <- spatial_correlation_sampler.spatial_correlation_sample
 *                              + the /C, leaky-ReLU, concat, ReLU that follow it
 *                              (layers/modules/track_to_segment_head.py:53-62,
 *                               layers/functions/TF_utils.py:28-31, STMask.py:291-297)
 *
 * Conventions
 *   - extern "C", plain-old-data only; no torch / C++ types cross this boundary.
 *   - All pointers are DEVICE pointers owned by the caller (activations, packed
 *     weights, workspace).  The library never allocates device memory, never
 *     synchronises and keeps no pointer after return; every kernel is enqueued
 *     on `stream` (a cudaStream_t passed as void*), so the calls are CUDA-graph
 *     capturable.
 *   - Activations are NHWC ("channels-last"): channel stride is 1, the other
 *     strides are given in ELEMENTS.  Offsets / masks / correlation outputs carry
 *     four explicit strides so both NCHW and NHWC buffers work without a copy.
 *   - Return value: 0 (STM_OK) or a negative StmStatus; stm_last_error() returns
 *     a thread-local message for the last failing call of the calling thread.
 *   - There is no CPU implementation behind this ABI.
 */
#ifndef STMASK_B200_H_
#define STMASK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define STM_ABI_VERSION 6

typedef enum StmStatus {
  STM_OK = 0,
  STM_ERR_INVALID_ARGUMENT = -1,   /* bad shape / stride / null pointer           */
  STM_ERR_UNSUPPORTED = -2,        /* valid request, no kernel for it             */
  STM_ERR_WORKSPACE = -3,          /* workspace too small                         */
  STM_ERR_CUDA = -4,               /* a CUDA runtime / driver call failed         */
  STM_ERR_NO_DEVICE = -5           /* no sm_100 device                            */
} StmStatus;

typedef enum StmDType {
  STM_F32 = 0,
  STM_BF16 = 1
} StmDType;

/* which kernel family to use; AUTO picks tcgen05 when the shape allows it */
typedef enum StmBackend {
  STM_BACKEND_AUTO = 0,
  STM_BACKEND_SIMT = 1,    /* CUDA-core kernels: any shape, fp32 or bf16 storage, fp32 math */
  STM_BACKEND_TCGEN05 = 2  /* tcgen05/TMEM/TMA kernels: bf16 storage, fp32 accumulate        */
} StmBackend;

/* ------------------------------------------------------------------------- */
/* Deformable convolution (DCNv1 when mask == NULL, DCNv2 otherwise)          */
/* ------------------------------------------------------------------------- */

enum {
  STM_DCN_RELU = 1,          /* y = max(y, 0) in the epilogue (Featurealign.py:72)                  */
  STM_DCN_MASK_SIGMOID = 2,  /* mask holds logits; apply sigmoid while sampling (dcn_v2.DCN.forward) */
  STM_DCN_ZERO_OFFSET = 4,   /* offset == NULL: plain convolution through the same pipeline          */
  STM_DCN_HINT_RASTER = 8,   /* tcgen05 sampling kernel: GEMM rows in raster order even where 8x8 pixel patches apply */
  /* Scheduling hints for the tcgen05 backend (results are identical; tests use them to force every CTA
   * shape through the same entry point).  Stateless: they travel with the call, there are no
   * environment variables or global knobs behind this ABI. */
  STM_DCN_HINT_ROWS128 = 16, /* 128 output pixels per CTA (one accumulator)                          */
  STM_DCN_HINT_ROWS256 = 32, /* 256 output pixels per CTA (two accumulators)                         */
  STM_DCN_HINT_NO_PAIR = 64, /* do not pair CTAs (tcgen05 cta_group::1 only)                         */
  STM_DCN_OUT_F32 = 128,     /* y is float32 [.., out_c] (strides in float elements) whatever conv->dtype:
                                the offset / mask-logit predictor of a DCN keeps its fp32 accumulators   */
  STM_DCN_HINT_DEEP_PIPE = 256, /* prefer more pipeline stages over L1 capacity                         */
  STM_DCN_HINT_TWO_CTAS = 512,  /* two 128-row CTAs with 8 producer warps each per SM                   */
  /* FCB, box-guided feature calibration (stm_deform_conv2d_fcb_fwd only): problem.offset is NOT the per-tap offset
   * tensor but the regressed box deltas [batch, 4 = (t_x, t_y, t_w, t_h), out_h, out_w] (strides off_stride_*,
   * dtype conv->offset_dtype); every tap's (dy, dx) is derived inside the sampling kernel. */
  STM_DCN_FCB_ADA = 1024,       /* offsets = 1x1 conv_offset(deltas)            (Featurealign.py:20-25,44)   */
  STM_DCN_FCB_ALI = 2048,       /* offsets = closed form of the box transform   (Featurealign.py:46-69)      */
  STM_DCN_HINT_TAP_MAJOR = 8192, /* K-block order (tap, chunk): sample records computed one tap ahead                        */
  STM_DCN_HINT_CHUNK_MAJOR = 16384, /* K-block order (chunk, tap) with every tap's sample records resident in shared memory
                                   (the default when deform_groups == 1)                                     */
  STM_DCN_HINT_NO_FUSE = 32768, /* STM_DCN_ZERO_OFFSET only: TMA kernel without the horizontal taps fused into N (tests)      */
  STM_DCN_OUT_PLANAR = 131072,  /* y is [batch, out_c, out_h, out_w] ("NCHW"): element (b, n, h, w) at b * y_stride_n + n * (out_h * y_stride_h)
                                   + h * y_stride_h + w * y_stride_w.  tcgen05 backend only.  A DCN's offset / mask predictor writes
                                   this layout so that the sampling kernel's per-tap offset loads of 32 neighbouring pixels are one
                                   128-byte line instead of 32 (they were 39 % of its L1 requests on the C = 128 stage)            */
  STM_DCN_HINT_GATHER = 4096    /* STM_DCN_ZERO_OFFSET only: keep the plain convolution on the gather main loop instead of
                                   the TMA shifted-view kernel (tests compare the two)                         */
};

/* Parameters shared by every problem of one call (one weight tensor). */
typedef struct StmDcnConv {
  int32_t in_c, out_c;             /* Cin, Cout (totals, not per group)                */
  int32_t kernel_h, kernel_w;
  int32_t stride_h, stride_w;
  int32_t pad_h, pad_w;
  int32_t dil_h, dil_w;
  int32_t groups;                  /* weight groups                                    */
  int32_t deform_groups;           /* offset groups; in_c % deform_groups == 0         */
  int32_t dtype;                   /* StmDType of x, packed weight, bias(always f32), y */
  int32_t offset_dtype;            /* StmDType of offset and mask                      */
  int32_t flags;                   /* STM_DCN_*                                        */
  int32_t backend;                 /* StmBackend                                       */
} StmDcnConv;

/* One feature map to convolve.  Several problems (e.g. the five FPN levels that
 * share the prediction-head weights, STMask.py:91-92) go into ONE launch. */
typedef struct StmDcnProblem {
  int32_t batch, in_h, in_w;
  int32_t out_h, out_w;            /* must equal floor((in + 2p - d(k-1) - 1)/s) + 1   */
  const void* x;                   /* [batch, in_h, in_w, in_c]  NHWC                   */
  int64_t x_stride_n, x_stride_h, x_stride_w;
  const void* offset;              /* logical [batch, dg*2*kh*kw, out_h, out_w]; ch 2k = dy, 2k+1 = dx */
  int64_t off_stride_n, off_stride_c, off_stride_h, off_stride_w;
  const void* mask;                /* logical [batch, dg*kh*kw, out_h, out_w] or NULL   */
  int64_t mask_stride_n, mask_stride_c, mask_stride_h, mask_stride_w;
  void* y;                         /* [batch, out_h, out_w, out_c]  NHWC                */
  int64_t y_stride_n, y_stride_h, y_stride_w;
} StmDcnProblem;

#define STM_DCN_MAX_PROBLEMS 8

/* Packed weight: [out_c][kernel_h][kernel_w][in_c/groups] ("OHWI"), conv->dtype.
 * Returns the number of BYTES the packed tensor needs. */
size_t stm_dcn_packed_weight_bytes(const StmDcnConv* conv);

/* w_oihw: contiguous [out_c, in_c/groups, kh, kw] in `src_dtype` -> w_packed. */
int stm_dcn_pack_weight(const StmDcnConv* conv, const void* w_oihw, int32_t src_dtype,
                        void* w_packed, void* stream);

/* Bytes of scratch the call below needs (may be 0). */
size_t stm_deform_conv2d_workspace(const StmDcnConv* conv, const StmDcnProblem* probs, int32_t n_probs);

/* y = conv(deform_sample(x, offset) * mask, W) + bias.  bias: float32[out_c] or NULL. */
int stm_deform_conv2d_fwd(const StmDcnConv* conv, const StmDcnProblem* probs, int32_t n_probs,
                          const void* w_packed, const float* bias,
                          void* workspace, size_t workspace_bytes, void* stream);

/* FeatureAlign's `conv_adaption(x, offset(deltas))` with the offsets computed INSIDE the sampling kernel from the box
 * deltas (conv->flags has STM_DCN_FCB_ADA or STM_DCN_FCB_ALI; see there): replaces stm_fcb_{ada,ali}_offsets + the
 * offset tensors + stm_deform_conv2d_fwd for the hot path's shapes.  fcb_weight: float32 [deform_groups*2*kh*kw][4] (the
 * 1x1 conv_offset weight) for ADA, NULL for ALI (deform_groups must be 1, odd kernel).  tcgen05 backend only:
 * returns STM_ERR_UNSUPPORTED when the shape needs the CUDA-core kernel (callers then use the two-step form). */
int stm_deform_conv2d_fcb_fwd(const StmDcnConv* conv, const StmDcnProblem* probs, int32_t n_probs,
                              const void* w_packed, const float* bias, const float* fcb_weight,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Which backend a call with these arguments would run: STM_BACKEND_SIMT / _TCGEN05,
 * or a negative StmStatus. */
int stm_deform_conv2d_backend(const StmDcnConv* conv, const StmDcnProblem* probs, int32_t n_probs);

/* Human-readable description of the kernel instantiation the call would launch on the current device
 * ("simt", or "tcgen05 rows=256 n=256 ... pair=1 ..."), written NUL-terminated into buf[len].
 * Parity tests assert it so that every benchmarked instantiation is known to have been compared
 * with the oracle. */
int stm_deform_conv2d_variant(const StmDcnConv* conv, const StmDcnProblem* probs, int32_t n_probs,
                              char* buf, size_t len);

/* FCB(ali) offsets from regressed box deltas (Featurealign.py:46-69), deform_groups = 1:
 *   shape[b, 0..3, h, w] = (t_x, t_y, t_w, t_h)
 *   offset[b, 2*(i*kw+j)  ] = 0.1*t_y*kh + (exp(0.2*t_h) - 1) * (i - kh/2)
 *   offset[b, 2*(i*kw+j)+1] = 0.1*t_x*kw + (exp(0.2*t_w) - 1) * (j - kw/2)
 * Both tensors are addressed through explicit element strides (n, c, h, w). */
int stm_fcb_ali_offsets(const void* shape, const int64_t shape_strides[4], int32_t shape_dtype,
                        void* offset, const int64_t offset_strides[4], int32_t offset_dtype,
                        int32_t batch, int32_t h, int32_t w, int32_t kernel_h, int32_t kernel_w,
                        void* stream);

/* FCB(ada) offsets: the bias-free 1x1 `conv_offset` over the 4 box-delta channels
 * (Featurealign.py:20-25,44).  weight: float32 [out_channels][4] (the conv weight viewed 2-D),
 * offset[b, o, h, w] = sum_c weight[o][c] * shape[b, c, h, w]. */
int stm_fcb_ada_offsets(const void* shape, const int64_t shape_strides[4], int32_t shape_dtype,
                        const float* weight, void* offset, const int64_t offset_strides[4],
                        int32_t offset_dtype, int32_t batch, int32_t h, int32_t w,
                        int32_t out_channels, void* stream);

/* ------------------------------------------------------------------------- */
/* Temporal-fusion correlation cost volume                                    */
/* ------------------------------------------------------------------------- */

enum {
  STM_CORR_LEAKY_RELU = 1,   /* v = v > 0 ? v : slope * v   (track_to_segment_head.py:62) */
  STM_CORR_RELU = 2,         /* v = max(v, 0)               (TF_utils.py:31)              */
  STM_CORR_COPY_FEATS = 4    /* also write relu?(feat_ref), relu?(feat_next) behind the P*P
                                correlation channels: the 633-channel concat of TF_utils.py:30 */
};

typedef struct StmCorrDesc {
  int32_t batch, h, w, c;          /* x1, x2: [batch, h, w, c] NHWC                    */
  int32_t patch;                   /* patch_size P (odd); output has P*P channels      */
  int32_t dilation_patch;
  int32_t dtype;                   /* StmDType of x1, x2                               */
  int32_t out_dtype;               /* StmDType of out                                  */
  int32_t flags;                   /* STM_CORR_*                                       */
  int32_t backend;                 /* StmBackend                                       */
  float scale;                     /* multiplies the raw dot product (1/C in correlate) */
  float leaky_slope;               /* 0.1 in correlate                                 */
  int64_t x1_stride_n, x1_stride_h, x1_stride_w;
  int64_t x2_stride_n, x2_stride_h, x2_stride_w;
  /* out[b, k, y, x], k = ph*P + pw  <->  displacement ((ph-P/2)*d, (pw-P/2)*d);
   * with STM_CORR_COPY_FEATS channels P*P.. hold feat_a then feat_b. */
  int64_t out_stride_n, out_stride_c, out_stride_h, out_stride_w;
  /* STM_CORR_COPY_FEATS only: two NHWC tensors with feat_c channels each */
  int32_t feat_c;
  int32_t feat_dtype;
  int64_t feat_a_stride_n, feat_a_stride_h, feat_a_stride_w;
  int64_t feat_b_stride_n, feat_b_stride_h, feat_b_stride_w;
  /* STM_CORR_COPY_FEATS only: channel at which feat_a starts; 0 means P*P (the reference's 633-channel
   * concat).  A larger value pads the correlation block — channels [P*P, feat_c_offset) are written as
   * zeros — e.g. 128 for P = 11, which makes every block of a channels-last bf16 pixel row 16-byte aligned
   * (the layout the fused temporal-fusion path uses: [corr 121 | 0 x 7 | T2S_ref 256 | T2S_next 256]). */
  int32_t feat_c_offset;
  int32_t reserved_;
  /* Optional pair indexing — temporal fusion straight out of a frame batch (no gathered copies of the
   * (t-1, t) pairs, STMask.py:289-297 generalised): when x1_index / x2_index are non-NULL DEVICE arrays of
   * `batch` int32, pair i correlates frame x1_index[i] of x1 with frame x2_index[i] of x2 and, with
   * STM_CORR_COPY_FEATS, copies feat_a[x1_index[i]] and feat_b[x2_index[i]].  x1 / feat_a hold x1_frames
   * frames, x2 / feat_b hold x2_frames.  An x1_index value v >= x1_frames selects frame v - x1_frames of the
   * ALTERNATE tensors x1_alt / feat_a_alt: the one-frame halos received from the neighbour rank.
   * tcgen05 backend only (STM_ERR_UNSUPPORTED otherwise). */
  const int32_t* x1_index;
  const int32_t* x2_index;
  int32_t x1_frames, x2_frames, alt_frames, reserved2_;
  const void* x1_alt;
  const void* feat_a_alt;
  int64_t x1_alt_stride_n, x1_alt_stride_h, x1_alt_stride_w;
  int64_t feat_a_alt_stride_n, feat_a_alt_stride_h, feat_a_alt_stride_w;
} StmCorrDesc;

/* out[b,ph,pw,y,x] = post( scale * sum_c x1[b,y,x,c] * x2[b, y+(ph-r)d, x+(pw-r)d, c] ), zero outside x2. */
int stm_correlation_fwd(const StmCorrDesc* desc, const void* x1, const void* x2,
                        const void* feat_a, const void* feat_b, void* out, void* stream);

int stm_correlation_backend(const StmCorrDesc* desc);

/* The same operator over n <= 8 feature maps in ONE launch (the FPN levels P3..P7 of the operator sweep): descs[i],
 * x1s[i], x2s[i], outs[i] describe feature map i; C, patch, dilation, dtypes, flags and scale must be equal across
 * them.  bf16 / tcgen05 only, no concat features, no pair indexing.  A desc with feat_c_offset > patch^2 (and no
 * STM_CORR_COPY_FEATS) asks for zero-padded channels-last rows of feat_c_offset channels (128 for patch 11: 256-byte
 * rows, written with 16-byte stores); channels [patch^2, feat_c_offset) are zeros. */
int stm_correlation_multi_fwd(const StmCorrDesc* descs, const void* const* x1s, const void* const* x2s,
                              void* const* outs, int32_t n, void* stream);

/* ------------------------------------------------------------------------- */
/* RoIAlign (average pooling) on an NHWC feature map                          */
/* ------------------------------------------------------------------------- */
typedef struct StmRoiAlignDesc {
  int32_t batch, h, w, c;          /* feat: [batch, h, w, c] NHWC (channel stride 1)       */
  int32_t n_rois;
  int32_t pooled_h, pooled_w;      /* 7 x 7 in the reference                              */
  int32_t sampling_ratio;          /* <= 0: adaptive grid ceil(roi_size / pooled_size)    */
  int32_t aligned;                 /* 1: shift the box by -0.5 pixel (mmcv aligned=True)  */
  int32_t dtype, out_dtype;        /* StmDType of feat / out                              */
  float spatial_scale;
  int64_t feat_stride_n, feat_stride_h, feat_stride_w;
  int64_t out_stride_n, out_stride_c, out_stride_h, out_stride_w;   /* out: logical [n_rois, c, pooled_h, pooled_w] */
} StmRoiAlignDesc;

/* rois: float32 DEVICE array [n_rois][5] = (batch index, x1, y1, x2, y2), contiguous.
 * out[r, c, i, j] = mean over the sample grid of bin (i, j) of the bilinearly interpolated feature;
 * sample points outside [-1, h] x [-1, w] contribute 0, others are clamped into the map. */
int stm_roi_align_fwd(const StmRoiAlignDesc* desc, const void* feat, const float* rois, void* out, void* stream);

/* ------------------------------------------------------------------------- */
/* Candidate generation + cross-class fast NMS for a batch of frames           */
/* (generate_candidate, TF_utils.py:54-82; Detect_TF.cc_fast_nms,              */
/*  detection_TF.py:85-134) — one launch, no host round trip                   */
/* ------------------------------------------------------------------------- */
/* conf [frames, n_priors, n_classes] class PROBABILITIES (class 0 = background), loc [frames, n_priors, 4] box
 * regressions, centerness [frames, n_priors] or NULL, priors [n_priors, 4] (cx, cy, w, h) — all float32, contiguous.
 * A prior is a candidate when max_{c>0} conf > conf_thresh; score = that probability x centerness; candidates are
 * sorted by score (descending), cut to top_k (<= 256) and a candidate survives iff no higher-scoring candidate overlaps
 * it with IoU > nms_thresh.  Outputs (device, fixed size, survivors in score order): count [frames],
 * index / cls / score [frames, top_k], box [frames, top_k, 4] (x1, y1, x2, y2; decoded with variances 0.1 / 0.2,
 * box_utils.py:238-283).  Entries past count[f] are left untouched. */
int stm_detect_fast_nms_fwd(const float* conf, const float* loc, const float* centerness, const float* priors,
                            int32_t frames, int32_t n_priors, int32_t n_classes, int32_t top_k,
                            float conf_thresh, float nms_thresh,
                            int32_t* count, int32_t* index, int32_t* cls, float* score, float* box, void* stream);

/* ------------------------------------------------------------------------- */
/* Mask assembly and mask IoU for a batch of frames                           */
/* (generate_mask, layers/mask_utils.py:111-128; crop, box_utils.py:341-364;  */
/*  mask_iou, box_utils.py:435-447; callers track_TF.py:76-112)               */
/* ------------------------------------------------------------------------- */
/* proto [frames, h, w, k] prototypes, coeff [frames, max_n, k] raw mask coefficients (tanh is applied here),
 * boxes [frames, max_n, 4] relative x1, y1, x2, y2, count [frames] valid detections per frame (NULL: max_n) — float32 /
 * int32, contiguous.  masks [frames, max_n, h, w] = sigmoid(proto . tanh(coeff)) inside the box grown by one pixel
 * (the reference's crop), 0 outside; mask_bits [frames, max_n, ceil(h*w/32)] = masks > 0.5, one bit per pixel.
 * Rows past count[f] are left untouched. */
int stm_mask_assembly_fwd(const float* proto, const float* coeff, const float* boxes, const int32_t* count,
                          float* masks, uint32_t* mask_bits, int32_t frames, int32_t h, int32_t w, int32_t k,
                          int32_t max_n, void* stream);

/* iou[f, i, j] = |A_i and B_j| / |A_i or B_j| (0 for an empty union) on the bit masks above:
 * bits_a [frames, max_a, words], bits_b [frames, max_b, words], count_a / count_b [frames] or NULL, iou [frames, max_a, max_b]. */
int stm_mask_iou_fwd(const uint32_t* bits_a, const uint32_t* bits_b, const int32_t* count_a, const int32_t* count_b,
                     float* iou, int32_t frames, int32_t max_a, int32_t max_b, int32_t words, void* stream);

/* ------------------------------------------------------------------------- */
/* Tracker state machine for a batch of independent clips                     */
/* (Track_TF.track, layers/functions/track_TF.py:52-181; compute_comp_scores, */
/*  layers/functions/TF_utils.py:98-123) — one launch per frame, no host sync */
/* ------------------------------------------------------------------------- */
/* The reference's `prev_candidate` dict of growing tensors as fixed-capacity DEVICE arrays owned by the caller
 * (rows past n_obj[clip] are unused).  `mask` / `centerness` may be NULL. */
typedef struct StmTrackState {
  int32_t* n_obj;        /* [clips]            objects tracked so far in each clip (0 = `prev_candidate is None`) */
  float* box;            /* [clips, cap, 4]    x1, y1, x2, y2 (relative)                                          */
  float* score;          /* [clips, cap]                                                                          */
  int32_t* cls;          /* [clips, cap]                                                                          */
  float* coeff;          /* [clips, cap, k]    mask coefficients                                                  */
  float* track;          /* [clips, cap, e]    L2-normalised track embeddings                                     */
  float* centerness;     /* [clips, cap]                                                                          */
  int32_t* tracked;      /* [clips, cap]       `tracked_mask`: frames since the object was last matched           */
  uint32_t* mask_bits;   /* [clips, cap, words] masks > 0.5 as bit planes (stm_mask_assembly_fwd)                  */
  float* mask;           /* [clips, cap, hw]   soft masks                                                         */
} StmTrackState;

/* This frame's detections after fast NMS, per clip (stm_detect_fast_nms_fwd / stm_mask_assembly_fwd outputs). */
typedef struct StmTrackDets {
  const int32_t* count;      /* [clips] valid rows, or NULL: max_det                */
  const float* box;          /* [clips, max_det, 4]                                 */
  const float* score;        /* [clips, max_det]                                    */
  const int32_t* cls;        /* [clips, max_det]                                    */
  const float* coeff;        /* [clips, max_det, k]                                 */
  const float* track;        /* [clips, max_det, e]                                 */
  const float* centerness;   /* [clips, max_det] or NULL                            */
  const uint32_t* mask_bits; /* [clips, max_det, words]                             */
  const float* mask;         /* [clips, max_det, hw] or NULL                        */
} StmTrackDets;

typedef struct StmTrackParams {
  int32_t clips, cap, max_det;   /* cap, max_det <= 256                                                              */
  int32_t k, e, words, hw;       /* mask coefficients, embedding size, words per bit plane, pixels per soft mask     */
  int32_t max_age;               /* an object is reported while tracked <= max_age (10, track_TF.py:160)             */
  float match_coeff[4];          /* cfg.match_coeff: weights of score, mask IoU, box IoU, same label                 */
  float bbox_dummy_iou;          /* IoU of the "new object" column (0.3, track_TF.py:126)                            */
  float conf_thresh;             /* cfg.eval_conf_thresh (track_TF.py:164)                                           */
} StmTrackParams;

/* One frame of every clip: ages the tracked objects, matches the detections (mask_iou [clips, max_det, cap] =
 * stm_mask_iou_fwd(det bits, state bits) after the caller's CandidateShift), runs the reference's sequential assignment
 * in detection order and copies the winning detections' rows into the state.  is_first [clips] (uint8, or NULL) resets a
 * clip's state (a new video).  Outputs: det_slot [clips, max_det] the state row each detection went to (-1: it lost its
 * object to a higher-scoring detection, or the state is full), keep [clips, cap] the output filter of track_TF.py:158-165;
 * an object's id is its row index (box_ids = arange, track_TF.py:158). */
int stm_track_update_fwd(const StmTrackParams* params, const StmTrackState* state, const StmTrackDets* dets,
                         const float* mask_iou, const uint8_t* is_first, int32_t* det_slot, uint8_t* keep, void* stream);

/* ------------------------------------------------------------------------- */
/* TemporalNet tail: y[n, :] = W * mean over the hw pixels of x[n] + b        */
/* (AvgPool2d(7x7) + fc + fc_coeff, track_to_segment_head.py:17-19,31-35)     */
/* ------------------------------------------------------------------------- */
/* x: NHWC activations [n, hw, c] (`dtype`), pixel p of box i at x + i*x_stride_n + p*x_stride_p (elements);
 * weight: float32 [out_features][c] (fc rows, then fc_coeff rows), bias: float32[out_features] or NULL;
 * y: float32 [n][out_features]. */
int stm_pool_fc_fwd(const void* x, int32_t dtype, int32_t n, int32_t hw, int32_t c,
                    int64_t x_stride_n, int64_t x_stride_p, const float* weight, const float* bias,
                    int32_t out_features, float* y, void* stream);

/* ------------------------------------------------------------------------- */
/* Layout helpers (NCHW <-> NHWC with dtype conversion), used at the module   */
/* boundary when a caller hands over contiguous NCHW tensors.                 */
/* ------------------------------------------------------------------------- */
int stm_nchw_to_nhwc(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype,
                     int32_t n, int32_t c, int32_t h, int32_t w, void* stream);
int stm_nhwc_to_nchw(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype,
                     int32_t n, int32_t c, int32_t h, int32_t w, void* stream);

/* ------------------------------------------------------------------------- */
int stm_version(void);               /* STM_ABI_VERSION of the loaded library            */
const char* stm_last_error(void);    /* thread-local; "" when the last call succeeded    */
int stm_device_supported(int32_t device); /* 1 if `device` is compute capability 10.x       */
/* number of kernels this library has launched from the calling process (bench accounting) */
uint64_t stm_kernel_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* STMASK_B200_H_ */
