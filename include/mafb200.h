/* mafb200.h — C ABI of libmafb200.so: the B200 (sm_100a) forward/detect hot path of MAF-YOLO.
 *
 * Every entry point replaces one group of PyTorch library calls on the reference's
 * inference path (citations are into the reference tree, yang-0201/MAF-YOLO):
 *
 *   mafb200_stem_conv3x3s2   RepVGGBlock layer 0 in deploy form           yolov6/layers/common.py:206,214-217
 *   mafb200_conv3x3s2        RepVGGBlock / ConvWrapper 3x3 stride 2        common.py:76-83,166-283,776-792
 *   mafb200_conv1x1          Conv(k=1) incl. the Concat that feeds it      common.py:29-50,148-154,938-946
 *   mafb200_dwconv           UniRepLKNetBlock / DilatedReparamBlock deploy common.py:2948-3100
 *   mafb200_bottleneck       DepthBottleneckUni, all three convs (K4)       common.py:898-927
 *   mafb200_maxpool2x2       MP inside MPRep                               common.py:667-673,787-792
 *   mafb200_sppf_pool        SPPF's three chained 5x5 max-pools            common.py:114-129
 *   mafb200_upsample2x       nn.Upsample(None, 2, 'nearest')               configs/yaml/MAF-YOLO-n.yaml:21,26
 *   mafb200_head_pred        Head_DepthUni cls_pred / reg_pred + sigmoid + common.py:1325-1336, yolov6/models/yolo.py:355-396
 *                            DFL decode in the GEMM epilogue (K7)
 *   mafb200_head_decode      Head_DepthUni sigmoid + Detect_yaml eval      common.py:1332, yolov6/models/yolo.py:355-396,
 *                            branch + generate_anchors + dist2bbox         yolov6/assigners/anchor_generator.py:11-25,
 *                                                                          yolov6/utils/general.py:29-40
 *   mafb200_letterbox_u8     letterbox + Inferer.precess_image             yolov6/data/data_augment.py:53-83,
 *                                                                          yolov6/core/inferer.py:168-178
 *   mafb200_nms              non_max_suppression + torchvision.ops.nms     yolov6/utils/nms.py:31-105
 *   mafb200_scale_detections Evaler.scale_coords / box_convert /           yolov6/core/evaler.py:382-434,
 *                            Inferer.rescale                               yolov6/core/inferer.py:181-195
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - all buffers are caller-owned DEVICE memory (except where "host" is stated);
 *     the library allocates nothing and never synchronises: work is enqueued on
 *     `stream` (a cudaStream_t passed as void*), so every call is CUDA-graph capturable.
 *   - return value: 0 on success, negative MAF_E_* code on error; the text of the last
 *     error of the calling thread is available from mafb200_last_error().  Nothing throws.
 *   - there is NO CPU fallback: on a machine without an sm_100 GPU every compute entry
 *     point returns MAF_E_ARCH.
 *   - activations are NHWC ("channels last") fp16: pixel (n,y,x) of a `maf_tensor` starts at
 *     ptr + ((n*h + y)*w + x) * c_stride elements and holds `c` valid channels.  c_stride >= c
 *     lets a tensor be a channel slice of a wider buffer — this is how Concat / split of the
 *     reference (common.py:148-154,940,944) cost no memory traffic.  For fp16, ptr must be
 *     16-byte aligned and c_stride a multiple of 8.
 */
#ifndef MAFB200_H_
#define MAFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAFB200_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define MAFB200_API __attribute__((visibility("default")))
#else
#define MAFB200_API
#endif

/* error codes */
#define MAF_OK 0
#define MAF_E_ARG (-1)     /* bad shape / null pointer / unsupported parameter */
#define MAF_E_ALIGN (-2)   /* pointer or stride alignment violated */
#define MAF_E_ARCH (-3)    /* no sm_100 device / driver entry point missing */
#define MAF_E_CUDA (-4)    /* a CUDA runtime/driver call failed (see mafb200_last_error) */
#define MAF_E_WORKSPACE (-5) /* workspace too small */

/* dtypes */
#define MAF_F16 0
#define MAF_F32 1
#define MAF_U8 2

/* activations fused into the conv epilogues */
#define MAF_ACT_NONE 0
#define MAF_ACT_SILU 1    /* x * sigmoid(x)  — Conv, common.py:29-50 */
#define MAF_ACT_RELU 2    /* RepVGGBlock, common.py:198 */
#define MAF_ACT_SIGMOID 3 /* Head_DepthUni cls branch, common.py:1332 */

#define MAF_MAX_SRC 4 /* a MAFPN fusion stage concatenates at most 4 maps (MAF-YOLO-n.yaml:36) */

typedef struct maf_tensor {
  void* ptr;        /* device pointer to channel 0 of pixel (0,0,0) of this view */
  int32_t n, h, w;  /* batch, height, width */
  int32_t c;        /* valid channels of this view */
  int32_t c_stride; /* elements between consecutive pixels */
  int32_t dtype;    /* MAF_F16 for activations */
} maf_tensor;

MAFB200_API int32_t mafb200_version(void);
MAFB200_API const char* mafb200_last_error(void);
/* 0 if device `device` (or the current one if < 0) can run the kernels (compute capability 10.x). */
MAFB200_API int32_t mafb200_device_ok(int32_t device);

/* ---- packed-weight geometry (pure host arithmetic; usable without a GPU) -------------------
 * A conv with `cout` output channels is computed as ceil-split N tiles of `tile_n` columns
 * (tile_n % 16 == 0, tile_n <= 128).  Packed weights are fp16 [n_tiles*tile_n][k_packed],
 * row = output channel (zero rows beyond cout), columns = for each source s (1x1) or each tap
 * (ky,kx) row-major (3x3) a zero-padded block of round_up(C, 64) input channels.
 * Bias is fp32 [n_tiles*tile_n], zero padded. */
MAFB200_API int32_t mafb200_gemm_tiling(int32_t cout, int32_t* n_tiles, int32_t* tile_n);
MAFB200_API int32_t mafb200_packed_k_1x1(const int32_t* src_channels, int32_t n_src);
MAFB200_API int32_t mafb200_packed_k_3x3(int32_t cin);

/* ---- convolutions on tcgen05 tensor cores --------------------------------------------------
 * dst[m, :] = act( sum_s srcs[s][m, :] @ W_s^T + bias ),  m over all n*h*w pixels.
 * All sources and dst share n,h,w.  If dst_up2x != NULL the result is additionally written
 * nearest-upsampled x2 (nn.Upsample fused into the producer) to that [n,2h,2w,c] tensor. */
MAFB200_API int32_t mafb200_conv1x1(const maf_tensor* srcs, int32_t n_src, const void* w_packed, const float* bias,
                        int32_t act, const maf_tensor* dst, const maf_tensor* dst_up2x, void* stream);

/* 3x3, stride 2, padding 1 implicit GEMM (im2col folded into the TMA descriptor).
 * src [n,h,w,cin] -> dst [n,h/2,w/2,cout]; h and w even. */
MAFB200_API int32_t mafb200_conv3x3s2(const maf_tensor* src, const void* w_packed, const float* bias, int32_t act,
                          const maf_tensor* dst, void* stream);

/* The same conv for a NARROW input (src->c_stride == 32 fp16: one 128-byte row per pixel PAIR, c <= 32): the map is
 * read as pixel pairs, so a tile needs 6 im2col boxes instead of 9 (the TMA cost is per box row, not per byte).
 * src padding channels [c, c_stride) must hold finite values.  w_packed: fp16 [n_tiles*tile_n][6*64], block
 * o = ky*2 + po; po = 0: kx = 0 at channel offset c_stride; po = 1: kx = 1 at offset 0, kx = 2 at offset c_stride. */
MAFB200_API int32_t mafb200_conv3x3s2_pair(const maf_tensor* src, const void* w_packed, const float* bias, int32_t act,
                               const maf_tensor* dst, void* stream);

/* First layer: reads the reference's input tensor directly — NCHW, `x_dtype` in {MAF_F32, MAF_F16
 * (values in [0,1], evaler.py:161-163), MAF_U8 (raw pixels; the /255 is folded in)} — 3 input
 * channels, 3x3 stride 2 pad 1.  w: fp32 [cout][3][3][3] (co, ky, kx, ci); bias fp32 [cout]. */
MAFB200_API int32_t mafb200_stem_conv3x3s2(const void* x_nchw, int32_t x_dtype, int32_t n, int32_t h, int32_t w,
                               const float* weight, const float* bias, int32_t act, const maf_tensor* dst,
                               void* stream);

/* ---- depth-wise k x k (k in 3,5,7,9), stride 1, padding k/2 -----------------------------------
 * weight fp32 [k][k][c] (tap-major, channel fastest); bias fp32 [c]. */
MAFB200_API int32_t mafb200_dwconv(const maf_tensor* src, const float* weight, const float* bias, int32_t k, int32_t act,
                       const maf_tensor* dst, void* stream);

/* ---- depth-wise k x k fused with the 1x1 conv that consumes it (k in 3,5; cout <= 128) -------------------
 * dst = act2(W2 * act1(DW_k(src) + dw_bias) + pw_bias): DepthBottleneckUni's conv2 -> SiLU -> one_conv
 * (common.py:915-926) and Head_DepthUni's cls_conv -> cls_conv_s / reg_conv -> reg_conv_s (common.py:1328-1336);
 * the depth-wise output never goes to HBM.  dw_weight / dw_bias as for mafb200_dwconv; pw_packed / pw_bias as for
 * mafb200_conv1x1 with one source of src->c channels.  act1 in {none, silu, relu}. */
MAFB200_API int32_t mafb200_dwconv_conv1x1(const maf_tensor* src, const float* dw_weight, const float* dw_bias, int32_t k,
                               int32_t act1, const void* pw_packed, const float* pw_bias, int32_t act2,
                               const maf_tensor* dst, void* stream);

/* ---- K4: the whole DepthBottleneckUni in one kernel (k in 3,5; c_in <= 64, mid <= 192, c_out <= 64) -------------
 * dst = SiLU(W2 * SiLU(DW_k(SiLU(W1 * src + b1)) + dw_bias) + b2): conv1 (1x1, c_in -> mid) -> UniRepLKNetBlock deploy
 * form (depth-wise k x k) -> SiLU -> one_conv (1x1, mid -> c_out), yolov6/layers/common.py:898-927.  The mid-wide
 * (3 c_) intermediate never goes to HBM: a persistent warp-specialised CTA per SM computes it per 10x20-pixel tile
 * (with its k-1 halo) into shared memory from a TMA halo tile of src and feeds the taps from there.
 * Packed operands, with mid_pad = round_up(mid, 64) and tile_n = round_up(dst->c, 16):
 *   w1_packed fp16 [mid_pad][64] (row = mid channel; zero beyond mid / c_in), b1 fp32 [mid_pad];
 *   dw_weight fp32 [k*k][mid_pad] tap-major, dw_bias fp32 [mid_pad];  w2_packed fp16 [tile_n][mid_pad], b2 fp32 [tile_n].
 * mafb200_bottleneck_supported: 1 if the shape fits this kernel (pure host arithmetic), else 0. */
MAFB200_API int32_t mafb200_bottleneck_supported(int32_t c_in, int32_t mid, int32_t c_out, int32_t k);
/* debug: clock64 trace of CTA 0 of the following mafb200_bottleneck calls of this thread (device int64 [3][64][4]) */
MAFB200_API int32_t mafb200_bottleneck_trace(long long* trace_device);
MAFB200_API int32_t mafb200_bottleneck(const maf_tensor* src, int32_t mid, const void* w1_packed, const float* b1,
                           const float* dw_weight, const float* dw_bias, int32_t k, const void* w2_packed,
                           const float* b2, const maf_tensor* dst, void* stream);

/* ---- 2x2 max pool fused with the 1x1 conv that consumes it (C <= 256, cout <= 128) ------------------------
 * dst = act(W * maxpool2x2(src) + bias): the first branch of MPRep, conv1(mp(x)) (common.py:787-792, MP at
 * common.py:667-673); the pooled map never goes to HBM.  packed / bias as for mafb200_conv1x1 with one source of
 * src->c channels; even h and w; dst [n, h/2, w/2, cout], 32-B aligned with c_stride % 16 == 0. */
MAFB200_API int32_t mafb200_maxpool2x2_conv1x1(const maf_tensor* src, const void* packed, const float* bias, int32_t act,
                                   const maf_tensor* dst, void* stream);

/* ---- pooling / resampling ------------------------------------------------------------------- */
MAFB200_API int32_t mafb200_maxpool2x2(const maf_tensor* src, const maf_tensor* dst, void* stream);
/* y1 = maxpool5(x), y2 = maxpool5(y1), y3 = maxpool5(y2) (stride 1, pad 2, -inf padding). */
MAFB200_API int32_t mafb200_sppf_pool(const maf_tensor* src, const maf_tensor* y1, const maf_tensor* y2,
                          const maf_tensor* y3, void* stream);
MAFB200_API int32_t mafb200_upsample2x(const maf_tensor* src, const maf_tensor* dst, void* stream);

/* layout converters used by the block-level nn.Module drop-ins (reference tensors are NCHW fp32) */
MAFB200_API int32_t mafb200_nchw_to_nhwc_f16(const void* src, int32_t src_dtype, const maf_tensor* dst, void* stream);
MAFB200_API int32_t mafb200_nhwc_f16_to_nchw(const maf_tensor* src, void* dst, int32_t dst_dtype, void* stream);

/* ---- detect head decode ----------------------------------------------------------------------
 * For each of `n_levels` pyramid levels: cls_logits[l] [n,h_l,w_l,nc] (raw cls_pred output, the
 * sigmoid of common.py:1332 is applied here) and reg[l] [n,h_l,w_l,4*(reg_max+1)] (raw reg_pred
 * output, side-major: ch = side*(reg_max+1)+bin, yolo.py:377).  Writes pred fp32
 * [n, sum(h_l*w_l), 5+nc] = (cx, cy, w, h in input pixels, 1.0, class probabilities).
 * cls_is_prob != 0: `cls_logits` already holds sigmoid outputs (what the reference's Head_DepthUni
 * returns, common.py:1332) and is copied through — used by the per-block Detect_yaml drop-in. */
MAFB200_API int32_t mafb200_head_decode(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides,
                            int32_t n_levels, int32_t reg_max, int32_t cls_is_prob, float* pred, void* stream);

/* ---- batched NMS ------------------------------------------------------------------------------
 * Same result as yolov6/utils/nms.py:31-105 on the same `pred` ([B, A, 5+nc] fp32), without its
 * 10 s wall-clock bail-out.  `class_filter`: optional device array of nc bytes (non-zero = keep
 * class; NULL = all classes; = the `classes` argument).  Outputs: det fp32 [B, max_det, 6]
 * (x1,y1,x2,y2,score,class; rows >= count are zero) and count int32 [B].  `conf_thres` is compared
 * in fp32 and `iou_thres` in fp64 exactly as torch / torchvision do on CPU. */
MAFB200_API size_t mafb200_nms_workspace_bytes(int32_t batch, int32_t anchors, int32_t nc);
MAFB200_API int32_t mafb200_nms(const float* pred, int32_t batch, int32_t anchors, int32_t nc, double conf_thres,
                    double iou_thres, int32_t multi_label, int32_t agnostic, const uint8_t* class_filter,
                    int32_t max_det, int32_t max_nms, float* det, int32_t* count, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- serving path: decode fused with the NMS threshold / compaction pass -------------------------------
 * mafb200_head_decode_detect = mafb200_head_decode + the first half of mafb200_nms (yolov6/utils/nms.py:48-84) in
 * one kernel: boxes fp32 [n, A, 4] (cx, cy, w, h; written for candidate rows only when pred is NULL) and, in `workspace` (size / layout of mafb200_nms), the unordered
 * candidate keys + per-image counts.  pred may be NULL — the [n, A, 5+nc] tensor (91 MB at bs32) is then never
 * written nor read back.  mafb200_nms_select = the second half of mafb200_nms (sort + greedy NMS, nms.py:90-100) on
 * those candidates; box_stride = 4 for `boxes`, 5+nc when boxes points at a pred tensor.  Together they give the
 * same det / count as mafb200_head_decode followed by mafb200_nms, bit for bit. */
MAFB200_API int32_t mafb200_head_decode_detect(const maf_tensor* cls_logits, const maf_tensor* reg, const float* strides,
                                   int32_t n_levels, int32_t reg_max, int32_t cls_is_prob, float* pred, float* boxes,
                                   double conf_thres, int32_t multi_label, const uint8_t* class_filter,
                                   void* workspace, size_t workspace_bytes, void* stream);
MAFB200_API int32_t mafb200_nms_select(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors, int32_t nc,
                           double iou_thres, int32_t agnostic, int32_t max_det, int32_t max_nms, float* det,
                           int32_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7: the head's prediction convs with the detect path finished in their epilogues ---------------------------
 * mafb200_head_pred = Head_DepthUni's `cls_pred` / `reg_pred` 1x1 conv (yolov6/layers/common.py:1325-1336) of ONE
 * pyramid level with, in the epilogue of the tcgen05 GEMM and on the fp32 accumulators (the logits are never rounded
 * to fp16 nor stored):
 *   kind MAF_HEAD_CLS: the class sigmoid (common.py:1332) -> pred[b, a, 5:5+nc] (if pred) and / or the threshold pass
 *       of non_max_suppression (yolov6/utils/nms.py:48-84) -> candidate keys + counts in `workspace` (if detect_cfg);
 *   kind MAF_HEAD_REG: DFL softmax expectation + anchor decode + dist2bbox 'xywh' x stride (yolov6/models/yolo.py:
 *       377-396, yolov6/assigners/anchor_generator.py:11-25, yolov6/utils/general.py:29-40) -> pred[b, a, 0:5]
 *       = (cx, cy, w, h, 1) (if pred) and / or boxes[b, a, 0:4] (if boxes).
 * src [n, h, w, C] is the level's cls / reg tower output; anchor a = anchor_off + y * w + x of `total_anchors`;
 * pred is [n, total_anchors, 5+nc] fp32, boxes [n, total_anchors, 4] fp32 (16-B aligned).
 * Weights: kind CLS — packed as for mafb200_conv1x1 (cout = nc <= 128); kind REG — reg_max must be 16 and the 68
 * rows are permuted before packing: packed row j < 64 = channel (j / 16) * 17 + j % 16, row 64 + s = channel
 * s * 17 + 16, rows 68..79 zero (so that a box side never straddles a 32-column TMEM load).
 * detect_cfg: DEVICE copy of a maf_detect_cfg filled by mafb200_detect_cfg_fill (host); it is read when the kernel
 * runs, so one captured CUDA graph serves every threshold.  workspace: layout / size of mafb200_nms; its per-image
 * counters must be zeroed by mafb200_detect_reset (stream-ordered) before the CLS launches of a batch.
 * With the three levels' CLS + REG launches followed by mafb200_nms_select the detections equal those of the
 * pred-writing launches followed by mafb200_nms, bit for bit.
 *   kind MAF_HEAD_CLS_TRAIN / MAF_HEAD_REG_TRAIN: the TRAIN-form outputs of Detect_yaml.forward (`self.training or val_loss`,
 *       yolov6/models/yolo.py:333-354) for frozen-BN (folded) weights: `pred` is then pred_scores [n, total_anchors, nc]
 *       (sigmoid class probabilities) resp. pred_distri [n, total_anchors, 68] (raw DFL logits, side-major), fp32 from the
 *       fp32 accumulators; `boxes` and `detect_cfg` must be NULL.  Same packed weights as CLS / REG.  These are the inputs
 *       of mafb200_detect_loss (validation loss). */
#define MAF_HEAD_CLS 0
#define MAF_HEAD_REG 1
#define MAF_HEAD_CLS_TRAIN 2
#define MAF_HEAD_REG_TRAIN 3
typedef struct maf_detect_cfg {
  float conf;          /* conf_thres as fp32 (torch compares in fp32) */
  float skip_below;    /* logit bound under which sigmoid(z) cannot exceed conf */
  int32_t multi_label; /* nms.py:57: multi_label && nc > 1 */
  int32_t has_filter;  /* class_filter valid (the `classes` argument) */
  uint8_t class_filter[256];
} maf_detect_cfg;
MAFB200_API int32_t mafb200_detect_cfg_fill(maf_detect_cfg* host_cfg, double conf_thres, int32_t multi_label, int32_t nc,
                                const uint8_t* class_filter_host);
MAFB200_API int32_t mafb200_detect_reset(void* workspace, int32_t batch, void* stream);
MAFB200_API int32_t mafb200_head_pred(const maf_tensor* src, const void* w_packed, const float* bias, int32_t kind,
                          int32_t anchor_off, int32_t total_anchors, float stride, int32_t nc, float* pred,
                          float* boxes, const maf_detect_cfg* detect_cfg, void* workspace, size_t workspace_bytes,
                          void* stream);

/* mafb200_nms_select writing the packed rows of the multi-GPU detection all-gather directly: image b's [max_det, 6]
 * detections at packed + b * row_floats and its count (int32 bits) in the float behind them (row_floats >= max_det*6+1).
 * The rank's slice of the gather buffer is filled in place; one NCCL all-gather then moves detections and counts. */
MAFB200_API int32_t mafb200_nms_select_packed(const float* boxes, int32_t box_stride, int32_t batch, int32_t anchors,
                                  int32_t nc, double iou_thres, int32_t agnostic, int32_t max_det, int32_t max_nms,
                                  float* packed, int32_t row_floats, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* ---- image pre-processing (the step right before the hot path) ------------------------------------
 * letterbox (yolov6/data/data_augment.py:53-83: cv2.resize INTER_LINEAR to new_w x new_h — reproduced bit for
 * bit — then a constant border of `fill`) + HWC -> CHW and BGR -> RGB (Inferer.precess_image,
 * yolov6/core/inferer.py:168-178) of ONE uint8 image, written into the uint8 [3, dst_h, dst_w] slot the stem
 * kernel reads (mafb200_stem_conv3x3s2 with MAF_U8 folds the /255).  Geometry comes from the host exactly as
 * letterbox() computes it: resized size (new_h, new_w; equal to the source size = no resize) and the
 * top / left border.  src: device memory, HWC, src_pitch_bytes per row. */
MAFB200_API int32_t mafb200_letterbox_u8(const void* src_hwc, int32_t src_h, int32_t src_w, int32_t src_pitch_bytes,
                             void* dst_chw, int32_t dst_h, int32_t dst_w, int32_t new_h, int32_t new_w,
                             int32_t top, int32_t left, int32_t fill, int32_t swap_rb, void* stream);

/* ---- post-NMS rescaling (the step right after the hot path) --------------------------------------
 * For every valid row of det [B, max_det, 6] (row i < count[b]): (x - pad_x) / gain_x, (y - pad_y) / gain_y,
 * clamped to [0, w0] x [0, h0] — Evaler.scale_coords (yolov6/core/evaler.py:391-418) and Inferer.rescale
 * (yolov6/core/inferer.py:181-195).  params fp32 [B][6] = gain_x, gain_y, pad_x, pad_y, w0, h0 (device).
 * mode 0: out = (x1,y1,x2,y2,score,cls); mode 1: out = (x_tl,y_tl,w,h,score,cls) as box_convert + the
 * top-left shift of evaler.py:382-390,428-429.  out_cat (optional) = category_ids[int(cls)] (evaler.py:433;
 * category_ids: device int32 [nc] or NULL = the class index), -1 for padding rows; padding rows of out are 0.
 * recip_mul = 0 reproduces torch-CPU division bit for bit, 1 torch-CUDA's x * (1 / gain).  out may alias det. */
MAFB200_API int32_t mafb200_scale_detections(const float* det, const int32_t* count, int32_t batch, int32_t max_det,
                                 const float* params, const int32_t* category_ids, int32_t nc, int32_t mode,
                                 int32_t recip_mul, float* out, int32_t* out_cat, void* stream);

/* number of kernel launches issued by this library in the calling process (bench.py gpu_launches) */
/* ---- training-side kernels -------------------------------------------------------------------------
 * Task-aligned label assignment + detection loss of the reference's ComputeLoss.__call__ for epoch >= warm-up
 * (yolov6/models/loss.py:56-162; TaskAlignedAssigner(topk=13, alpha=1, beta=6), yolov6/assigners/tal_assigner.py:22-151,
 * assigner_utils.py:25-89; VarifocalLoss loss.py:181-192; BboxLoss GIoU + DFL loss.py:195-254, figure_iou.py:27-65),
 * forward value and the gradients w.r.t. the head's train-form outputs.  Replaces the numpy / python target loop of
 * loss.py:164-172 and ~40 torch ops on [B,G,8400] float64 tensors.
 *   pred_scores  [B,A,nc] fp32 class probabilities, pred_distri [B,A,68] fp32 DFL logits (Detect_yaml train branch,
 *                yolov6/models/yolo.py:333-354); A = (s/8)^2 + (s/16)^2 + (s/32)^2 anchors for image size s
 *   targets      [T,6] fp32 rows (image index, class, cx, cy, w, h normalised), as the dataloader collates them
 *   gt_cap       capacity G of boxes per image (>= the largest count; rows beyond it are counted in scalars[6])
 *   boxes_override  optional [B,A,4] fp32 xyxy in stride units used instead of decoding pred_distri (tests)
 *   scalars_out  device double[8]: loss, 2.5*loss_iou, 0.5*loss_dfl, 1.0*loss_cls (the reference's return values),
 *                target_scores_sum, number of foreground anchors, dropped target rows, 0
 *   grad_scores / grad_distri  optional fp32 [B,A,nc] / [B,A,68]: d loss / d pred_scores, d loss / d pred_distri
 *   out_gt_idx / out_fg / out_target_score  optional [B,A] int32 / uint8 / double: the assignment (target_gt_idx,
 *                fg_mask, and the one non-zero entry of each row of target_scores)
 *   assigner     MAF_ASSIGN_TAL or MAF_ASSIGN_ATSS (warm-up: target scores = IoU with the predicted box, fp32 as the reference)
 * Nothing allocates or synchronises; all kernels are enqueued on `stream`. */
#define MAF_ASSIGN_TAL 0  /* TaskAlignedAssigner(topk=13, alpha=1, beta=6): epoch_num >= warmup_epoch (loss.py:92-100) */
#define MAF_ASSIGN_ATSS 1 /* ATSSAssigner(topk=9), yolov6/assigners/atss_assigner.py:17-87: the warm-up epochs (loss.py:83-91) */
MAFB200_API size_t mafb200_loss_workspace_bytes(int32_t batch, int32_t anchors, int32_t gt_cap);
MAFB200_API int32_t mafb200_detect_loss(const float* pred_scores, const float* pred_distri, const float* targets,
                                        int32_t num_targets, int32_t batch, int32_t img_size, int32_t num_classes,
                                        int32_t gt_cap, const float* boxes_override, void* workspace,
                                        size_t workspace_bytes, double* scalars_out, float* grad_scores,
                                        float* grad_distri, int32_t* out_gt_idx, uint8_t* out_fg,
                                        double* out_target_score, int32_t assigner, void* stream);

MAFB200_API int64_t mafb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MAFB200_H_ */
