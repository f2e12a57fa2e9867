/* dana_b200 -- C ABI of the Blackwell-native DAnA forward hot path.
 *
 * Every entry point takes raw device pointers, plain sizes and a CUDA stream
 * (passed as void* so the header needs no CUDA include).  Rules shared by all
 * calls: inputs are never mutated; outputs and workspaces are caller-allocated
 * (sized by the matching *_workspace_bytes query); nothing allocates, nothing
 * synchronises the stream, nothing throws.  Return value: 0 on success, a
 * negative DANA_E* code otherwise (dana_error_string() names it).
 *
 * Each declaration cites the reference interface it replaces
 * (paths relative to the reference repo root).
 */
#ifndef DANA_B200_H_
#define DANA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DANA_OK 0
#define DANA_EINVAL (-1)   /* bad argument (shape, alignment, null pointer)          */
#define DANA_ECUDA (-2)    /* CUDA runtime / driver error, see dana_last_cuda_error  */
#define DANA_EDEVICE (-3)  /* device-side protocol error recorded by a kernel        */
#define DANA_ENOTSUP (-4)  /* valid request this build does not implement           */

int dana_abi_version(void);
const char* dana_error_string(int code);
/* Last CUDA error code seen by the library on this thread (cudaError_t as int). */
int dana_last_cuda_error(void);
/* Reads (and clears) the sticky device-side error word; synchronises the device. */
int dana_device_error(void);

/* ------------------------------------------------------------------------
 * NMS -- replaces model._C.nms  (lib/model/csrc/vision.cpp:8, csrc/nms.h:10-28,
 * algorithm and comparison semantics of csrc/cpu/nms_cpu.cpp:6-75: "+1" box
 * extents, fp32, IoU >= thresh suppresses, ties broken by lower input index).
 *
 * boxes  [n,4] fp32 (x1,y1,x2,y2), scores [n] fp32.
 * keep   [n]   int64 out: kept INPUT indices in ascending order (first *count valid)
 * count  [1]   int32 out (device memory).
 * If scores == NULL the boxes are taken as already sorted by descending score.
 * ------------------------------------------------------------------------ */
int64_t dana_nms_workspace_bytes(int n);
int dana_nms(const float* boxes, const float* scores, int n, float thresh, int64_t* keep, int32_t* count,
             void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * RPN proposals -- replaces _ProposalLayer.forward
 * (lib/model/rpn/proposal_layer.py:49-190 with bbox_transform_inv / clip_boxes of
 * lib/model/rpn/bbox_transform.py:77-103,125-133 and the anchor grid of
 * lib/model/rpn/generate_anchors.py:45-105 + proposal_layer.py:79-93).
 *
 * fg_scores [B, H*W*A] fp32 in (h, w, a) order; deltas [B, H*W*A, 4] fp32, same order.
 * base_anchors [A,4] fp32; im_info [B,3] fp32 (height, width, scale).
 * rois [B, post_nms_top_n, 5] fp32 out: (batch index, x1, y1, x2, y2), zero padded.
 * roi_scores [B, post_nms_top_n] fp32 out (may be NULL); roi_counts [B] int32 out (may be NULL).
 * ------------------------------------------------------------------------ */
int64_t dana_proposals_workspace_bytes(int batch, int num_anchors_total, int pre_nms_top_n);
int dana_proposals(const float* fg_scores, const float* deltas, const float* base_anchors, const float* im_info,
                   int batch, int feat_h, int feat_w, int num_base_anchors, int feat_stride, int pre_nms_top_n,
                   int post_nms_top_n, float nms_thresh, float* rois, float* roi_scores, int32_t* roi_counts,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Detections after the forward pass -- replaces the torch ops of inference.py:108-142 and NMS() of
 * utils.py:312-317: box_deltas * STDS + MEANS (host arrays of 4 floats), bbox_transform_inv against the rois,
 * clip_boxes, / im_info scale, fg score > score_thresh, sort by score, nms(nms_thresh).
 * rois [B,R,5], cls_prob [B*R,2] (column 1 = fg), bbox_pred [B*R,4], im_info [B,3].
 * dets [B,R,5] fp32 out: (x1, y1, x2, y2, score) in kept order, zero padded; counts [B] int32 out.
 * ------------------------------------------------------------------------ */
int64_t dana_detections_workspace_bytes(int batch, int rois_per_image);
int dana_detections(const float* rois, const float* cls_prob, const float* bbox_pred, const float* im_info, int batch,
                    int rois_per_image, const float* stds, const float* means, float score_thresh, float nms_thresh,
                    float* dets, int32_t* counts, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * RoIAlign -- replaces model._C.roi_align_forward / roi_align_backward
 * (lib/model/csrc/vision.cpp:9-10, csrc/ROIAlign.h:11-45; sampling rules of
 * csrc/cpu/ROIAlign_cpu.cpp:18-219: no rounding, no half-pixel shift, adaptive
 * ceil(roi/pooled) grid when sampling_ratio <= 0).
 *
 * layout: 0 = NCHW fp32 in / [R,C,ph,pw] fp32 out (the reference's layout)
 *         1 = NHWC fp32 in / [R,ph,pw,C] out, written as fp32 (out) and/or as a
 *             bf16 hi/lo pair (out_hi/out_lo) for the tensor-core head.
 * rois [R,5] fp32 (batch index, x1, y1, x2, y2) in image pixels.
 * ------------------------------------------------------------------------ */
int64_t dana_roi_align_workspace_bytes(int batch, int channels, int height, int width, int layout);
int dana_roi_align_forward(const float* input, const float* rois, int num_rois, int batch, int channels, int height,
                           int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int layout,
                           float* out, void* out_hi, void* out_lo, void* workspace, int64_t workspace_bytes,
                           void* stream);
/* Head variant of the forward (7x7 bins): NHWC fp32 map -> [R,7,7,C] written as any of fp32 (out), bf16 pair
 * (out_hi/out_lo), bf16 pair of value + pe[bin][c] (qpe_*; the positional-encoded query of rcnn_head,
 * lib/model/framework/dana.py:259) and one fp16 plane (out_f16; the layer4 input of the mixed-precision mode).
 * pe is fp32 [49, C]. */
int dana_roi_align_head(const float* feat_nhwc, const float* rois, int num_rois, int batch, int channels, int height,
                        int width, float spatial_scale, int sampling_ratio, float* out, void* out_hi, void* out_lo,
                        const float* pe, void* qpe_hi, void* qpe_lo, void* out_f16, void* stream);
/* Backward.  layout 0: grad_out [R,C,ph,pw], grad_input [B,C,H,W] (the reference's operator, csrc/ROIAlign.h:29-45);
 * layout 1: grad_out [R,ph*pw,C], grad_input [B,H,W,C] (pipeline layout, 7x7 bins only).  grad_input is ZEROED by the
 * call and then accumulated with vector float reductions.  7x7 bins with channels % 4 == 0 run as the transpose of the
 * separable forward on NHWC (layout 0 stages through `workspace`: dana_roi_align_backward_workspace_bytes(...) bytes,
 * 256-byte aligned); other bin counts take a per-element scatter (layout 0 only, no workspace needed). */
int64_t dana_roi_align_backward_workspace_bytes(int num_rois, int batch, int channels, int height, int width,
                                                int pooled_h, int pooled_w, int layout);
int dana_roi_align_backward(const float* grad_out, const float* rois, int num_rois, int batch, int channels,
                            int height, int width, int pooled_h, int pooled_w, float spatial_scale,
                            int sampling_ratio, int layout, float* grad_input, void* workspace, int64_t workspace_bytes,
                            void* stream);

/* ------------------------------------------------------------------------
 * Episode construction -- replaces, on the device, the per-image host work of
 *   prep_im_for_blob (lib/model/utils/blob.py:35-52): astype(float32) - PIXEL_MEANS, cv2.resize(INTER_LINEAR)
 *   the support crop -> resize -> zero-pad of lib/roi_data_layer/fs_loader.py:113-138 and
 *   lib/roi_data_layer/inference_loader.py:95-109, and the HWC -> CHW permute of the loaders.
 * src: HWC image, 3 channels (BGR), u8 (src_is_f32 = 0) or f32, `src_row_pitch` elements between rows.
 * The window (crop_x, crop_y, crop_w, crop_h) is what cv2.resize would see as its source; scale_x / scale_y are
 * source pixels per destination pixel (1 / fx when the reference passes fx, crop_w / dst_w when it passes dsize).
 * mean0..2 are subtracted BEFORE interpolation (pass 0 for an already prepared image).
 * out: [3][out_h][out_w] fp32, planes of the resized dst_h x dst_w image, zero elsewhere (the padded canvas).
 * cv2's CV_32F arithmetic order is followed (row pass, then column pass; see csrc/episode.cuh). */
int dana_episode_resize(const void* src, int src_is_f32, int src_h, int src_w, int64_t src_row_pitch, int crop_x,
                        int crop_y, int crop_w, int crop_h, double scale_x, double scale_y, int dst_w, int dst_h,
                        float mean0, float mean1, float mean2, float* out, int out_h, int out_w, void* stream);

/* ------------------------------------------------------------------------
 * Tensor-core implicit-GEMM convolution / GEMM (tcgen05 + TMA).  Replaces the
 * cuDNN / cuBLAS calls behind nn.Conv2d + BatchNorm2d(eval) + ReLU + residual
 * (lib/model/framework/resnet.py:66-102, lib/model/rpn/rpn.py:28-36,63-72) and
 * nn.Linear / torch.bmm of the attention block (lib/model/framework/dana.py:124,
 * 140,142,147,266-290).
 *
 * A is an NHWC activation seen as a 4-D tensor (c, x, y, n) with arbitrary
 * x/y/n element strides (a stride-2 1x1 convolution is a strided view); a plain
 * GEMM uses x = row, y = 1.  B is [n_out][taps*c_in] (tap-major, then channel),
 * optionally one matrix per n (b_batch_stride != 0).  Operands are bf16; when the
 * *_lo planes are given each product is evaluated as hi*hi + hi*lo + lo*hi
 * (fp32-equivalent).  Output = relu?(alpha * acc * scale[co] + bias[co] + residual).
 * ------------------------------------------------------------------------ */
typedef struct dana_conv_gemm_args {
  const void* a_hi;
  const void* a_lo; /* NULL -> plain bf16 */
  int64_t a_c, a_w, a_h, a_n;       /* extents of the (strided) input view        */
  int64_t a_sx, a_sy, a_sn;         /* element strides of x, y, n (c stride is 1) */
  int32_t taps_r, taps_s;           /* filter taps (rows, cols): 1x1, 3x3, 4x1 ... */
  int32_t pad_y, pad_x;             /* tap (r,s) reads input (y + r - pad_y, x + s - pad_x) */
  const void* b_hi;
  const void* b_lo;
  int64_t b_pitch;        /* elements between consecutive output channels (>= taps*c_in) */
  int64_t b_batch_stride; /* elements between per-n matrices, 0 = shared                 */
  int32_t n_out;
  int32_t tile_w, tile_h, tile_n; /* output tile, tile_w*tile_h*tile_n == 128 (0 = choose) */
  int32_t out_w, out_h, out_n;    /* output grid                                           */
  int64_t o_sx, o_sy, o_sn;       /* output element strides (channel stride 1)             */
  void* out_hi;
  void* out_lo;
  float* out_f32;
  const float* scale; /* [n_out] or NULL */
  const float* bias;  /* [n_out] or NULL */
  int64_t bias_sn;    /* elements between per-n bias vectors, 0 = shared */
  const void* res_hi; /* residual, bf16 hi (+lo) or fp32, strides r_s* */
  const void* res_lo;
  const float* res_f32;
  int64_t r_sx, r_sy, r_sn;
  float alpha;
  int32_t relu;
  /* Optional stream-K scratch (dana_conv_gemm_workspace_bytes() bytes, zero-initialised once, private to one
   * stream).  sk_epoch must be non-zero and different for consecutive launches that share the workspace;
   * workspace == NULL or sk_epoch == 0 selects whole-tile scheduling. */
  void* workspace;
  int64_t workspace_bytes;
  int32_t sk_epoch;
  /* Fused attention softmax (dana.py:142-143): when softmax_ns > 0 the n_out columns are consecutive segments of
   * softmax_ns keys (one per shot, softmax_ns <= 256); the epilogue writes softmax_j(alpha * acc) of every segment
   * as a bf16 pair at column segment * softmax_pitch (softmax_pitch % 8 == 0, pad columns zeroed). */
  int32_t softmax_ns;
  int32_t softmax_pitch;
  /* fp16 planes (ABI 2).  ab_f16: a_hi / b_hi are IEEE fp16 planes (a_lo / b_lo must be NULL): one MMA per product,
   * 11 significant bits per operand -- the tensor-bound layers of the mixed-precision mode (layer4, RPN conv).
   * io_f16: out_hi (and res_hi, when given) are single fp16 planes, whatever the operand planes are (out_lo, res_lo,
   * out_f32, res_f32 must be NULL; n_out % 32 == 0, 16-byte aligned planes, strides % 8 == 0; saturates at 65504). */
  int32_t ab_f16;
  int32_t io_f16;
} dana_conv_gemm_args;
int64_t dana_conv_gemm_workspace_bytes(void);
int dana_conv_gemm(const dana_conv_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------
 * BA + CISA block, RPN level -- replaces the inline PyTorch of lib/model/framework/dana.py:117-151 (there is no such
 * operator in the reference; SURVEY.md section 8b introduces it behind the module boundary): positional encoding of
 * the support maps, background-attenuation gate, q / k projections with mean-centring, softmax(Q K^T / sqrt(d)) per
 * shot plus the unary term, attention-weighted values, mean over shots.
 *
 * q: the query feature, bf16 pair [batch*nq][c], row pitch q_pitch (e.g. the first channel half of the RPN input).
 * s: support maps, bf16 pair [batch*sets*shots][ns][c], image-major; set 0 of every image drives the block.
 * pe fp32 [ns][c]; wq / wk bf16 pairs [d][c] (the Linear biases cancel under the centring); un_w [c], un_b [1] the
 * unary Linear; ba_w [c], ba_b [1] the BA Linear (NULL: no BA block, semantic_enhance=False).
 * out: attended support feature [batch*nq][c] with row pitch out_pitch, as a bf16 pair (out_hi/out_lo) or, with
 * out_f16, as one fp16 plane in out_hi (out_lo NULL).  lo planes all NULL = plain-bf16 operands.
 * workspace: dana_cisa_workspace_bytes(...) bytes, 256-byte aligned, uninitialised, private to one stream.  ns <= 512. */
typedef struct dana_cisa_args {
  const void* q_hi;
  const void* q_lo;
  int64_t q_pitch;
  const void* s_hi;
  const void* s_lo;
  int32_t batch, nq, sets, shots, ns, c, d;
  const float* pe;
  const void* wq_hi;
  const void* wq_lo;
  const void* wk_hi;
  const void* wk_lo;
  const float* un_w;
  const float* un_b;
  float unary_gamma;
  const float* ba_w;
  const float* ba_b;
  float gamma;
  void* out_hi;
  void* out_lo;
  int64_t out_pitch;
  int32_t out_f16;
  void* workspace;
  int64_t workspace_bytes;
} dana_cisa_args;
int64_t dana_cisa_workspace_bytes(int batch, int nq, int sets, int shots, int ns, int c, int d);
int dana_cisa_fwd(const dana_cisa_args* args, void* stream);

/* ------------------------------------------------------------------------
 * CUDA-core stages of the path (each replaces the torch ops named).
 * bf16 "pairs" are (hi, lo) planes with x ~= hi + lo; lo may be NULL.
 * ------------------------------------------------------------------------ */
/* Stem, lib/model/framework/resnet.py:109-113.  conv1 7x7/2 pad 3 is run by dana_conv_gemm as a 4-tap
 * K=256 GEMM over the 2x2 space-to-depth image written by dana_stem_s2d:
 *   in [B,3,H,W] fp32 NCHW -> pair [B, ceil(H/2), ceil(W/2)+4, 16] (2 zero pixels left, >=2 right).
 * dana_maxpool3x3s2: MaxPool2d(3, 2, 0, ceil_mode=True) on an NHWC pair. */
int dana_stem_s2d(const float* in_nchw, int batch, int height, int width, void* out_hi, void* out_lo, void* stream);
int dana_maxpool3x3s2(const void* in_hi, const void* in_lo, int batch, int h, int w, int c, void* out_hi, void* out_lo,
                      void* stream);
/* nn.AvgPool2d(k, stride=1) on NHWC pairs -> fp32 NHWC (lib/model/framework/dana.py:42,114). */
int dana_avgpool(const void* in_hi, const void* in_lo, int maps, int h, int w, int c, int k, float* out, void* stream);
/* Support side of BA + CISA (dana.py:126-147; rcnn_head :255-276 with ba_w == NULL): positional
 * encoding, background-attenuation gate, unary term r, mean-centred k-projection input (vc) and the
 * transposed values vt[set][c][shot*seg_pitch + n] (row pitch vt_pitch, seg_pitch >= ns; pad columns zeroed by the
 * call) for the P.V contraction.  cbar_hi/lo (optional, [sets][c] pair): rbar + mean over shots of the column means -- the row-constant
 * part of the attended value when the head contracts P with the centred values. */
int dana_support_prepare(const void* in_hi, const void* in_lo, const float* in_f32, const float* pe, int maps,
                         int shots, int ns, int c, const float* ba_w, const float* ba_b, float gamma,
                         const float* un_w, const float* un_b, float unary_gamma, float* v, float* logit, float* g,
                         float* r, float* colmean, void* vc_hi, void* vc_lo, void* vt_hi, void* vt_lo,
                         int64_t vt_pitch, int seg_pitch, float* rbar, void* cbar_hi, void* cbar_lo, void* stream);
/* x - x.mean(1, keepdim=True) over groups of rows (dana.py:125,141,267,272) -> pair.
 * sums: fp32 scratch of groups*c*(1 + ceil(group_rows/64)) elements, needed when group_rows > 256 (column sums are
 * reduced in a fixed order, no float atomics: results are bit-reproducible); may be NULL for small groups. */
int dana_center_rows(const float* in, int groups, int group_rows, int c, void* out_hi, void* out_lo, float* sums,
                     void* stream);
/* Same for an input whose rows are in_pitch elements apart (a column slice of a wider matrix); group_rows <= 256. */
int dana_center_rows_pitched(const float* in, int64_t in_pitch, int groups, int group_rows, int c, void* out_hi,
                             void* out_lo, void* stream);
/* F.softmax(logits, dim=2) per shot segment (dana.py:143,274) -> pair, pad columns zeroed. */
int dana_attn_softmax(const float* logits, int64_t rows, int segs, int ns, int pitch, void* p_hi, void* p_lo,
                      void* stream);
/* RPN (bg, fg) pair softmax + delta repack (lib/model/rpn/rpn.py:47-72, proposal_layer.py:67,97-103). */
int dana_rpn_fg_prob(const float* in, int64_t pixels, int num_a, int pitch, float* fg, float* deltas, void* stream);
int dana_add_pe_split(const float* in, const float* pe, int64_t rows, int c, int period, int64_t out_pitch,
                      void* out_hi, void* out_lo, void* stream);
int dana_split_f32(const float* in, int64_t n, void* out_hi, void* out_lo, void* stream);
/* pair [rows][c] (row pitch in_pitch elements) -> contiguous fp32 [rows][c] (out, may be NULL) and / or one fp16
 * plane with row pitch f16_pitch (out_f16, may be NULL; saturating) */
int dana_merge_pair(const void* in_hi, const void* in_lo, int64_t rows, int c, int64_t in_pitch, float* out,
                    void* out_f16, int64_t f16_pitch, void* stream);
/* .mean(3).mean(2) of the layer4 output (dana.py:387-389): [items][sp][c] -> [items][c] fp32 and / or pair.
 * Input: bf16 pair, or (in_is_f16) one fp16 plane in in_hi.  c % 8 == 0, 16-byte aligned planes. */
int dana_spatial_mean(const void* in_hi, const void* in_lo, int in_is_f16, int64_t items, int sp, int c, float* out,
                      void* out_hi, void* out_lo, void* stream);
int dana_softmax2(const float* in, int64_t rows, float* out, void* stream);
int dana_nhwc_pair_to_nchw(const void* in_hi, const void* in_lo, int batch, int c, int hw, float* out, void* stream);
/* Key-major relayout: in fp32 [maps][ns][c] -> pair out[maps/shots][c][vt_pitch], element (set, ch, slot*seg_pitch + n),
 * pad columns zeroed.  Builds the B operand (V W^T)^T of the head, where the Linear(2048->64) of dana.py:288 is applied
 * to the support values before the attention-weighted sum of :281 ((P V) W^T = P (V W^T)). */
int dana_transpose_segments(const float* in, int maps, int shots, int ns, int c, int seg_pitch, int64_t vt_pitch,
                            void* out_hi, void* out_lo, void* stream);

/* ------------------------------------------------------------------------
 * Training branch: the loss layers.  The target layers that feed them are host code
 * (dana_b200/targets.py: numpy RNG sampling like the reference's); all tensors below are device memory.
 *
 * dana_rpn_losses -- lib/model/rpn/rpn.py:96-116 with _smooth_l1_loss of lib/model/utils/net_utils.py:71-85 (sigma 3).
 *   rpn_out [pixels][pitch] fp32 NHWC RPN head output: channels [0,A) bg scores, [A,2A) fg scores, [2A,6A) deltas;
 *   labels int8 [pixels*A] (1 fg, 0 bg, -1 ignored), bbox_targets [pixels*A][4], inside_w / outside_w [pixels*A], all
 *   in (pixel, anchor) order.  losses[0] = rpn_loss_cls (mean cross entropy over labels >= 0), losses[1] = rpn_loss_box
 *   (sum over anchors and coordinates, mean over the batch).
 * dana_rcnn_losses -- lib/model/framework/dana.py:199-215.  cls_scores [2*rois][2]: rows [0,rois) from the positive
 *   support set, [rois,2*rois) from the negative one; labels fp32 [rois] (0 / 1) of the first half (the second half is
 *   background); bbox_pred / bbox_targets / inside_w / outside_w [rois][4].  losses[0] = RCNN_loss_cls: cross entropy
 *   over every foreground row, the max(1, min(2 fg, rows/4)) hardest background rows of the first half and the
 *   max(1, min(fg, that)) hardest of the second; losses[1] = RCNN_loss_bbox (smooth-L1 sigma 1, mean over rois). */
int64_t dana_rpn_losses_workspace_bytes(void);
int dana_rpn_losses(const float* rpn_out, int batch, int64_t pixels, int num_a, int pitch, const int8_t* labels,
                    const float* bbox_targets, const float* inside_w, const float* outside_w, float* losses,
                    void* workspace, int64_t workspace_bytes, void* stream);
int dana_rcnn_losses(const float* cls_scores, const float* labels, int rois, const float* bbox_pred,
                     const float* bbox_targets, const float* inside_w, const float* outside_w, float* losses,
                     void* stream);

/* ------------------------------------------------------------------------
 * Sibling model FSOD, the attention-RPN feature of lib/model/framework/fsod.py:96-112:
 * dana_group_mean     support_feats[:, :n_shot].mean(1) on fp32 [groups*k][n] -> [groups][n]
 * dana_depthwise_xcorr  F.conv2d(feat, kernel.view(C,1,kh,kw), groups=C), no padding: NHWC pair [B,h,w,C] and a
 *                     per-image kernel fp32 [B][kh*kw][C] -> [B,h-kh+1,w-kw+1,C] as fp32 (out) and / or bf16 pair. */
int dana_group_mean(const float* in, int groups, int k, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward pass (training step, BASELINE configs[3]).  The gradient GEMMs run on dana_conv_gemm: the data-gradient of
 * a convolution is the same implicit GEMM with transposed / rotated weights; the weight-gradient
 *     dW[co][tap][ci] = sum_p g[p][co] * x[p + tap][ci]
 * is a K-major GEMM over the pixel index on channel-major operands, which the two layout kernels below write.
 * Together they replace the cuDNN backward-data / backward-filter calls that torch.autograd issues for
 * lib/model/framework/resnet.py:66-102 and dana.py:120-151,244-290 under train.py:138 (loss.backward()).
 *
 * dana_grad_prepare   g' = grad * (relu_out > 0) (relu_out NULL: no mask) on fp32 [pixels][channels].
 *                     Any of: out_f32 (masked gradient, the residual branch's share), out_hi/out_lo (NHWC bf16 pair,
 *                     operand of the data-gradient), t_hi/t_lo (channel-major pair [channels][t_pitch], t_pitch % 8 == 0,
 *                     t_pitch >= pixels; operand of the weight-gradient).
 * dana_im2col_t       channel-major im2col of an NHWC bf16 pair x[batch][height][width][channels] (element strides,
 *                     channel stride 1, channels % 8 == 0): t[(tap*channels + c)][(n*OH + oy)*OW + ox] =
 *                     x[n][oy*s + r - pad][ox*s + t - pad][c], zero outside.  ksize 1 (any stride s) or 3 (s = 1, pad 1).
 * dana_sgd_momentum   torch.optim.SGD.step with momentum (train.py:89,139) on a flat fp32 buffer:
 *                     d = grad*grad_scale + weight_decay*p;  m = momentum*m + d;  p -= lr*m.
 */
int dana_grad_prepare(const float* grad, const float* relu_out, int64_t pixels, int channels, float* out_f32,
                      void* out_hi, void* out_lo, void* t_hi, void* t_lo, int64_t t_pitch, void* stream);
int dana_im2col_t(const void* x_hi, const void* x_lo, int batch, int height, int width, int channels, int64_t stride_n,
                  int64_t stride_y, int64_t stride_x, int ksize, int conv_stride, void* t_hi, void* t_lo, int64_t t_pitch,
                  void* stream);
/* dana_pack_conv_weight    operand planes of an fp32 weight w[co][ci][kh][kw] (x scale[co], the frozen-BN fold; NULL: none)
 *                          for the forward / weight-gradient GEMM  f[co][(r*kw+s)*ci_n + ci]  and (dgrad_* non-NULL) for the
 *                          data-gradient GEMM  d[ci][(r*kw+s)*co_n + co] = w[co][ci][kh-1-r][kw-1-s] * scale[co].
 * dana_unpack_conv_wgrad   dW[co][ci][tap] = wgrad[co][tap*ci_n + ci] * scale[co]: the weight-gradient GEMM's output in the
 *                          parameter's own layout (what torch.autograd accumulates into .grad). */
int dana_pack_conv_weight(const float* weight, const float* scale, int out_channels, int in_channels, int kh, int kw,
                          void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* dgrad_lo, void* stream);
int dana_unpack_conv_wgrad(const float* wgrad, const float* scale, int out_channels, int in_channels, int taps,
                           float* weight_grad, int accumulate, void* stream);

/* dana_conv_backward: the whole backward of y = relu?(conv(x, W*s) + t (+ res)) in one call (what torch.autograd does with
 * cudnn_convolution_backward + threshold_backward for resnet.py:66-102 under train.py:138): dana_grad_prepare on grad_out
 * (mask by relu_out when given), the data-gradient GEMM into dx (NULL: skipped; zero-filled first when stride > 1), the
 * transposing im2col of the saved input pair + the weight-gradient GEMM + unpack into dw (NULL: skipped; dw_accumulate:
 * dw += instead of dw =, so that a parameter used twice -- query and support trunk -- or a gradient arena is written
 * in place), dres = the masked fp32 gradient (the residual branch's share / the bias gradient's summand; NULL: skipped).
 * x is the saved NHWC input pair [batch][height][width][in_channels] (element strides x_s*); wd_* the data-gradient weight
 * planes of dana_pack_conv_weight; workspace >= dana_conv_backward_workspace_bytes(), 256-byte aligned, private to the stream;
 * gemm_workspace / sk_epoch as in dana_conv_gemm_args (two consecutive epochs are used). */
typedef struct dana_conv_bwd_args {
  int32_t batch, height, width, in_channels, out_channels, ksize, stride;
  const float* grad_out;
  const float* relu_out;
  const void* x_hi;
  const void* x_lo;
  int64_t x_sn, x_sy, x_sx;
  const void* wd_hi;
  const void* wd_lo;
  const float* scale;
  float* dx;
  float* dw;
  float* dres;
  int32_t dw_accumulate;
  void* workspace;
  int64_t workspace_bytes;
  void* gemm_workspace;
  int64_t gemm_workspace_bytes;
  int32_t sk_epoch;
} dana_conv_bwd_args;
int64_t dana_conv_backward_workspace_bytes(int batch, int height, int width, int in_channels, int out_channels, int ksize,
                                           int stride);
int dana_conv_backward(const dana_conv_bwd_args* args, void* stream);
int dana_sgd_momentum(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                      float weight_decay, float grad_scale, void* stream);
int dana_depthwise_xcorr(const void* in_hi, const void* in_lo, int batch, int h, int w, int c, const float* kernel,
                         int kh, int kw, float* out, void* out_hi, void* out_lo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DANA_B200_H_ */
