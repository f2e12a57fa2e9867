// libdana_b200.so -- single translation unit: kernels + extern "C" entry points (include/dana_b200.h).
#include <math.h>
#include <string.h>

#include "api_common.cuh"
#include "tc_common.cuh"
#include "conv_gemm_api.cuh"
#include "nms.cuh"
#include "proposals.cuh"
#include "roi_align.cuh"
#include "dana_ops.cuh"
#include "episode.cuh"
#include "train_ops.cuh"

using namespace dana;

extern "C" {

int dana_abi_version(void) { return 2; }

const char* dana_error_string(int code) {
  switch (code) {
    case DANA_OK: return "ok";
    case DANA_EINVAL: return "invalid argument";
    case DANA_ECUDA: return "CUDA error";
    case DANA_EDEVICE: return "device-side protocol error";
    case DANA_ENOTSUP: return "not supported";
    default: return "unknown error";
  }
}

int dana_last_cuda_error(void) { return t_last_cuda_error; }

int dana_device_error(void) {
  int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_device_error, sizeof(int)) != cudaSuccess) return -1;
  if (v != 0) {
    const int zero = 0;
    cudaMemcpyToSymbol(g_device_error, &zero, sizeof(int));
  }
  return v;
}

int64_t dana_nms_workspace_bytes(int n) { return n <= 0 ? 256 : nms_workspace_layout(n).total; }

int dana_nms(const float* boxes, const float* scores, int n, float thresh, int64_t* keep, int32_t* count,
             void* workspace, int64_t workspace_bytes, void* stream) {
  return nms_run(boxes, scores, n, thresh, keep, count, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int64_t dana_proposals_workspace_bytes(int batch, int num_anchors_total, int pre_nms_top_n) {
  if (batch <= 0 || num_anchors_total <= 0) return 256;
  return proposals_workspace_layout(batch, num_anchors_total, pre_nms_top_n).total;
}

int dana_proposals(const float* fg_scores, const float* deltas, const float* base_anchors, const float* im_info,
                   int batch, int feat_h, int feat_w, int num_base_anchors, int feat_stride, int pre_nms_top_n,
                   int post_nms_top_n, float nms_thresh, float* rois, float* roi_scores, int32_t* roi_counts,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  return proposals_run(fg_scores, deltas, base_anchors, im_info, batch, feat_h, feat_w, num_base_anchors, feat_stride,
                       pre_nms_top_n, post_nms_top_n, nms_thresh, rois, roi_scores, roi_counts, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream));
}

int64_t dana_detections_workspace_bytes(int batch, int rois_per_image) {
  if (batch <= 0 || rois_per_image <= 0) return 256;
  return proposals_workspace_layout(batch, rois_per_image, 0).total;
}

int dana_detections(const float* rois, const float* cls_prob, const float* bbox_pred, const float* im_info, int batch,
                    int rois_per_image, const float* stds, const float* means, float score_thresh, float nms_thresh,
                    float* dets, int32_t* counts, void* workspace, int64_t workspace_bytes, void* stream) {
  return detections_run(rois, cls_prob, bbox_pred, im_info, batch, rois_per_image, stds, means, score_thresh,
                        nms_thresh, dets, counts, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int64_t dana_roi_align_workspace_bytes(int batch, int channels, int height, int width, int layout) {
  if (layout != 0) return 256;
  return 4LL * batch * channels * height * width + 256;
}

int dana_roi_align_forward(const float* input, const float* rois, int num_rois, int batch, int channels, int height,
                           int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int layout,
                           float* out, void* out_hi, void* out_lo, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  return roi_align_forward_run(input, rois, num_rois, batch, channels, height, width, pooled_h, pooled_w, spatial_scale,
                               sampling_ratio, layout, out, out_hi, out_lo, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

int dana_roi_align_head(const float* feat_nhwc, const float* rois, int num_rois, int batch, int channels, int height,
                        int width, float spatial_scale, int sampling_ratio, float* out, void* out_hi, void* out_lo,
                        const float* pe, void* qpe_hi, void* qpe_lo, void* out_f16, void* stream) {
  return roi_align_head_run(feat_nhwc, rois, num_rois, batch, channels, height, width, spatial_scale, sampling_ratio,
                            out, out_hi, out_lo, pe, qpe_hi, qpe_lo, out_f16, static_cast<cudaStream_t>(stream));
}

int64_t dana_roi_align_backward_workspace_bytes(int num_rois, int batch, int channels, int height, int width,
                                                int pooled_h, int pooled_w, int layout) {
  return roi_align_backward_workspace(num_rois, batch, channels, height, width, pooled_h, pooled_w, layout);
}

int dana_roi_align_backward(const float* grad_out, const float* rois, int num_rois, int batch, int channels,
                            int height, int width, int pooled_h, int pooled_w, float spatial_scale,
                            int sampling_ratio, int layout, float* grad_input, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  return roi_align_backward_run(grad_out, rois, num_rois, batch, channels, height, width, pooled_h, pooled_w,
                                spatial_scale, sampling_ratio, layout, grad_input, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

int64_t dana_conv_gemm_workspace_bytes(void) { return 4096 + static_cast<int64_t>(sm_count()) * 128 * 256 * 4; }

int dana_conv_gemm(const dana_conv_gemm_args* args, void* stream) {
  return conv_gemm_dispatch(args, static_cast<cudaStream_t>(stream));
}

int dana_episode_resize(const void* src, int src_is_f32, int src_h, int src_w, int64_t src_row_pitch, int crop_x,
                        int crop_y, int crop_w, int crop_h, double scale_x, double scale_y, int dst_w, int dst_h,
                        float mean0, float mean1, float mean2, float* out, int out_h, int out_w, void* stream) {
  return episode_resize_run(src, src_is_f32, src_h, src_w, src_row_pitch, crop_x, crop_y, crop_w, crop_h, scale_x,
                            scale_y, dst_w, dst_h, mean0, mean1, mean2, out, out_h, out_w,
                            static_cast<cudaStream_t>(stream));
}

#include "dana_ops_api.inc"
#include "cisa_api.inc"
#include "train_api.inc"

}  // extern "C"
