// RPN proposal layer on the device, batched over images
// (lib/model/rpn/proposal_layer.py:67-190, bbox_transform.py:77-103,125-133).
//
//   decode+clip  ->  per-image stable descending sort of fg scores  ->  top pre_nms_top_n
//   ->  NMS (nms.cuh, early exit at post_nms_top_n)  ->  rows (i, x1, y1, x2, y2), zero padded
//
// The arithmetic order of the decode mirrors the reference expression by expression with explicit
// round-to-nearest intrinsics (torch CPU evaluates mul and add as separate fp32 ops, never FMA).
#pragma once
#include "nms.cuh"

namespace dana {

// grid over B*HWA anchors.  boxes_all [B][HWA] float4, idx_all [B][HWA] int (0..HWA-1)
__global__ void proposals_decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ base_anchors,
                                        const float* __restrict__ im_info, int batch, int feat_w, int num_a,
                                        int hwa, int feat_stride, float4* __restrict__ boxes_all,
                                        int* __restrict__ idx_all) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(batch) * hwa) return;
  const int b = static_cast<int>(gid / hwa);
  const int i = static_cast<int>(gid - static_cast<long long>(b) * hwa);
  const int a = i % num_a;
  const int cell = i / num_a;
  const int gx = cell % feat_w, gy = cell / feat_w;
  const float4 ba = __ldg(base_anchors + a);
  const float sx = static_cast<float>(gx * feat_stride), sy = static_cast<float>(gy * feat_stride);
  const float ax1 = __fadd_rn(ba.x, sx), ay1 = __fadd_rn(ba.y, sy);
  const float ax2 = __fadd_rn(ba.z, sx), ay2 = __fadd_rn(ba.w, sy);
  const float4 d = deltas[gid];
  // bbox_transform_inv
  const float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f);
  const float h = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
  const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w));
  const float cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx);
  const float pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
  const float pw = __fmul_rn(expf(d.z), w);
  const float ph = __fmul_rn(expf(d.w), h);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  float y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  float y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  // clip_boxes: x in [0, W-1], y in [0, H-1]
  const float xmax = __fsub_rn(__ldg(im_info + b * 3 + 1), 1.0f);
  const float ymax = __fsub_rn(__ldg(im_info + b * 3 + 0), 1.0f);
  x1 = fminf(fmaxf(x1, 0.0f), xmax);
  y1 = fminf(fmaxf(y1, 0.0f), ymax);
  x2 = fminf(fmaxf(x2, 0.0f), xmax);
  y2 = fminf(fmaxf(y2, 0.0f), ymax);
  boxes_all[gid] = make_float4(x1, y1, x2, y2);
  idx_all[gid] = i;
}

__global__ void proposals_segs_kernel(int* segs, int batch, int hwa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= batch) segs[i] = i * hwa;
}

// sorted_boxes [B][n_pre] = boxes_all[b][order[b][r]]
__global__ void proposals_gather_kernel(const float4* __restrict__ boxes_all, const int* __restrict__ order, int batch,
                                        int hwa, int n_pre, float4* __restrict__ sorted_boxes) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(batch) * n_pre) return;
  const int b = static_cast<int>(gid / n_pre);
  const int r = static_cast<int>(gid - static_cast<long long>(b) * n_pre);
  sorted_boxes[gid] = boxes_all[static_cast<long long>(b) * hwa + order[static_cast<long long>(b) * hwa + r]];
}

// rois [B][post][5]; one CTA per image
__global__ void proposals_write_kernel(const float4* __restrict__ sorted_boxes, const float* __restrict__ sorted_scores,
                                       const int* __restrict__ kept_ranks, const int* __restrict__ kept_count,
                                       int n_pre, int hwa, int post, float* __restrict__ rois,
                                       float* __restrict__ roi_scores, int* __restrict__ roi_counts) {
  const int b = blockIdx.x;
  const int cnt = min(kept_count[b], post);
  for (int k = threadIdx.x; k < post; k += blockDim.x) {
    float* row = rois + (static_cast<long long>(b) * post + k) * 5;
    row[0] = static_cast<float>(b);
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    if (k < cnt) {
      const int r = kept_ranks[static_cast<long long>(b) * n_pre + k];
      bx = sorted_boxes[static_cast<long long>(b) * n_pre + r];
      sc = sorted_scores[static_cast<long long>(b) * hwa + r];
    }
    row[1] = bx.x;
    row[2] = bx.y;
    row[3] = bx.z;
    row[4] = bx.w;
    if (roi_scores) roi_scores[static_cast<long long>(b) * post + k] = sc;
  }
  if (threadIdx.x == 0 && roi_counts) roi_counts[b] = cnt;
}

struct ProposalsWorkspace {
  int64_t off_boxes_all, off_idx_all, off_keys_out, off_order, off_segs, off_sorted, off_mask, off_kept, off_count,
      off_cub;
  int64_t cub_bytes, total;
  int n_pre;
};

// Reference quirk kept on purpose (proposal_layer.py:148): the truncation to pre_nms_top_n is only
// applied when pre_nms_top_n < numel of the WHOLE batch score tensor.
inline int proposals_n_pre(int batch, int hwa, int pre_nms_top_n) {
  if (pre_nms_top_n > 0 && static_cast<long long>(pre_nms_top_n) < static_cast<long long>(batch) * hwa)
    return hwa < pre_nms_top_n ? hwa : pre_nms_top_n;
  return hwa;
}

inline ProposalsWorkspace proposals_workspace_layout(int batch, int hwa, int pre_nms_top_n) {
  ProposalsWorkspace w;
  w.n_pre = proposals_n_pre(batch, hwa, pre_nms_top_n);
  const int64_t tot = static_cast<int64_t>(batch) * hwa;
  const int64_t nw = (w.n_pre + 63) / 64;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  w.off_boxes_all = take(16 * tot);
  w.off_idx_all = take(4 * tot);
  w.off_keys_out = take(4 * tot);
  w.off_order = take(4 * tot);
  w.off_segs = take(4LL * (batch + 1));
  w.off_sorted = take(16LL * batch * w.n_pre);
  w.off_mask = take(8LL * batch * nw * w.n_pre);
  w.off_kept = take(4LL * batch * w.n_pre);
  w.off_count = take(4LL * batch);
  size_t cub_bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, cub_bytes, static_cast<const float*>(nullptr),
                                                     static_cast<float*>(nullptr), static_cast<const int*>(nullptr),
                                                     static_cast<int*>(nullptr), static_cast<int>(tot), batch,
                                                     static_cast<const int*>(nullptr), static_cast<const int*>(nullptr));
  w.cub_bytes = static_cast<int64_t>(cub_bytes);
  w.off_cub = take(w.cub_bytes + 256);
  w.total = o;
  return w;
}

inline int proposals_run(const float* fg_scores, const float* deltas, const float* base_anchors, const float* im_info,
                         int batch, int feat_h, int feat_w, int num_a, int feat_stride, int pre_nms_top_n,
                         int post_nms_top_n, float nms_thresh, float* rois, float* roi_scores, int32_t* roi_counts,
                         void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (!fg_scores || !deltas || !base_anchors || !im_info || !rois || !workspace) return DANA_EINVAL;
  if (batch <= 0 || feat_h <= 0 || feat_w <= 0 || num_a <= 0 || post_nms_top_n <= 0) return DANA_EINVAL;
  if ((reinterpret_cast<uintptr_t>(deltas) & 15) || (reinterpret_cast<uintptr_t>(base_anchors) & 15))
    return DANA_EINVAL;
  const int hwa = feat_h * feat_w * num_a;
  const ProposalsWorkspace w = proposals_workspace_layout(batch, hwa, pre_nms_top_n);
  if (workspace_bytes < w.total) return DANA_EINVAL;
  if (4LL * w.n_pre > 200 * 1024) return DANA_ENOTSUP;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float4* boxes_all = reinterpret_cast<float4*>(ws + w.off_boxes_all);
  int* idx_all = reinterpret_cast<int*>(ws + w.off_idx_all);
  float* keys_out = reinterpret_cast<float*>(ws + w.off_keys_out);
  int* order = reinterpret_cast<int*>(ws + w.off_order);
  int* segs = reinterpret_cast<int*>(ws + w.off_segs);
  float4* sorted = reinterpret_cast<float4*>(ws + w.off_sorted);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + w.off_mask);
  int* kept = reinterpret_cast<int*>(ws + w.off_kept);
  int* kcount = reinterpret_cast<int*>(ws + w.off_count);
  const long long tot = static_cast<long long>(batch) * hwa;
  const int tb = 256;
  proposals_decode_kernel<<<static_cast<int>((tot + tb - 1) / tb), tb, 0, stream>>>(
      reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(base_anchors), im_info, batch, feat_w,
      num_a, hwa, feat_stride, boxes_all, idx_all);
  proposals_segs_kernel<<<1, 256, 0, stream>>>(segs, batch, hwa);
  size_t cub_bytes = static_cast<size_t>(w.cub_bytes);
  DANA_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortPairsDescending(ws + w.off_cub, cub_bytes, fg_scores, keys_out,
                                                                     idx_all, order, static_cast<int>(tot), batch, segs,
                                                                     segs + 1, 0, 32, stream));
  const long long tg = static_cast<long long>(batch) * w.n_pre;
  proposals_gather_kernel<<<static_cast<int>((tg + tb - 1) / tb), tb, 0, stream>>>(boxes_all, order, batch, hwa,
                                                                                     w.n_pre, sorted);
  const int nblk = (w.n_pre + 63) / 64;
  const long long mstride = static_cast<long long>(nblk) * w.n_pre;
  nms_mask_kernel<<<dim3(nblk, nblk, batch), 256, 0, stream>>>(sorted, nullptr, w.n_pre, nms_thresh, mask, mstride);
  static bool configured = false;
  if (!configured) {
    DANA_CUDA_CHECK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  nms_scan_kernel<<<batch, 256, 4 * w.n_pre, stream>>>(mask, mstride, nullptr, w.n_pre, post_nms_top_n, kept, kcount);
  proposals_write_kernel<<<batch, 256, 0, stream>>>(sorted, keys_out, kept, kcount, w.n_pre, hwa, post_nms_top_n, rois,
                                                    roi_scores, roi_counts);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // namespace dana
