// RPN proposal layer on the device, batched over images
// (lib/model/rpn/proposal_layer.py:67-190, bbox_transform.py:77-103,125-133).
//
//   decode+clip  ->  per-image stable descending sort of fg scores  ->  top pre_nms_top_n
//   ->  NMS (nms.cuh, early exit at post_nms_top_n)  ->  rows (i, x1, y1, x2, y2), zero padded
//
// The arithmetic order of the decode mirrors the reference expression by expression with explicit
// round-to-nearest intrinsics (torch CPU evaluates mul and add as separate fp32 ops, never FMA).
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "nms.cuh"

namespace dana {

// grid over B*HWA anchors.  boxes_all [B][HWA] float4, idx_all [B][HWA] int (0..HWA-1)
__global__ void proposals_decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ base_anchors,
                                        const float* __restrict__ im_info, int batch, int feat_w, int num_a,
                                        int hwa, int feat_stride, const float* __restrict__ fg_scores,
                                        float4* __restrict__ boxes_all, unsigned long long* __restrict__ keys,
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
  // One global ascending radix sort replaces B per-image sorts: key = (image << 32) | ~orderable(score).
  // orderable() is the usual monotone float->uint map, so within an image the order is descending score;
  // the sort is stable, so ties keep ascending anchor index (the documented tie rule).
  const unsigned int bits = __float_as_uint(fg_scores[gid]);
  const unsigned int ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  keys[gid] = (static_cast<unsigned long long>(b) << 32) | static_cast<unsigned long long>(~ord);
}

// Greedy NMS against the kept set, one CTA (1024 threads) per image, early exit at max_keep.
// Sorted candidates are taken 64 at a time: (1) each candidate is tested against every box kept so far
// (16 threads per candidate, kept boxes in shared memory), (2) the 64x64 IoU bits inside the chunk are built
// with warp ballots, (3) one thread resolves the chunk serially in registers.  Same fp32 arithmetic and
// ">=" rule as nms.cuh (bit-exact keep set); work is O(examined x kept) instead of O(n^2).
// The chain of chunks is serial, so its per-chunk latency is what matters: the next chunk's candidates
// (order -> box, two dependent global loads) and keys are fetched while the current chunk is resolved.
// Writes rois [B][post][5] (zero padded), optional scores / counts.
constexpr int kPropNmsThreads = 1024;
__global__ void __launch_bounds__(kPropNmsThreads)
proposals_nms_write_kernel(const float4* __restrict__ boxes_all, const int* __restrict__ order,
                           const unsigned long long* __restrict__ keys_sorted, int hwa, int n_pre, int post,
                           float thresh, float min_score, int det_format, float* __restrict__ rois,
                           float* __restrict__ roi_scores, int* __restrict__ roi_counts) {
  extern __shared__ float4 s_keep_box[];            // [post]
  float* s_keep_area = reinterpret_cast<float*>(s_keep_box + post);   // [post]
  __shared__ float4 s_cand[64];
  __shared__ float s_cand_area[64];
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned int s_dead[64];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_nkept;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long seg = static_cast<long long>(b) * hwa;
  if (tid == 0) s_nkept = 0;
  // software pipeline of the candidate fetch: box + key one chunk ahead, the `order` index two chunks ahead (the box
  // address depends on it)
  float4 next_bx = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long next_key = 0ull;
  int ord_ahead = 0;
  if (tid < 64) {
    if (tid < n_pre) {
      next_bx = boxes_all[seg + order[seg + tid]];
      next_key = keys_sorted[seg + tid];
    }
    if (64 + tid < n_pre) ord_ahead = order[seg + 64 + tid];
  }
  unsigned long long my_key = 0ull;
  __syncthreads();
  for (int base = 0; base < n_pre; base += 64) {
    const int nkept = s_nkept;
    const int valid = min(64, n_pre - base);
    if (tid < 64) {
      const float4 bx = next_bx;
      my_key = next_key;
      s_cand[tid] = bx;
      s_cand_area[tid] = box_area_rn(bx);
      // candidates at or below the score floor never enter (detections: score > thresh, inference.py:130)
      unsigned int dead0 = 0;
      if (tid < valid && min_score > -INFINITY) {
        const unsigned int ord = ~static_cast<unsigned int>(my_key & 0xFFFFFFFFull);
        const unsigned int bits = (ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord;
        dead0 = !(__uint_as_float(bits) > min_score);
      }
      s_dead[tid] = dead0;
      // prefetch the next chunk (consumed at the top of the next iteration)
      next_bx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + 64 + tid < n_pre) {
        next_bx = boxes_all[seg + ord_ahead];
        next_key = keys_sorted[seg + base + 64 + tid];
      }
      if (base + 128 + tid < n_pre) ord_ahead = order[seg + base + 128 + tid];
    }
    __syncthreads();
    // (1) against the kept set: candidate = tid / 16, sixteen threads stride the kept list
    {
      const int cj = tid >> 4, part = tid & 15;
      const float4 cb = s_cand[cj];
      const float ca = s_cand_area[cj];
      bool dead = false;
      for (int k = part; k < nkept && !dead; k += 16) dead = iou_suppresses(s_keep_box[k], s_keep_area[k], cb, ca, thresh);
      if (dead) s_dead[cj] = 1;   // benign race: all writers store 1
    }
    // (2) IoU bits inside the chunk: warp w handles rows 2w, 2w+1
    {
      const float4 c0 = s_cand[lane], c1 = s_cand[lane + 32];
      const float a0 = s_cand_area[lane], a1 = s_cand_area[lane + 32];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ri = warp * 2 + r;
        const float4 rb = s_cand[ri];
        const float ra = s_cand_area[ri];
        const bool p0 = (ri < valid) && (lane < valid) && (lane > ri) && iou_suppresses(rb, ra, c0, a0, thresh);
        const bool p1 = (ri < valid) && (lane + 32 < valid) && (lane + 32 > ri) && iou_suppresses(rb, ra, c1, a1, thresh);
        const unsigned lo = __ballot_sync(0xffffffffu, p0);
        const unsigned hi = __ballot_sync(0xffffffffu, p1);
        if (lane == 0) s_diag[ri] = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
      }
    }
    __syncthreads();
    // (3) serial resolution of the chunk
    if (warp == 0) {
      // the chunk's dead bits by ballot; the 64 diagonal words go to registers first so that the serial chain below
      // is pure ALU (it used to pay one shared-memory round trip per kept box)
      const unsigned dlo = __ballot_sync(0xffffffffu, s_dead[lane] != 0);
      const unsigned dhi = __ballot_sync(0xffffffffu, s_dead[lane + 32] != 0);
      if (lane == 0) {
        unsigned long long removed = static_cast<unsigned long long>(dlo) | (static_cast<unsigned long long>(dhi) << 32);
        unsigned long long d[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) d[j] = s_diag[j];
        unsigned long long keep = 0;
        int room = post - nkept;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const bool alive = (j < valid) && !((removed >> j) & 1ull) && room > 0;
          if (alive) {
            keep |= (1ull << j);
            removed |= d[j];
            --room;
          }
        }
        s_keepbits = keep;
      }
    }
    __syncthreads();
    const unsigned long long keep = s_keepbits;
    if (tid < 64 && ((keep >> tid) & 1ull)) {
      const int pos = nkept + __popcll(keep & ((1ull << tid) - 1ull));
      const float4 bx = s_cand[tid];
      s_keep_box[pos] = bx;
      s_keep_area[pos] = s_cand_area[tid];
      float* row = rois + (static_cast<long long>(b) * post + pos) * 5;
      // recover the score from the sorted key (low 32 bits hold ~orderable(score))
      const unsigned int ord = ~static_cast<unsigned int>(my_key & 0xFFFFFFFFull);
      const unsigned int bits = (ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord;
      if (det_format) {           // (x1, y1, x2, y2, score), utils.py:312-317
        row[0] = bx.x;
        row[1] = bx.y;
        row[2] = bx.z;
        row[3] = bx.w;
        row[4] = __uint_as_float(bits);
      } else {                    // (image, x1, y1, x2, y2), proposal_layer.py:186-188
        row[0] = static_cast<float>(b);
        row[1] = bx.x;
        row[2] = bx.y;
        row[3] = bx.z;
        row[4] = bx.w;
      }
      if (roi_scores) roi_scores[static_cast<long long>(b) * post + pos] = __uint_as_float(bits);
    }
    __syncthreads();
    if (tid == 0) s_nkept = nkept + __popcll(keep);
    __syncthreads();
    if (s_nkept >= post) break;
  }
  const int cnt = s_nkept;
  for (int k = cnt + tid; k < post; k += kPropNmsThreads) {   // zero padding (proposal_layer.py:137,188-190)
    float* row = rois + (static_cast<long long>(b) * post + k) * 5;
    row[0] = det_format ? 0.0f : static_cast<float>(b);
    row[1] = row[2] = row[3] = row[4] = 0.0f;
    if (roi_scores) roi_scores[static_cast<long long>(b) * post + k] = 0.0f;
  }
  if (tid == 0 && roi_counts) roi_counts[b] = cnt;
}

struct ProposalsWorkspace {
  int64_t off_boxes_all, off_idx_all, off_keys, off_keys_out, off_order, off_cub;
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

inline int proposals_key_bits(int batch) {
  int bits = 1;
  while ((1 << bits) < batch) ++bits;
  return 32 + bits;
}

inline ProposalsWorkspace proposals_workspace_layout(int batch, int hwa, int pre_nms_top_n) {
  ProposalsWorkspace w;
  w.n_pre = proposals_n_pre(batch, hwa, pre_nms_top_n);
  const int64_t tot = static_cast<int64_t>(batch) * hwa;
  int64_t o = 0;
  auto take = [&](int64_t bytes) {
    const int64_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  w.off_boxes_all = take(16 * tot);
  w.off_idx_all = take(4 * tot);
  w.off_keys = take(8 * tot);
  w.off_keys_out = take(8 * tot);
  w.off_order = take(4 * tot);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<const unsigned long long*>(nullptr),
                                  static_cast<unsigned long long*>(nullptr), static_cast<const int*>(nullptr),
                                  static_cast<int*>(nullptr), static_cast<int>(tot), 0, proposals_key_bits(batch));
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
  const size_t keep_smem = static_cast<size_t>(post_nms_top_n) * 20;
  if (keep_smem > 200 * 1024) return DANA_ENOTSUP;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float4* boxes_all = reinterpret_cast<float4*>(ws + w.off_boxes_all);
  int* idx_all = reinterpret_cast<int*>(ws + w.off_idx_all);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + w.off_keys);
  unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws + w.off_keys_out);
  int* order = reinterpret_cast<int*>(ws + w.off_order);
  const long long tot = static_cast<long long>(batch) * hwa;
  const int tb = 256;
  proposals_decode_kernel<<<static_cast<int>((tot + tb - 1) / tb), tb, 0, stream>>>(
      reinterpret_cast<const float4*>(deltas), reinterpret_cast<const float4*>(base_anchors), im_info, batch, feat_w,
      num_a, hwa, feat_stride, fg_scores, boxes_all, keys, idx_all);
  size_t cub_bytes = static_cast<size_t>(w.cub_bytes);
  DANA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, keys, keys_out, idx_all, order,
                                                  static_cast<int>(tot), 0, proposals_key_bits(batch), stream));
  static size_t configured_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[current_device()];
  if (keep_smem > 40 * 1024 && keep_smem > configured) {
    DANA_CUDA_CHECK(cudaFuncSetAttribute(proposals_nms_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(keep_smem)));
    configured = keep_smem;
  }
  proposals_nms_write_kernel<<<batch, kPropNmsThreads, keep_smem, stream>>>(boxes_all, order, keys_out, hwa, w.n_pre,
                                                                post_nms_top_n, nms_thresh, -INFINITY, 0, rois,
                                                                roi_scores, roi_counts);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

// ---------------------------------------------------------------------------------------------
// Detections after the forward pass (inference.py:108-142, utils.py:312-317):
// de-normalise the head's deltas (x STDS + MEANS), decode against the rois, clip to the image, divide by the
// image scale, keep fg score > score_thresh, sort by score, NMS(nms_thresh).  Same decode arithmetic, key
// sort and greedy NMS kernel as the proposal layer.  dets [B][R][5] = (x1, y1, x2, y2, score), zero padded.
// ---------------------------------------------------------------------------------------------
__global__ void detections_decode_kernel(const float* __restrict__ rois, const float* __restrict__ cls_prob,
                                         const float4* __restrict__ bbox_pred, const float* __restrict__ im_info,
                                         int batch, int r, float4 stds, float4 means, float4* __restrict__ boxes_all,
                                         unsigned long long* __restrict__ keys, int* __restrict__ idx_all) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= batch * r) return;
  const int b = gid / r;
  const float* roi = rois + static_cast<long long>(gid) * 5;
  const float4 dn = bbox_pred[gid];
  const float4 d = make_float4(__fadd_rn(__fmul_rn(dn.x, stds.x), means.x), __fadd_rn(__fmul_rn(dn.y, stds.y), means.y),
                               __fadd_rn(__fmul_rn(dn.z, stds.z), means.z), __fadd_rn(__fmul_rn(dn.w, stds.w), means.w));
  const float w = __fadd_rn(__fsub_rn(roi[3], roi[1]), 1.0f);
  const float h = __fadd_rn(__fsub_rn(roi[4], roi[2]), 1.0f);
  const float cx = __fadd_rn(roi[1], __fmul_rn(0.5f, w));
  const float cy = __fadd_rn(roi[2], __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx);
  const float pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
  const float pw = __fmul_rn(expf(d.z), w);
  const float ph = __fmul_rn(expf(d.w), h);
  const float xmax = __fsub_rn(__ldg(im_info + b * 3 + 1), 1.0f);
  const float ymax = __fsub_rn(__ldg(im_info + b * 3 + 0), 1.0f);
  const float scale = __ldg(im_info + b * 3 + 2);
  const float x1 = __fdiv_rn(fminf(fmaxf(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), 0.0f), xmax), scale);
  const float y1 = __fdiv_rn(fminf(fmaxf(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), 0.0f), ymax), scale);
  const float x2 = __fdiv_rn(fminf(fmaxf(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 0.0f), xmax), scale);
  const float y2 = __fdiv_rn(fminf(fmaxf(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 0.0f), ymax), scale);
  boxes_all[gid] = make_float4(x1, y1, x2, y2);
  idx_all[gid] = gid - b * r;
  const unsigned int bits = __float_as_uint(cls_prob[static_cast<long long>(gid) * 2 + 1]);
  const unsigned int ord = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
  keys[gid] = (static_cast<unsigned long long>(b) << 32) | static_cast<unsigned long long>(~ord);
}

inline int detections_run(const float* rois, const float* cls_prob, const float* bbox_pred, const float* im_info,
                          int batch, int r, const float* stds, const float* means, float score_thresh,
                          float nms_thresh, float* dets, int32_t* counts, void* workspace, int64_t workspace_bytes,
                          cudaStream_t stream) {
  if (!rois || !cls_prob || !bbox_pred || !im_info || !dets || !workspace || !stds || !means) return DANA_EINVAL;
  if (batch <= 0 || r <= 0) return DANA_EINVAL;
  if (reinterpret_cast<uintptr_t>(bbox_pred) & 15) return DANA_EINVAL;
  const ProposalsWorkspace w = proposals_workspace_layout(batch, r, 0);
  if (workspace_bytes < w.total) return DANA_EINVAL;
  const size_t keep_smem = static_cast<size_t>(r) * 20;
  if (keep_smem > 200 * 1024) return DANA_ENOTSUP;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float4* boxes_all = reinterpret_cast<float4*>(ws + w.off_boxes_all);
  int* idx_all = reinterpret_cast<int*>(ws + w.off_idx_all);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + w.off_keys);
  unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws + w.off_keys_out);
  int* order = reinterpret_cast<int*>(ws + w.off_order);
  const int tot = batch * r;
  detections_decode_kernel<<<(tot + 255) / 256, 256, 0, stream>>>(
      rois, cls_prob, reinterpret_cast<const float4*>(bbox_pred), im_info, batch, r,
      make_float4(stds[0], stds[1], stds[2], stds[3]), make_float4(means[0], means[1], means[2], means[3]), boxes_all,
      keys, idx_all);
  size_t cub_bytes = static_cast<size_t>(w.cub_bytes);
  DANA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, keys, keys_out, idx_all, order, tot, 0,
                                                  proposals_key_bits(batch), stream));
  static size_t configured_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[current_device()];
  if (keep_smem > 40 * 1024 && keep_smem > configured) {
    DANA_CUDA_CHECK(cudaFuncSetAttribute(proposals_nms_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(keep_smem)));
    configured = keep_smem;
  }
  proposals_nms_write_kernel<<<batch, kPropNmsThreads, keep_smem, stream>>>(boxes_all, order, keys_out, r, r, r, nms_thresh,
                                                                score_thresh, 1, dets, nullptr, counts);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // namespace dana
