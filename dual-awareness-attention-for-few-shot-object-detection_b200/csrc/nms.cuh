// Greedy NMS and the RPN proposal layer, entirely on the device (no mask D2H, no host loop).
//
// Numerics contract (bit-exact keep sets vs lib/model/csrc/cpu/nms_cpu.cpp:17-64):
//   area  = (x2 - x1 + 1) * (y2 - y1 + 1)             fp32, every op rounded separately
//   inter = max(0, xx2 - xx1 + 1) * max(0, yy2 - yy1 + 1)
//   suppress j (lower score) by kept i when  inter / (area_i + area_j - inter) >= thresh
// nvcc would contract a*b+c into FMA, the x86 build of the reference does not, so every
// operation below is an explicit round-to-nearest intrinsic.
//
// Pipeline: boxes sorted by descending score (stable: ties -> lower input index first)
//   1. nms_mask_kernel   64x64 tiles of the upper triangle; each warp builds the 64-bit
//                        suppression word of a row with two __ballot_sync
//   2. nms_scan_kernel   one CTA per image walks the 64-box chunks in order: OR-reduce the
//                        mask column over the kept rows, resolve the diagonal word serially
//                        in registers, append kept ranks; optional early exit at max_keep
#pragma once
#include <cub/device/device_segmented_radix_sort.cuh>

#include "api_common.cuh"

namespace dana {

__device__ __forceinline__ float box_area_rn(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

__device__ __forceinline__ bool iou_suppresses(const float4 a, const float area_a, const float4 b, const float area_b,
                                               const float thresh) {
  const float xx1 = fmaxf(a.x, b.x);
  const float yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z);
  const float yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  // disjoint boxes (the common case): 0 / union is 0, -0 or NaN, none of which is >= a positive threshold -- same
  // decision as the division below, without the division
  if (inter == 0.0f && thresh > 0.0f) return false;
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr >= thresh;
}

// mask_t layout: [image][word w][row i]  (bit b of mask_t[w][i] <=> box w*64+b suppressed by box i, w*64+b > i)
// grid: (nblk, nblk, images), block 256 (8 warps x 8 rows each)
__global__ void __launch_bounds__(256) nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ counts,
                                                       int n_stride, float thresh, unsigned long long* __restrict__ mask_t,
                                                       long long mask_stride) {
  const int img = blockIdx.z;
  const int n = counts ? counts[img] : n_stride;
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;
  if (row_blk * 64 >= n || col_blk * 64 >= n) return;
  const float4* bx = boxes + static_cast<long long>(img) * n_stride;
  unsigned long long* mt = mask_t + static_cast<long long>(img) * mask_stride;

  __shared__ float4 s_row[64];
  __shared__ float s_row_area[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) {
    const int i = row_blk * 64 + tid;
    const float4 b = (i < n) ? bx[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    s_row[tid] = b;
    s_row_area[tid] = box_area_rn(b);
  }
  const int j0 = col_blk * 64 + lane, j1 = j0 + 32;
  const float4 c0 = (j0 < n) ? bx[j0] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 c1 = (j1 < n) ? bx[j1] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float a0 = box_area_rn(c0), a1 = box_area_rn(c1);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int ri = warp * 8 + r;
    const int i = row_blk * 64 + ri;
    const float4 rb = s_row[ri];
    const float ra = s_row_area[ri];
    const bool p0 = (i < n) && (j0 < n) && (j0 > i) && iou_suppresses(rb, ra, c0, a0, thresh);
    const bool p1 = (i < n) && (j1 < n) && (j1 > i) && iou_suppresses(rb, ra, c1, a1, thresh);
    const unsigned lo = __ballot_sync(0xffffffffu, p0);
    const unsigned hi = __ballot_sync(0xffffffffu, p1);
    if (lane == 0 && i < n)
      mt[static_cast<long long>(col_blk) * n_stride + i] =
          static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
  }
}

// One CTA (256 threads) per image.  kept_ranks [image][n_stride] int32 out, kept_count [image].
__global__ void __launch_bounds__(256) nms_scan_kernel(const unsigned long long* __restrict__ mask_t,
                                                       long long mask_stride, const int* __restrict__ counts,
                                                       int n_stride, int max_keep, int* __restrict__ kept_ranks,
                                                       int* __restrict__ kept_count) {
  extern __shared__ int s_kept[];  // n_stride entries
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_part[8];
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_nkept;
  const int img = blockIdx.x;
  const int n = counts ? counts[img] : n_stride;
  const unsigned long long* mt = mask_t + static_cast<long long>(img) * mask_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwords = (n + 63) / 64;
  if (tid == 0) s_nkept = 0;
  __syncthreads();
  for (int w = 0; w < nwords; ++w) {
    const int nkept = s_nkept;
    // (a) suppression state of chunk w = OR over kept rows of column word w
    unsigned long long part = 0;
    const unsigned long long* col = mt + static_cast<long long>(w) * n_stride;
    for (int k = tid; k < nkept; k += 256) part |= col[s_kept[k]];
    if (tid < 64) {
      const int i = w * 64 + tid;
      s_diag[tid] = (i < n) ? col[i] : 0ull;
    }
    unsigned plo = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(part));
    unsigned phi = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(part >> 32));
    if (lane == 0) s_part[warp] = static_cast<unsigned long long>(plo) | (static_cast<unsigned long long>(phi) << 32);
    __syncthreads();
    // (b) serial resolution of the 64 boxes of the chunk, in registers, by one thread
    if (tid == 0) {
      unsigned long long removed = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) removed |= s_part[i];
      unsigned long long d[64];
#pragma unroll
      for (int b = 0; b < 64; ++b) d[b] = s_diag[b];
      unsigned long long keep = 0;
      const int valid = min(64, n - w * 64);
#pragma unroll
      for (int b = 0; b < 64; ++b) {
        const bool alive = (b < valid) && !((removed >> b) & 1ull);
        if (alive) {
          keep |= (1ull << b);
          removed |= d[b];
        }
      }
      s_keepbits = keep;
    }
    __syncthreads();
    // (c) append kept ranks (ascending) with a popc prefix
    const unsigned long long keep = s_keepbits;
    if (tid < 64) {
      if ((keep >> tid) & 1ull) {
        const int pos = nkept + __popcll(keep & ((1ull << tid) - 1ull));
        s_kept[pos] = w * 64 + tid;
      }
    }
    __syncthreads();
    if (tid == 0) s_nkept = nkept + __popcll(keep);
    __syncthreads();
    if (max_keep > 0 && s_nkept >= max_keep) break;
  }
  const int total = s_nkept;
  const int out_n = (max_keep > 0 && total > max_keep) ? max_keep : total;
  int* out = kept_ranks + static_cast<long long>(img) * n_stride;
  for (int k = tid; k < out_n; k += 256) out[k] = s_kept[k];
  if (tid == 0) kept_count[img] = out_n;
}

// ------------------------------------------------------------------ standalone NMS glue
__global__ void iota_kernel(int* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void gather_boxes_kernel(const float4* __restrict__ boxes, const int* __restrict__ order, int n,
                                    float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = boxes[order ? order[i] : i];
}
// flags[input index] = 1 for kept ranks, then ascending compaction (single CTA, chunked scan)
__global__ void nms_flags_kernel(const int* __restrict__ kept_ranks, const int* __restrict__ kept_count,
                                 const int* __restrict__ order, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kept_count[0]) flags[order ? order[kept_ranks[i]] : kept_ranks[i]] = 1;
}
__global__ void __launch_bounds__(1024) compact_flags_kernel(const int* __restrict__ flags, int n,
                                                             long long* __restrict__ keep, int* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + tid;
    const int f = (i < n) ? flags[i] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f != 0);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int warp_off = 0;
    for (int k = 0; k < warp; ++k) warp_off += s_warp[k];
    const int base = s_base;
    if (f) keep[base + warp_off + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int k = 0; k < 32; ++k) tot += s_warp[k];
      s_base = base + tot;
    }
    __syncthreads();
  }
  if (tid == 0) count[0] = s_base;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct NmsWorkspace {
  int64_t off_keys_out, off_order, off_iota, off_sorted_boxes, off_mask, off_kept, off_count, off_flags, off_cub,
      off_segs;
  int64_t cub_bytes, total;
};

inline NmsWorkspace nms_workspace_layout(int n) {
  NmsWorkspace w;
  int64_t o = 0;
  const int64_t nw = (n + 63) / 64;
  auto take = [&](int64_t bytes) {
    const int64_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  w.off_keys_out = take(4LL * n);
  w.off_order = take(4LL * n);
  w.off_iota = take(4LL * n);
  w.off_sorted_boxes = take(16LL * n);
  w.off_mask = take(8LL * nw * n);
  w.off_kept = take(4LL * n);
  w.off_count = take(16);
  w.off_flags = take(4LL * n);
  w.off_segs = take(16);
  size_t cub_bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, cub_bytes, static_cast<const float*>(nullptr),
                                                     static_cast<float*>(nullptr), static_cast<const int*>(nullptr),
                                                     static_cast<int*>(nullptr), n, 1, static_cast<const int*>(nullptr),
                                                     static_cast<const int*>(nullptr));
  w.cub_bytes = static_cast<int64_t>(cub_bytes);
  w.off_cub = take(w.cub_bytes + 256);
  w.total = o;
  return w;
}

__global__ void set_segs_kernel(int* segs, int n) {
  segs[0] = 0;
  segs[1] = n;
}

inline int nms_run(const float* boxes, const float* scores, int n, float thresh, int64_t* keep, int32_t* count,
                   void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (n < 0 || keep == nullptr || count == nullptr) return DANA_EINVAL;
  if (n == 0) {
    DANA_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int32_t), stream));
    return DANA_OK;
  }
  if (boxes == nullptr || workspace == nullptr) return DANA_EINVAL;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) != 0) return DANA_EINVAL;
  const NmsWorkspace w = nms_workspace_layout(n);
  if (workspace_bytes < w.total) return DANA_EINVAL;
  if (4LL * n > 200 * 1024) return DANA_ENOTSUP;  // kept list lives in shared memory
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* keys_out = reinterpret_cast<float*>(ws + w.off_keys_out);
  int* order = reinterpret_cast<int*>(ws + w.off_order);
  int* iota = reinterpret_cast<int*>(ws + w.off_iota);
  float4* sorted = reinterpret_cast<float4*>(ws + w.off_sorted_boxes);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + w.off_mask);
  int* kept = reinterpret_cast<int*>(ws + w.off_kept);
  int* kcount = reinterpret_cast<int*>(ws + w.off_count);
  int* flags = reinterpret_cast<int*>(ws + w.off_flags);
  int* segs = reinterpret_cast<int*>(ws + w.off_segs);
  const int nblk = (n + 63) / 64;
  const int tb = 256, gb = (n + tb - 1) / tb;
  const int* order_used = nullptr;
  if (scores != nullptr) {
    iota_kernel<<<gb, tb, 0, stream>>>(iota, n);
    set_segs_kernel<<<1, 1, 0, stream>>>(segs, n);
    size_t cub_bytes = static_cast<size_t>(w.cub_bytes);
    DANA_CUDA_CHECK(cub::DeviceSegmentedRadixSort::SortPairsDescending(ws + w.off_cub, cub_bytes, scores, keys_out,
                                                                       iota, order, n, 1, segs, segs + 1, 0, 32,
                                                                       stream));
    order_used = order;
  }
  gather_boxes_kernel<<<gb, tb, 0, stream>>>(reinterpret_cast<const float4*>(boxes), order_used, n, sorted);
  nms_mask_kernel<<<dim3(nblk, nblk, 1), 256, 0, stream>>>(sorted, nullptr, n, thresh, mask,
                                                           static_cast<long long>(nblk) * n);
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    DANA_CUDA_CHECK(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  nms_scan_kernel<<<1, 256, 4 * n, stream>>>(mask, static_cast<long long>(nblk) * n, nullptr, n, 0, kept, kcount);
  DANA_CUDA_CHECK(cudaMemsetAsync(flags, 0, 4LL * n, stream));
  nms_flags_kernel<<<gb, tb, 0, stream>>>(kept, kcount, order_used, flags);
  compact_flags_kernel<<<1, 1024, 0, stream>>>(flags, n, reinterpret_cast<long long*>(keep), count);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // namespace dana
