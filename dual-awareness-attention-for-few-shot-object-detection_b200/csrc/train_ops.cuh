// Loss kernels of the training branch.  Small, latency-bound reductions; every sum is
// taken in a fixed order (block partials combined by one block), so the losses are bit-reproducible run to run.
//   rpn_loss_*   lib/model/rpn/rpn.py:96-116  (cross entropy over the sampled anchors, smooth-L1 sigma 3)
//   rcnn_loss    lib/model/framework/dana.py:199-215 (smooth-L1 sigma 1; 2-way cross entropy over all foreground rows,
//                the hardest 2 x fg background rows of the positive-support half and the hardest fg of the negative half)
//   smooth-L1    lib/model/utils/net_utils.py:71-85
#pragma once
#include "api_common.cuh"
#include "tc_common.cuh"

namespace dana {

__device__ __forceinline__ double block_sum_double(double v, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];   // fixed order
  __syncthreads();
  return t;   // valid in thread 0
}

__device__ __forceinline__ float smooth_l1(float d, float sigma2) {
  const float a = fabsf(d);
  return (a < 1.0f / sigma2) ? 0.5f * sigma2 * d * d : a - 0.5f / sigma2;
}

// in: RPN head output [pixels][pitch] fp32, channels [0,A) bg scores, [A,2A) fg scores, [2A,6A) deltas (a*4 + j);
// labels / targets / weights in (pixel, a) order.  partial[block] = (ce sum, ce count, weighted smooth-L1 sum).
__global__ void __launch_bounds__(256)
rpn_loss_partial_kernel(const float* __restrict__ in, long long pixels, int num_a, int pitch,
                        const signed char* __restrict__ labels, const float* __restrict__ tgt,
                        const float* __restrict__ in_w, const float* __restrict__ out_w, double* __restrict__ partial) {
  __shared__ double s_red[8];
  const long long total = pixels * num_a;
  double ce = 0.0, cnt = 0.0, box = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long px = i / num_a;
    const int a = static_cast<int>(i - px * num_a);
    const float* row = in + px * pitch;
    const int l = labels[i];
    if (l >= 0) {
      const float s0 = row[a], s1 = row[num_a + a];
      const float m = fmaxf(s0, s1);
      const float lse = m + logf(expf(s0 - m) + expf(s1 - m));
      ce += static_cast<double>(lse - (l ? s1 : s0));
      cnt += 1.0;
    }
    const float wo = out_w[i];
    if (wo != 0.0f) {
      const float wi = in_w[i];
      const float4 t = *reinterpret_cast<const float4*>(tgt + 4 * i);
      const float* d = row + 2 * num_a + a * 4;
      const float s = smooth_l1(wi * (d[0] - t.x), 9.0f) + smooth_l1(wi * (d[1] - t.y), 9.0f) +
                      smooth_l1(wi * (d[2] - t.z), 9.0f) + smooth_l1(wi * (d[3] - t.w), 9.0f);
      box += static_cast<double>(wo * s);
    }
  }
  const double a0 = block_sum_double(ce, s_red);
  const double a1 = block_sum_double(cnt, s_red);
  const double a2 = block_sum_double(box, s_red);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x + 0] = a0;
    partial[3 * blockIdx.x + 1] = a1;
    partial[3 * blockIdx.x + 2] = a2;
  }
}

__global__ void rpn_loss_final_kernel(const double* __restrict__ partial, int blocks, int batch, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ce = 0.0, cnt = 0.0, box = 0.0;
  for (int b = 0; b < blocks; ++b) {
    ce += partial[3 * b];
    cnt += partial[3 * b + 1];
    box += partial[3 * b + 2];
  }
  out[0] = static_cast<float>(cnt > 0.0 ? ce / cnt : 0.0);          // F.cross_entropy: mean over the kept anchors
  out[1] = static_cast<float>(box / static_cast<double>(batch));    // sum over (C,H,W), mean over the batch
}

// One block.  scores [2R][2] (rows [0,R) positive-support pass, [R,2R) negative-support pass), labels [R] (fp32 0/1) of
// the first half (the second half is all background), bbox_pred / targets / weights [R][4].
// out[0] = RCNN_loss_cls, out[1] = RCNN_loss_bbox.  Dynamic smem: (float key + int idx) * npow2.
__global__ void __launch_bounds__(1024)
rcnn_loss_kernel(const float* __restrict__ scores, const float* __restrict__ labels, int r, const float* __restrict__ bbox_pred,
                 const float* __restrict__ tgt, const float* __restrict__ in_w, const float* __restrict__ out_w, int npow2,
                 float* __restrict__ out) {
  extern __shared__ unsigned char s_raw[];
  float* key = reinterpret_cast<float*>(s_raw);
  int* idx = reinterpret_cast<int*>(key + npow2);
  __shared__ double s_red[32];
  __shared__ int s_nfg;
  const int n = 2 * r;
  if (threadIdx.x == 0) s_nfg = 0;
  __syncthreads();
  // keys: fg probability of the background rows (descending sort), -inf elsewhere
  int my_fg = 0;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    float k = -INFINITY;
    if (i < n) {
      const bool fg = (i < r) && (labels[i] == 1.0f);
      if (fg) {
        ++my_fg;
      } else {
        const float s0 = scores[2 * i], s1 = scores[2 * i + 1];
        const float m = fmaxf(s0, s1);
        const float e0 = expf(s0 - m), e1 = expf(s1 - m);
        k = e1 / (e0 + e1);
      }
    }
    key[i] = k;
    idx[i] = i;
  }
  if (my_fg) atomicAdd(&s_nfg, my_fg);
  __syncthreads();
  // bitonic sort, descending by key, ties by ascending row index
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < npow2 / 2; t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float ka = key[lo], kb = key[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool a_first = (ka > kb) || (ka == kb && ia < ib);     // a belongs before b in descending order
        if (a_first != desc) {
          key[lo] = kb; key[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  // smooth-L1 (sigma 1) of the box regression: sum over the 4 coordinates, mean over the R rows
  double box = 0.0;
  for (int i = threadIdx.x; i < r * 4; i += blockDim.x)
    box += static_cast<double>(out_w[i] * smooth_l1(in_w[i] * (bbox_pred[i] - tgt[i]), 1.0f));
  const double box_sum = block_sum_double(box, s_red);
  // cross entropy of every foreground row (label 1)
  double ce = 0.0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    if (labels[i] == 1.0f) {
      const float s0 = scores[2 * i], s1 = scores[2 * i + 1];
      const float m = fmaxf(s0, s1);
      ce += static_cast<double>(m + logf(expf(s0 - m) + expf(s1 - m)) - s1);
    }
  }
  const double ce_fg = block_sum_double(ce, s_red);
  if (threadIdx.x == 0) {
    const int nfg = s_nfg;
    const int half = static_cast<int>(n * 0.5);
    int q0 = min(nfg * 2, static_cast<int>(n * 0.25));
    q0 = max(1, q0);
    int q1 = max(1, min(nfg, q0));
    double ce_bg = 0.0;
    int t0 = 0, t1 = 0;
    for (int j = 0; j < npow2 && (t0 < q0 || t1 < q1); ++j) {
      if (key[j] == -INFINITY) break;                                   // past the background rows
      const int i = idx[j];
      const bool first = i < half;
      if (first ? (t0 >= q0) : (t1 >= q1)) continue;
      const float s0 = scores[2 * i], s1 = scores[2 * i + 1];
      const float m = fmaxf(s0, s1);
      ce_bg += static_cast<double>(m + logf(expf(s0 - m) + expf(s1 - m)) - s0);     // label 0
      if (first) ++t0; else ++t1;
    }
    const int cnt = nfg + t0 + t1;
    out[0] = static_cast<float>(cnt > 0 ? (ce_fg + ce_bg) / cnt : 0.0);
    out[1] = static_cast<float>(box_sum / static_cast<double>(r));
  }
}

// ---------------------------------------------------------------------------------------------
// Sibling model FSOD: the attention-RPN feature of lib/model/framework/fsod.py:96-112 --
// shot-mean of the support maps, AvgPool2d(14) to a 7x7 kernel per channel, depth-wise cross-correlation of the query
// feature with it: F.conv2d(feat [1,C,h,w], kernel [C,1,7,7], groups=C) -> [C, h-6, w-6] (no padding).
// NHWC: thread = (output pixel, 4 channels); the kh*kw taps of a pixel are 16-byte pair loads that neighbouring
// pixels share through L1, the per-image kernel (kh*kw*C fp32) is read through the read-only path.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
depthwise_xcorr_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int batch, int h, int w,
                       int c, const float* __restrict__ kern /*[B][kh*kw][C]*/, int kh, int kw, float* __restrict__ out,
                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int oh = h - kh + 1, ow = w - kw + 1, cg = c >> 2;
  const long long total = static_cast<long long>(batch) * oh * ow * cg;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(i % cg) * 4;
    long long t = i / cg;
    const int x = static_cast<int>(t % ow);
    t /= ow;
    const int y = static_cast<int>(t % oh);
    const int b = static_cast<int>(t / oh);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* kb = kern + static_cast<long long>(b) * kh * kw * c + c0;
    for (int dy = 0; dy < kh; ++dy) {
      const long long row = ((static_cast<long long>(b) * h + y + dy) * w + x) * c + c0;
      for (int dx = 0; dx < kw; ++dx) {
        const uint2 ph = __ldg(reinterpret_cast<const uint2*>(hi + row + static_cast<long long>(dx) * c));
        float f0 = __uint_as_float(ph.x << 16), f1 = __uint_as_float(ph.x & 0xFFFF0000u);
        float f2 = __uint_as_float(ph.y << 16), f3 = __uint_as_float(ph.y & 0xFFFF0000u);
        if (lo != nullptr) {
          const uint2 pl = __ldg(reinterpret_cast<const uint2*>(lo + row + static_cast<long long>(dx) * c));
          f0 += __uint_as_float(pl.x << 16); f1 += __uint_as_float(pl.x & 0xFFFF0000u);
          f2 += __uint_as_float(pl.y << 16); f3 += __uint_as_float(pl.y & 0xFFFF0000u);
        }
        const float4 k4 = __ldg(reinterpret_cast<const float4*>(kb + static_cast<long long>(dy * kw + dx) * c));
        a0 = fmaf(f0, k4.x, a0); a1 = fmaf(f1, k4.y, a1); a2 = fmaf(f2, k4.z, a2); a3 = fmaf(f3, k4.w, a3);
      }
    }
    const long long o = ((static_cast<long long>(b) * oh + y) * ow + x) * c + c0;
    if (out != nullptr) *reinterpret_cast<float4*>(out + o) = make_float4(a0, a1, a2, a3);
    if (out_hi != nullptr) {
      __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
      split_bf16(a0, h0, l0); split_bf16(a1, h1, l1); split_bf16(a2, h2, l2); split_bf16(a3, h3, l3);
      *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
      if (out_lo != nullptr) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
    }
  }
}

// out[g][i] = mean_k in[g*k + j][i]   (support_feats[:, :n_shot].mean(1), fsod.py:96,103)
__global__ void group_mean_kernel(const float* __restrict__ in, int groups, int k, long long n, float* __restrict__ out) {
  const long long total = static_cast<long long>(groups) * n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long g = i / n, e = i - g * n;
    float s = 0.0f;
    for (int j = 0; j < k; ++j) s += in[(g * k + j) * n + e];
    out[i] = s / static_cast<float>(k);
  }
}

}  // namespace dana
