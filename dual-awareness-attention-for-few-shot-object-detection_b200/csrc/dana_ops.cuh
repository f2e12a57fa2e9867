// CUDA-core kernels around the tensor-core GEMMs of the DAnA forward path: the 7x7 stem with
// fused BN+ReLU+max-pool, support-side BA block and unary term, positional encoding,
// mean-centering, the attention softmax, RPN pair softmax, small pooling / conversion kernels.
// Every kernel cites the reference lines it restates.
#pragma once
#include "api_common.cuh"
#include "tc_common.cuh"

namespace dana {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float ld_pair(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long i) {
  float v = __bfloat162float(hi[i]);
  if (lo != nullptr) v += __bfloat162float(lo[i]);
  return v;
}
__device__ __forceinline__ void st_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, long long i, float v) {
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[i] = h;
  if (lo != nullptr) lo[i] = l;
}

// ---------------------------------------------------------------------------------------------
// Stem (lib/model/framework/resnet.py:109-113): conv1 7x7/2 pad 3 (3->64) + frozen BN + ReLU, then
// MaxPool 3x3/2 pad 0 ceil_mode.  The 7x7 stride-2 convolution on 3 channels is rewritten as a 4x4
// stride-1 convolution on the 2x2 space-to-depth image (12 -> 16 channels): four horizontally adjacent
// s2d pixels are 64 contiguous bf16 = one 128-byte TMA/UMMA K-block, so conv1 runs on the tensor-core
// implicit-GEMM kernel as a 4-tap (one per kernel row) K=256 GEMM with fused BN+ReLU.
//   s2d[b][j_y][2 + j_x][(sy*2+sx)*3 + c] = im[b][c][2*j_y+sy][2*j_x+sx]   (0 outside, channels 12..15 = 0)
//   rows are stored with 2 zero pixels on the left and >= 2 on the right so every 4-pixel window of an
//   output column lies inside its own row.
// ---------------------------------------------------------------------------------------------
// one thread per stored s2d pixel (incl. the zero border): 16 channels -> 32 B hi + 32 B lo
__global__ void stem_s2d_kernel(const float* __restrict__ in, int batch, int height, int width, int h2, int w2,
                                int wp, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long total = static_cast<long long>(batch) * h2 * wp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(i % wp);
    const int jy = static_cast<int>((i / wp) % h2);
    const int b = static_cast<int>(i / wp / h2);
    const int jx = xp - 2;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.0f;
    if (jx >= 0 && jx < w2) {
      const float* img = in + static_cast<long long>(b) * 3 * height * width;
#pragma unroll
      for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
          const int y = 2 * jy + sy, x = 2 * jx + sx;
          if (y < height && x < width) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              v[(sy * 2 + sx) * 3 + c] = __ldg(img + (static_cast<long long>(c) * height + y) * width + x);
          }
        }
    }
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * k], h0, l0);
      split_bf16(v[2 * k + 1], h1, l1);
      ph[k] = pack_bf16x2(h0, h1);
      pl[k] = pack_bf16x2(l0, l1);
    }
    uint4* oh = reinterpret_cast<uint4*>(out_hi + i * 16);
    oh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    oh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
    if (out_lo != nullptr) {
      uint4* ol = reinterpret_cast<uint4*>(out_lo + i * 16);
      ol[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      ol[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
  }
}

// MaxPool2d(3, stride 2, pad 0, ceil_mode) on an NHWC pair; thread per (pixel, 8-channel group)
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    int batch, int h, int w, int c, int oh, int ow,
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int cg = c / 8;
  const long long total = static_cast<long long>(batch) * oh * ow * cg;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % cg);
    const int ox = static_cast<int>((i / cg) % ow);
    const int oy = static_cast<int>((i / cg / ow) % oh);
    const int b = static_cast<int>(i / cg / ow / oh);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * oy + dy;
      if (y >= h) break;
      for (int dx = 0; dx < 3; ++dx) {
        const int x = 2 * ox + dx;
        if (x >= w) break;
        const long long off = ((static_cast<long long>(b) * h + y) * w + x) * c + g * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + off));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        float v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[2 * e] = __uint_as_float(aw[e] << 16);
          v[2 * e + 1] = __uint_as_float(aw[e] & 0xFFFF0000u);
        }
        if (lo != nullptr) {
          const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
          const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] += __uint_as_float(lw[e] << 16);
            v[2 * e + 1] += __uint_as_float(lw[e] & 0xFFFF0000u);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
      }
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(m[2 * e], h0, l0);
      split_bf16(m[2 * e + 1], h1, l1);
      ph[e] = pack_bf16x2(h0, h1);
      pl[e] = pack_bf16x2(l0, l1);
    }
    const long long o = ((static_cast<long long>(b) * oh + oy) * ow + ox) * c + g * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// AvgPool2d(k, stride 1) on NHWC bf16 pairs -> fp32 NHWC (dana.py:42,114: 20x20 -> 7x7, k = 14)
// ---------------------------------------------------------------------------------------------
// grid (maps, c/32), block (32, 8): the map's 32-channel slab is staged in shared memory once, then pooled
// separably (row sums over k columns, then k row-sum rows) -- h*w*(1 + ...) loads instead of oh*ow*k*k.
__global__ void __launch_bounds__(256)
avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int maps, int h, int w,
                    int c, int k, float* __restrict__ out) {
  extern __shared__ float s_pool[];
  const int oh = h - k + 1, ow = w - k + 1;
  float* s_in = s_pool;                 // [h*w][32]
  float* s_rs = s_pool + h * w * 32;    // [h][ow][32]
  const int m = blockIdx.x;
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const bool ok = ch < c;
  const long long base = static_cast<long long>(m) * h * w * c + ch;
  for (int p = threadIdx.y; p < h * w; p += blockDim.y)
    s_in[p * 32 + threadIdx.x] = ok ? ld_pair(hi, lo, base + static_cast<long long>(p) * c) : 0.0f;
  __syncthreads();
  for (int i = threadIdx.y; i < h * ow; i += blockDim.y) {
    const int y = i / ow, ox = i - y * ow;
    float s = 0.0f;
    for (int dx = 0; dx < k; ++dx) s += s_in[(y * w + ox + dx) * 32 + threadIdx.x];
    s_rs[i * 32 + threadIdx.x] = s;
  }
  __syncthreads();
  const float inv = 1.0f / static_cast<float>(k * k);
  for (int i = threadIdx.y; i < oh * ow; i += blockDim.y) {
    const int oy = i / ow, ox = i - oy * ow;
    float s = 0.0f;
    for (int dy = 0; dy < k; ++dy) s += s_rs[((oy + dy) * ow + ox) * 32 + threadIdx.x];
    if (ok) out[(static_cast<long long>(m) * oh * ow + i) * c + ch] = s * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// Support side of BA + CISA (dana.py:126-147 per shot; rcnn_head dana.py:255-276 with enhance off).
//   V  = S + PE                                              (:128)
//   w  = softmax_n(V c + c0);  g = w^T V;  V' = V + gamma * leaky_relu(g)      (:133-137)
//   u  = softmax_n(V' a + a0); r = unary_gamma * u^T V'       (:144-146 folded: adds r to every output row)
//   Vc = V' - mean_n V'  -> A operand of the k projection (the Linear bias cancels, :140-141)
//   V'^T -> B operand of the P.V contraction (:147)
// ---------------------------------------------------------------------------------------------
// Three passes over the support maps instead of five, from three identities (exact in real arithmetic):
//   * V' = V + 1 h^T with h = gamma * leaky_relu(g) a row-constant shift, so softmax_n(V' a + a0) = softmax_n(V a + a0):
//     both logit vectors (BA and unary) come from ONE pass over V;
//   * u^T V' = u^T V + h (the weights sum to one): g, the unary sum and the column mean come from ONE more pass;
//   * V' - mean_n V' = V - mean_n V: the centred k-projection operand does not depend on g at all.
// V itself (= S + PE) is never written back: every pass rebuilds it from the input pair (4 bytes per element, the
// same traffic as reading an fp32 copy, minus the copy's write).
//   pass 1  support_logits2_kernel    l1[row] = V c + c0,  l2[row] = V a + a0
//   pass 2  support_wsum2_kernel      w = softmax(l1), u = softmax(l2);  g = w^T V;  r = u^T V + h;  colmean
//   pass 3  support_finalize2_kernel  Vc = V - colmean (pair, A operand of the k projection),  V'^T = (V + h)^T (pair)
__device__ __forceinline__ float support_x(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const float* f32,
                                           const float* pe_row, long long i, int ch) {
  float v = f32 ? __ldg(f32 + i) : ld_pair(hi, lo, i);
  if (pe_row) v += __ldg(pe_row + ch);
  return v;
}
// 8 consecutive channels (c % 8 == 0, 16-byte aligned rows)
__device__ __forceinline__ void support_x8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const float* f32,
                                           const float* pe_row, long long i, int ch, float (&x)[8]) {
  if (f32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(f32 + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(f32 + i) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + i));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      x[2 * e] = __uint_as_float(hw[e] << 16);
      x[2 * e + 1] = __uint_as_float(hw[e] & 0xFFFF0000u);
    }
    if (lo) {
      const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + i));
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[2 * e] += __uint_as_float(lw[e] << 16);
        x[2 * e + 1] += __uint_as_float(lw[e] & 0xFFFF0000u);
      }
    }
  }
  if (pe_row) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(pe_row + ch));
    const float4 b = __ldg(reinterpret_cast<const float4*>(pe_row + ch) + 1);
    x[0] += a.x; x[1] += a.y; x[2] += a.z; x[3] += a.w; x[4] += b.x; x[5] += b.y; x[6] += b.z; x[7] += b.w;
  }
}

// warp per row (rows = maps * ns)
__global__ void __launch_bounds__(256)
support_logits2_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                       const float* __restrict__ in_f32, const float* __restrict__ pe, int rows, int ns, int c,
                       const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                       const float* __restrict__ b2, float* __restrict__ l1, float* __restrict__ l2, int vec8) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long base = static_cast<long long>(row) * c;
  const float* per = pe ? pe + static_cast<long long>(row % ns) * c : nullptr;
  float d1 = 0.0f, d2 = 0.0f;
  if (vec8) {   // c % 8 == 0 and every row pointer 16-byte aligned (checked by the launcher)
    for (int ch = lane * 8; ch < c; ch += 256) {
      float x[8];
      support_x8(in_hi, in_lo, in_f32, per, base + ch, ch, x);
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(w2 + ch)), a1 = __ldg(reinterpret_cast<const float4*>(w2 + ch) + 1);
      d2 += x[0] * a0.x + x[1] * a0.y + x[2] * a0.z + x[3] * a0.w + x[4] * a1.x + x[5] * a1.y + x[6] * a1.z + x[7] * a1.w;
      if (w1) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(w1 + ch)), c1 = __ldg(reinterpret_cast<const float4*>(w1 + ch) + 1);
        d1 += x[0] * c0.x + x[1] * c0.y + x[2] * c0.z + x[3] * c0.w + x[4] * c1.x + x[5] * c1.y + x[6] * c1.z + x[7] * c1.w;
      }
    }
  } else {
    for (int ch = lane; ch < c; ch += 32) {
      const float x = support_x(in_hi, in_lo, in_f32, per, base + ch, ch);
      d2 += x * __ldg(w2 + ch);
      if (w1) d1 += x * __ldg(w1 + ch);
    }
  }
  d2 = warp_sum(d2);
  if (w1) d1 = warp_sum(d1);
  if (lane == 0) {
    l2[row] = d2 + __ldg(b2);
    if (w1) l1[row] = d1 + __ldg(b1);
  }
}

// softmax weights of one map's logits into shared memory (all threads of the CTA); returns nothing, p[] normalised
__device__ __forceinline__ void cta_softmax_to_smem(const float* __restrict__ lg, int ns, float* s_p, float* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = tid; i < ns; i += blockDim.x) mx = fmaxf(mx, lg[i]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int i = 1; i < nw; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (int i = tid; i < ns; i += blockDim.x) {
    const float e = expf(lg[i] - mx);
    s_p[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
  for (int i = 0; i < nw; ++i) sum += s_red[i];
  const float inv = 1.0f / sum;
  __syncthreads();
  for (int i = tid; i < ns; i += blockDim.x) s_p[i] *= inv;
  __syncthreads();
}

// grid (maps, ceil(c / 64)), block 256: lane = 2 channels of the CTA's 64-channel slab, 8 warps stride the positions
__global__ void __launch_bounds__(256)
support_wsum2_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                     const float* __restrict__ in_f32, const float* __restrict__ pe, int ns, int c,
                     const float* __restrict__ l1, const float* __restrict__ l2, float gamma, float* __restrict__ g,
                     float* __restrict__ r, float* __restrict__ colmean) {
  extern __shared__ float s_dyn2[];     // p1[ns], p2[ns]
  __shared__ float s_red[32];
  __shared__ float s_acc[8][3][64];
  float* s_p1 = s_dyn2;
  float* s_p2 = s_dyn2 + ns;
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (l1) cta_softmax_to_smem(l1 + static_cast<long long>(m) * ns, ns, s_p1, s_red);
  cta_softmax_to_smem(l2 + static_cast<long long>(m) * ns, ns, s_p2, s_red);
  const int ch = blockIdx.y * 64 + lane * 2;
  float a1[2] = {0.f, 0.f}, a2[2] = {0.f, 0.f}, mean[2] = {0.f, 0.f};
  const long long mbase = static_cast<long long>(m) * ns * c;
  for (int n = warp; n < ns; n += 8) {
    const float* per = pe ? pe + static_cast<long long>(n) * c : nullptr;
    const float p1 = l1 ? s_p1[n] : 0.0f, p2 = s_p2[n];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (ch + e < c) {
        const float x = support_x(in_hi, in_lo, in_f32, per, mbase + static_cast<long long>(n) * c + ch + e, ch + e);
        a1[e] += p1 * x;
        a2[e] += p2 * x;
        mean[e] += x;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    s_acc[warp][0][lane * 2 + e] = a1[e];
    s_acc[warp][1][lane * 2 + e] = a2[e];
    s_acc[warp][2][lane * 2 + e] = mean[e];
  }
  __syncthreads();
  if (tid < 64 && blockIdx.y * 64 + tid < c) {
    float t1 = 0.f, t2 = 0.f, tm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t1 += s_acc[i][0][tid];
      t2 += s_acc[i][1][tid];
      tm += s_acc[i][2][tid];
    }
    const long long o = static_cast<long long>(m) * c + blockIdx.y * 64 + tid;
    float h = 0.0f;
    if (l1) {
      g[o] = t1;
      h = gamma * (t1 > 0.0f ? t1 : 0.01f * t1);
    }
    r[o] = t2 + h;                       // u^T V' = u^T V + h
    colmean[o] = tm / static_cast<float>(ns);
  }
}

// Same as support_wsum2_kernel with 8 channels per lane (c % 256 == 0, 16-byte aligned rows): grid (maps, c / 256),
// block 512 -- a warp reads a 512-byte run per plane and row instead of 128 bytes (the 2-channel version ran at
// 0.7 TB/s: 57 us for the 41 MB of the step's 24 support maps).
__global__ void __launch_bounds__(512)
support_wsum8_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                     const float* __restrict__ in_f32, const float* __restrict__ pe, int ns, int c,
                     const float* __restrict__ l1, const float* __restrict__ l2, float gamma, float* __restrict__ g,
                     float* __restrict__ r, float* __restrict__ colmean) {
  extern __shared__ float s_dyn8[];     // p1[ns], p2[ns], then acc[16][3][256]
  __shared__ float s_red[32];
  float* s_p1 = s_dyn8;
  float* s_p2 = s_dyn8 + ns;
  float* s_acc = s_dyn8 + 2 * ns;
  const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (l1) cta_softmax_to_smem(l1 + static_cast<long long>(m) * ns, ns, s_p1, s_red);
  cta_softmax_to_smem(l2 + static_cast<long long>(m) * ns, ns, s_p2, s_red);
  const int ch = blockIdx.y * 256 + lane * 8;
  float a1[8], a2[8], mean[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a1[e] = a2[e] = mean[e] = 0.0f;
  const long long mbase = static_cast<long long>(m) * ns * c;
  for (int n = warp; n < ns; n += 16) {
    const float p1 = l1 ? s_p1[n] : 0.0f, p2 = s_p2[n];
    float x[8];
    support_x8(in_hi, in_lo, in_f32, pe ? pe + static_cast<long long>(n) * c : nullptr,
               mbase + static_cast<long long>(n) * c + ch, ch, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a1[e] += p1 * x[e];
      a2[e] += p2 * x[e];
      mean[e] += x[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s_acc[(warp * 3 + 0) * 256 + lane * 8 + e] = a1[e];
    s_acc[(warp * 3 + 1) * 256 + lane * 8 + e] = a2[e];
    s_acc[(warp * 3 + 2) * 256 + lane * 8 + e] = mean[e];
  }
  __syncthreads();
  if (tid < 256) {
    float t1 = 0.f, t2 = 0.f, tm = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {       // fixed order
      t1 += s_acc[(i * 3 + 0) * 256 + tid];
      t2 += s_acc[(i * 3 + 1) * 256 + tid];
      tm += s_acc[(i * 3 + 2) * 256 + tid];
    }
    const long long o = static_cast<long long>(m) * c + blockIdx.y * 256 + tid;
    float h = 0.0f;
    if (l1) {
      g[o] = t1;
      h = gamma * (t1 > 0.0f ? t1 : 0.01f * t1);
    }
    r[o] = t2 + h;                       // u^T V' = u^T V + h
    colmean[o] = tm / static_cast<float>(ns);
  }
}

// Vc = V - colmean -> bf16 pair [rows][c];  V'^T = (V + h)^T -> bf16 pair vt[set][ch][slot * seg_pitch + n]
// grid (ceil(ns / 64), ceil(c / 64), maps), block 256; 64 x 64 tiles, two elements per lane on both sides so that
// every store is a 4-byte bf16 pair (128 bytes per warp and row)
__global__ void __launch_bounds__(256)
support_finalize2_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                         const float* __restrict__ in_f32, const float* __restrict__ pe,
                         const float* __restrict__ colmean, const float* __restrict__ g, float gamma, int ns, int c,
                         int shots, int seg_pitch, long long vt_pitch, __nv_bfloat16* __restrict__ vc_hi,
                         __nv_bfloat16* __restrict__ vc_lo, __nv_bfloat16* __restrict__ vt_hi,
                         __nv_bfloat16* __restrict__ vt_lo) {
  __shared__ float tile[64][65];
  const int m = blockIdx.z;
  const int n0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long mbase = static_cast<long long>(m) * ns * c;
  const bool vec = ((c & 1) == 0);
  for (int i = warp; i < 64; i += 8) {
    const int n = n0 + i;
    const int ch = c0 + lane * 2;
    float x[2] = {0.f, 0.f};
    if (n < ns) {
      const float* per = pe ? pe + static_cast<long long>(n) * c : nullptr;
      const long long o = mbase + static_cast<long long>(n) * c + ch;
      float d[2] = {0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (ch + e < c) {
          x[e] = support_x(in_hi, in_lo, in_f32, per, o + e, ch + e);
          d[e] = x[e] - colmean[static_cast<long long>(m) * c + ch + e];
        }
      }
      if (vec && ch + 1 < c) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(d[0], h0, l0);
        split_bf16(d[1], h1, l1);
        *reinterpret_cast<uint32_t*>(vc_hi + o) = pack_bf16x2(h0, h1);
        if (vc_lo) *reinterpret_cast<uint32_t*>(vc_lo + o) = pack_bf16x2(l0, l1);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (ch + e < c) st_pair(vc_hi, vc_lo, o + e, d[e]);
      }
    }
    tile[i][lane * 2] = x[0];
    tile[i][lane * 2 + 1] = x[1];
  }
  __syncthreads();
  const int set = m / shots, slot = m % shots;
  const bool vec_t = ((seg_pitch & 1) == 0) && ((vt_pitch & 1) == 0);
  for (int i = warp; i < 64; i += 8) {
    const int ch = c0 + i;
    if (ch >= c) continue;
    float h = 0.0f;
    if (g) {
      const float gg = g[static_cast<long long>(m) * c + ch];
      h = gamma * (gg > 0.0f ? gg : 0.01f * gg);
    }
    const int n = n0 + lane * 2;
    const long long o = (static_cast<long long>(set) * c + ch) * vt_pitch + static_cast<long long>(slot) * seg_pitch + n;
    const float v0 = tile[lane * 2][i] + h, v1 = tile[lane * 2 + 1][i] + h;
    if (vec_t && n + 1 < ns) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v0, h0, l0);
      split_bf16(v1, h1, l1);
      *reinterpret_cast<uint32_t*>(vt_hi + o) = pack_bf16x2(h0, h1);
      if (vt_lo) *reinterpret_cast<uint32_t*>(vt_lo + o) = pack_bf16x2(l0, l1);
    } else {
      if (n < ns) st_pair(vt_hi, vt_lo, o, v0);
      if (n + 1 < ns) st_pair(vt_hi, vt_lo, o + 1, v1);
    }
  }
}

// Key-major relayout of a per-position matrix: in fp32 [maps][ns][c] -> pair out[set][ch][slot * seg_pitch + n] with row
// pitch vt_pitch, zero in the pad columns (the B operand of a contraction over the keys of a support set; used for
// (V W^T)^T in the head, where the 2048->64 transform is applied to the support values BEFORE the attention-weighted
// sum -- (P V) W^T = P (V W^T), dana.py:281-288).  grid (sets, c), one output row per CTA; the data is a few hundred KB.
__global__ void transpose_segments_kernel(const float* __restrict__ in, int shots, int ns, int c, int seg_pitch,
                                          int vt_pitch, __nv_bfloat16* __restrict__ out_hi,
                                          __nv_bfloat16* __restrict__ out_lo) {
  const int set = blockIdx.x, ch = blockIdx.y;
  const long long orow = (static_cast<long long>(set) * c + ch) * vt_pitch;
  for (int j = threadIdx.x; j < vt_pitch; j += blockDim.x) {
    const int slot = j / seg_pitch, n = j - slot * seg_pitch;
    float v = 0.0f;
    if (slot < shots && n < ns) v = in[((static_cast<long long>(set) * shots + slot) * ns + n) * c + ch];
    st_pair(out_hi, out_lo, orow + j, v);
  }
}

// rbar[set][c] = (unary_gamma / shots) * sum_k r[set*shots + k][c]     (:146 + shot mean :150)
// optionally also cbar = rbar + mean_k colmean (as a bf16 pair): the row-constant part of the attended value once the
// values are mean-centred (engine.py, head: V = Vc + 1 m^T and the rows of P sum to one)
__global__ void support_rbar_kernel(const float* __restrict__ r, int sets, int shots, int c, float unary_gamma,
                                    float* __restrict__ rbar, const float* __restrict__ colmean,
                                    __nv_bfloat16* __restrict__ cbar_hi, __nv_bfloat16* __restrict__ cbar_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sets * c) return;
  const int set = i / c, ch = i - set * c;
  float s = 0.0f;
  for (int k = 0; k < shots; ++k) s += r[(static_cast<long long>(set) * shots + k) * c + ch];
  const float rb = s * unary_gamma / static_cast<float>(shots);
  rbar[i] = rb;
  if (cbar_hi != nullptr) {
    float m = 0.0f;
    for (int k = 0; k < shots; ++k) m += colmean[(static_cast<long long>(set) * shots + k) * c + ch];
    st_pair(cbar_hi, cbar_lo, i, rb + m / static_cast<float>(shots));
  }
}

// ---------------------------------------------------------------------------------------------
// Mean-centering over groups of rows (q/k projections, dana.py:125,141,267,272) -> bf16 pair.
// grid (groups, ceil(c/128)), block 128: thread per column, two passes over the group's rows.
// ---------------------------------------------------------------------------------------------
// (in_pitch: elements between input rows -- the input may be a column slice of a wider matrix; the output is dense)
__global__ void center_rows_kernel(const float* __restrict__ in, long long in_pitch, int group_rows, int c,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int g = blockIdx.x;
  const int ch = blockIdx.y * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const long long ibase = static_cast<long long>(g) * group_rows * in_pitch + ch;
  const long long obase = static_cast<long long>(g) * group_rows * c + ch;
  float s = 0.0f;
  for (int r = 0; r < group_rows; ++r) s += in[ibase + static_cast<long long>(r) * in_pitch];
  const float mean = s / static_cast<float>(group_rows);
  for (int r = 0; r < group_rows; ++r)
    st_pair(hi, lo, obase + static_cast<long long>(r) * c, in[ibase + static_cast<long long>(r) * in_pitch] - mean);
}
// large groups (RPN level, thousands of rows): per-block partial column sums, combined in a fixed order (no float
// atomics: the forward is bit-reproducible run to run), then subtract
__global__ void colsum_partial_kernel(const float* __restrict__ in, int group_rows, int c, int rows_per_block,
                                      float* __restrict__ partial /*[gridDim.x][groups][c]*/) {
  const int g = blockIdx.z;
  const int ch = blockIdx.y * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(group_rows, r0 + rows_per_block);
  const long long base = static_cast<long long>(g) * group_rows * c + ch;
  float s = 0.0f;
  for (int r = r0; r < r1; ++r) s += in[base + static_cast<long long>(r) * c];
  partial[(static_cast<long long>(blockIdx.x) * gridDim.z + g) * c + ch] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblk, int groups, int c,
                                    float* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * c) return;
  float s = 0.0f;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<long long>(b) * groups * c + i];
  sums[i] = s;
}
__global__ void center_apply_kernel(const float* __restrict__ in, const float* __restrict__ sums, int group_rows, int c,
                                    long long total, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const long long row = i / c;
    const long long g = row / group_rows;
    st_pair(hi, lo, i, in[i] - sums[g * c + ch] / static_cast<float>(group_rows));
  }
}

// ---------------------------------------------------------------------------------------------
// Attention softmax (dana.py:142-143,273-274): logits [rows][pitch] fp32 (already scaled by
// 1/sqrt(d) in the GEMM epilogue), `segs` segments of `ns` columns per row -> P bf16 pair, pad = 0.
// warp per row.
// ---------------------------------------------------------------------------------------------
// One warp per (row, segment): the segment lives in registers (<= 16 values per lane, ns <= 512), one expf per
// element, warp-shuffle max / sum, coalesced lane-strided loads and bf16 stores.
__global__ void __launch_bounds__(256)
attn_softmax_kernel(const float* __restrict__ s, long long rows, int segs, int ns, int pitch,
                    __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo) {
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= rows * segs) return;
  const long long row = wid / segs;
  const int k = static_cast<int>(wid - row * segs);
  const long long base = row * pitch + static_cast<long long>(k) * ns;
  float v[16];
  float mx = -INFINITY;
  const int nt = (ns + 31) >> 5;   // live register slots (uniform)
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (t < nt) {
      const int j = lane + 32 * t;
      v[t] = (j < ns) ? s[base + j] : -INFINITY;
      mx = fmaxf(mx, v[t]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (t < nt) {
      v[t] = (lane + 32 * t < ns) ? expf(v[t] - mx) : 0.0f;
      sum += v[t];
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    if (t < nt) {
      const int j = lane + 32 * t;
      if (j < ns) st_pair(p_hi, p_lo, base + j, v[t] * inv);
    }
  }
  if (k == segs - 1)   // zero the row's pad columns
    for (int j = segs * ns + lane; j < pitch; j += 32) st_pair(p_hi, p_lo, row * pitch + j, 0.0f);
}

// Same, four consecutive keys per lane (ns % 4 == 0, pitch % 4 == 0, 16-byte aligned planes): 16-byte logit loads and
// 8-byte bf16 stores per plane -- the lane-strided version above moves 2 bytes per lane and store instruction
// (measured 49 us for the 46 MB of logits of one step at Ns = 400).
__global__ void __launch_bounds__(256)
attn_softmax4_kernel(const float* __restrict__ s, long long rows, int segs, int ns, int pitch,
                     __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo) {
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= rows * segs) return;
  const long long row = wid / segs;
  const int k = static_cast<int>(wid - row * segs);
  const long long base = row * pitch + static_cast<long long>(k) * ns;
  float4 v[4];
  float mx = -INFINITY;
  const int nt = (ns + 127) >> 7;   // live register slots (uniform), ns <= 512
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < nt) {
      const int j = 4 * (lane + 32 * t);
      v[t] = (j < ns) ? __ldg(reinterpret_cast<const float4*>(s + base + j))
                      : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      mx = fmaxf(fmaxf(mx, fmaxf(v[t].x, v[t].y)), fmaxf(v[t].z, v[t].w));
    }
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < nt) {
      const bool ok = 4 * (lane + 32 * t) < ns;
      v[t].x = ok ? expf(v[t].x - mx) : 0.0f;
      v[t].y = ok ? expf(v[t].y - mx) : 0.0f;
      v[t].z = ok ? expf(v[t].z - mx) : 0.0f;
      v[t].w = ok ? expf(v[t].w - mx) : 0.0f;
      sum += (v[t].x + v[t].y) + (v[t].z + v[t].w);
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < nt) {
      const int j = 4 * (lane + 32 * t);
      if (j < ns) {
        const float f[4] = {v[t].x * inv, v[t].y * inv, v[t].z * inv, v[t].w * inv};
        uint32_t ph[2], pl[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
          ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - __uint_as_float(ph[e] << 16),
                                                          f[2 * e + 1] - __uint_as_float(ph[e] & 0xFFFF0000u));
          pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        *reinterpret_cast<uint2*>(p_hi + base + j) = make_uint2(ph[0], ph[1]);
        if (p_lo != nullptr) *reinterpret_cast<uint2*>(p_lo + base + j) = make_uint2(pl[0], pl[1]);
      }
    }
  }
  if (k == segs - 1)   // zero the row's pad columns
    for (int j = segs * ns + lane; j < pitch; j += 32) st_pair(p_hi, p_lo, row * pitch + j, 0.0f);
}

// ---------------------------------------------------------------------------------------------
// RPN pair softmax + delta repack (rpn.py:47-72, proposal_layer.py:67,97-103).
// in: [pixels][2A + 4A] fp32 NHWC (cls scores first, channel a = bg, A + a = fg).
// out: fg [pixels*A], deltas [pixels*A][4]
// ---------------------------------------------------------------------------------------------
__global__ void rpn_fg_prob_kernel(const float* __restrict__ in, long long pixels, int num_a, int pitch,
                                   float* __restrict__ fg, float4* __restrict__ deltas) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= pixels * num_a) return;
  const long long px = i / num_a;
  const int a = static_cast<int>(i - px * num_a);
  const float* row = in + px * pitch;
  const float s0 = row[a], s1 = row[num_a + a];
  const float m = fmaxf(s0, s1);
  const float e0 = expf(s0 - m), e1 = expf(s1 - m);
  fg[i] = e1 / (e0 + e1);
  const float* d = row + 2 * num_a + a * 4;
  deltas[i] = make_float4(d[0], d[1], d[2], d[3]);
}

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// out pair [rows][c_pitch] (+col offset) = in fp32 [rows][c] + pe[row % period][c]
__global__ void add_pe_split_kernel(const float* __restrict__ in, const float* __restrict__ pe, long long rows, int c,
                                    int period, long long out_pitch, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo) {
  const long long total = rows * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / c;
    const int ch = static_cast<int>(i - row * c);
    float v = in[i];
    if (pe) v += pe[static_cast<long long>(row % period) * c + ch];
    st_pair(hi, lo, row * out_pitch + ch, v);
  }
}
// fp32 -> bf16 pair, same shape
__global__ void split_f32_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    st_pair(hi, lo, i, in[i]);
}
// bf16 pair [rows][c] with row pitch in_pitch -> fp32 contiguous [rows][c] and / or one fp16 plane with row pitch
// f16_pitch (the query half of the fp16 RPN input).  8 channels per thread when VEC (16-byte accesses).
template <bool VEC>
__global__ void merge_pair_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                  long long n, int c, long long in_pitch, float* __restrict__ out,
                                  __half* __restrict__ out16, long long f16_pitch) {
  if constexpr (VEC) {
    const int cg = c >> 3;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n / 8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long row = i / cg;
      const int ch = static_cast<int>(i - row * cg) * 8;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + row * in_pitch + ch));
      const uint32_t wh[4] = {h.x, h.y, h.z, h.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f[2 * e] = __uint_as_float(wh[e] << 16);
        f[2 * e + 1] = __uint_as_float(wh[e] & 0xFFFF0000u);
      }
      if (lo != nullptr) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + row * in_pitch + ch));
        const uint32_t wl[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          f[2 * e] += __uint_as_float(wl[e] << 16);
          f[2 * e + 1] += __uint_as_float(wl[e] & 0xFFFF0000u);
        }
      }
      if (out != nullptr) {
        float4* o = reinterpret_cast<float4*>(out + row * c + ch);
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
      if (out16 != nullptr) {
        uint32_t ph[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 h2 = __floats2half2_rn(fminf(fmaxf(f[2 * e], -65504.0f), 65504.0f),
                                               fminf(fmaxf(f[2 * e + 1], -65504.0f), 65504.0f));
          ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(out16 + row * f16_pitch + ch) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      }
    }
  } else {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long row = i / c;
      const float v = ld_pair(hi, lo, row * in_pitch + (i - row * c));
      if (out != nullptr) out[i] = v;
      if (out16 != nullptr) out16[row * f16_pitch + (i - row * c)] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
    }
  }
}
// mean over `sp` spatial positions: in [items][sp][c] (bf16 pair, or one fp16 plane when F16) -> fp32 [items][c] and
// pair; 8 channels per thread (c % 8 == 0, 16-byte aligned planes: checked by the launcher)
template <bool F16>
__global__ void spatial_mean_kernel(const void* __restrict__ hi_v, const __nv_bfloat16* __restrict__ lo,
                                    long long items, int sp, int c, float* __restrict__ out,
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int cg = c >> 3;
  const long long total = items * cg;
  const __nv_bfloat16* hi = static_cast<const __nv_bfloat16*>(hi_v);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long it = i / cg;
    const int ch = static_cast<int>(i - it * cg) * 8;
    // reference: .mean(3).mean(2) -- mean over w first, then over h (sp = h*w, square)
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int p = 0; p < sp; ++p) {
      const long long off = (it * sp + p) * c + ch;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi + off));
      const uint32_t wh[4] = {h.x, h.y, h.z, h.w};
      if constexpr (F16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&wh[e]));
          s[2 * e] += v.x;
          s[2 * e + 1] += v.y;
        }
      } else {
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          f[2 * e] = __uint_as_float(wh[e] << 16);
          f[2 * e + 1] = __uint_as_float(wh[e] & 0xFFFF0000u);
        }
        if (lo != nullptr) {
          const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo + off));
          const uint32_t wl[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[2 * e] += __uint_as_float(wl[e] << 16);
            f[2 * e + 1] += __uint_as_float(wl[e] & 0xFFFF0000u);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += f[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float m = s[e] / static_cast<float>(sp);
      if (out) out[it * c + ch + e] = m;
      if (out_hi) st_pair(out_hi, out_lo, it * c + ch + e, m);
    }
  }
}
// 2-class softmax over rows of [rows][2]
__global__ void softmax2_kernel(const float* __restrict__ in, long long rows, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float a = in[2 * i], b = in[2 * i + 1];
  const float m = fmaxf(a, b);
  const float ea = expf(a - m), eb = expf(b - m);
  out[2 * i] = ea / (ea + eb);
  out[2 * i + 1] = eb / (ea + eb);
}
// NHWC pair -> NCHW fp32 (boundary export of feature maps)
__global__ void nhwc_pair_to_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                         int c, int hw, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const long long base = static_cast<long long>(b) * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && ch < c) ? ld_pair(hi, lo, base + static_cast<long long>(p) * c + ch) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    if (ch < c && p < hw) out[base + static_cast<long long>(ch) * hw + p] = tile[threadIdx.x][i];
  }
}

inline int grid_for(long long n, int tb) {
  long long g = (n + tb - 1) / tb;
  const long long cap = 148LL * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dana
