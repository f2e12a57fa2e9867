// libdana_b200.so, second translation unit: the layout kernels of the backward pass (SURVEY.md section 8 row a15,
// BASELINE configs[3]) and the fused SGD update.  The gradient GEMMs themselves run on the forward's tcgen05 kernel
// (dana_conv_gemm): the data-gradient of a convolution is the same implicit GEMM with transposed / rotated weights, the
// weight-gradient  dW[co][tap][ci] = sum_p g[p][co] * x[p + tap][ci]  is a plain K-major GEMM over the PIXEL index once
// both operands are laid out channel-major -- which is what the two kernels below produce, straight from the NHWC
// tensors, in the bf16 (hi, lo) operand format.
//
// Replaces, for the training step: the cuDNN backward-data / backward-filter calls under torch.autograd of
// lib/model/framework/resnet.py:66-102, dana.py:120-151,244-290 (train.py:138 loss.backward()) and torch.optim.SGD.step
// (train.py:89,139).
#include <cuda_bf16.h>
#include <string.h>

#include "api_common.cuh"

namespace dana {
namespace {

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

constexpr int kTP = 64;        // pixels per tile
constexpr int kTC = 64;        // channels per tile
constexpr int kTPitch = 72;    // smem row pitch in elements (144 B: rows stay 16-byte aligned, banks spread)

// Phase B of both kernels: the [channel][pixel] tile in shared memory goes out as rows of 64 consecutive pixels.
__device__ __forceinline__ void store_tile_t(const __nv_bfloat16 (*th)[kTPitch], const __nv_bfloat16 (*tl)[kTPitch],
                                             __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, long long t_pitch, int c0,
                                             int channels, long long p0, long long pixels) {
  const int ch = threadIdx.x >> 2;            // 0..63
  const int seg = (threadIdx.x & 3) * 16;     // 16 pixels = 32 B per plane
  if (c0 + ch >= channels) return;
  __nv_bfloat16* oh = out_hi + static_cast<long long>(c0 + ch) * t_pitch + p0 + seg;
  __nv_bfloat16* ol = out_lo ? out_lo + static_cast<long long>(c0 + ch) * t_pitch + p0 + seg : nullptr;
  if (p0 + seg + 16 <= pixels) {
    const uint4* sh = reinterpret_cast<const uint4*>(&th[ch][seg]);
    reinterpret_cast<uint4*>(oh)[0] = sh[0];
    reinterpret_cast<uint4*>(oh)[1] = sh[1];
    if (ol) {
      const uint4* sl = reinterpret_cast<const uint4*>(&tl[ch][seg]);
      reinterpret_cast<uint4*>(ol)[0] = sl[0];
      reinterpret_cast<uint4*>(ol)[1] = sl[1];
    }
  } else {
    for (int i = 0; i < 16 && p0 + seg + i < pixels; ++i) {
      oh[i] = th[ch][seg + i];
      if (ol) ol[i] = tl[ch][seg + i];
    }
  }
}

// g' = g * (y > 0) (mask optional) on an fp32 [pixels][channels] gradient; writes any of: the masked fp32 gradient (the
// residual branch's share), its NHWC bf16 pair (operand of the data-gradient GEMM) and the channel-major pair
// [channels][t_pitch] (operand of the weight-gradient GEMM).
__global__ void __launch_bounds__(256) grad_prepare_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                           long long pixels, int channels, float* __restrict__ out_f32,
                                                           __nv_bfloat16* __restrict__ nh, __nv_bfloat16* __restrict__ nl,
                                                           __nv_bfloat16* __restrict__ th_out,
                                                           __nv_bfloat16* __restrict__ tl_out, long long t_pitch,
                                                           const bool vec) {
  __shared__ __align__(16) __nv_bfloat16 th[kTC][kTPitch];
  __shared__ __align__(16) __nv_bfloat16 tl[kTC][kTPitch];
  const long long p0 = static_cast<long long>(blockIdx.x) * kTP;
  const int c0 = blockIdx.y * kTC;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + 256 * i;
    const int row = idx >> 4, c4 = (idx & 15) * 4;
    const long long p = p0 + row;
    const int c = c0 + c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ok = p < pixels && c < channels;
    __align__(8) __nv_bfloat16 h[4], l[4];
    if (vec) {                                       // channels % 4 == 0 and 16-byte aligned rows
      if (ok) {
        v = *reinterpret_cast<const float4*>(g + p * channels + c);
        if (y != nullptr) {
          const float4 m = *reinterpret_cast<const float4*>(y + p * channels + c);
          v.x = m.x > 0.f ? v.x : 0.f;
          v.y = m.y > 0.f ? v.y : 0.f;
          v.z = m.z > 0.f ? v.z : 0.f;
          v.w = m.w > 0.f ? v.w : 0.f;
        }
        if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + p * channels + c) = v;
      }
    } else if (ok) {                                 // any channel count (the C -> 1 layers): element by element
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < 4 && c + k < channels; ++k) {
        const long long o = p * channels + c + k;
        e[k] = (y == nullptr || y[o] > 0.f) ? g[o] : 0.f;
        if (out_f32 != nullptr) out_f32[o] = e[k];
      }
      v = make_float4(e[0], e[1], e[2], e[3]);
    }
    split2(v.x, h[0], l[0]);
    split2(v.y, h[1], l[1]);
    split2(v.z, h[2], l[2]);
    split2(v.w, h[3], l[3]);
    if (ok && nh != nullptr) {
      if (vec) {
        *reinterpret_cast<uint2*>(nh + p * channels + c) = *reinterpret_cast<const uint2*>(h);
        if (nl != nullptr) *reinterpret_cast<uint2*>(nl + p * channels + c) = *reinterpret_cast<const uint2*>(l);
      } else {
        for (int k = 0; k < 4 && c + k < channels; ++k) {
          nh[p * channels + c + k] = h[k];
          if (nl != nullptr) nl[p * channels + c + k] = l[k];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      th[c4 + e][row] = h[e];
      tl[c4 + e][row] = l[e];
    }
  }
  if (th_out == nullptr) return;
  __syncthreads();
  store_tile_t(th, tl, th_out, tl_out, t_pitch, c0, channels, p0, pixels);
}

// Channel-major im2col of an NHWC bf16 pair: out[(tap * C + c)][(n * OH + oy) * OW + ox] = x[n, oy*s + r - pad,
// ox*s + t - pad, c] (zero outside), tap = r * S + t.  1 x 1 with s = 2 is the strided first convolution of a stage.
struct Im2colT {
  const __nv_bfloat16 *xh, *xl;
  long long sn, sy, sx;      // element strides of x (channel stride 1)
  int n, h, w, c;
  int oh, ow, stride, pad, taps_s;
  __nv_bfloat16 *oh_t, *ol_t;
  long long t_pitch;
};

__global__ void __launch_bounds__(256) im2col_t_kernel(const Im2colT a) {
  __shared__ __align__(16) __nv_bfloat16 th[kTC][kTPitch];
  __shared__ __align__(16) __nv_bfloat16 tl[kTC][kTPitch];
  const long long pixels = static_cast<long long>(a.n) * a.oh * a.ow;
  const long long p0 = static_cast<long long>(blockIdx.x) * kTP;
  const int c0 = blockIdx.y * kTC;
  const int tap = blockIdx.z;
  const int r = tap / a.taps_s, t = tap % a.taps_s;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = threadIdx.x + 256 * i;       // 64 rows x 8 groups of 8 channels
    const int row = idx >> 3, c8 = (idx & 7) * 8;
    const long long p = p0 + row;
    uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
    if (p < pixels && c0 + c8 < a.c) {            // channels % 8 == 0
      const int ox = static_cast<int>(p % a.ow);
      const long long q = p / a.ow;
      const int oy = static_cast<int>(q % a.oh);
      const int n = static_cast<int>(q / a.oh);
      const int iy = oy * a.stride + r - a.pad, ix = ox * a.stride + t - a.pad;
      if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
        const long long off = n * a.sn + iy * a.sy + ix * a.sx + c0 + c8;
        vh = *reinterpret_cast<const uint4*>(a.xh + off);
        if (a.xl != nullptr) vl = *reinterpret_cast<const uint4*>(a.xl + off);
      }
    }
    const __nv_bfloat16* eh = reinterpret_cast<const __nv_bfloat16*>(&vh);
    const __nv_bfloat16* el = reinterpret_cast<const __nv_bfloat16*>(&vl);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      th[c8 + e][row] = eh[e];
      tl[c8 + e][row] = el[e];
    }
  }
  __syncthreads();
  const long long plane = static_cast<long long>(tap) * a.c * a.t_pitch;
  store_tile_t(th, tl, a.oh_t + plane, a.ol_t ? a.ol_t + plane : nullptr, a.t_pitch, c0, a.c, p0, pixels);
}

// torch.optim.SGD with momentum (train.py:89): d = g * grad_scale + wd * p; m = mu * m + d; p -= lr * m
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    long long n, float lr, float mu, float wd, float grad_scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      mv.x = mu * mv.x + gv.x * grad_scale + wd * pv.x;
      mv.y = mu * mv.y + gv.y * grad_scale + wd * pv.y;
      mv.z = mu * mv.z + gv.z * grad_scale + wd * pv.z;
      mv.w = mu * mv.w + gv.w * grad_scale + wd * pv.w;
      pv.x -= lr * mv.x;
      pv.y -= lr * mv.y;
      pv.z -= lr * mv.z;
      pv.w -= lr * mv.w;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(p + i) = pv;
    } else {
      for (long long j = i; j < n; ++j) {
        const float mv = mu * m[j] + g[j] * grad_scale + wd * p[j];
        m[j] = mv;
        p[j] -= lr * mv;
      }
    }
  }
}

// Operand planes of a convolution weight w[co][ci][r][s] (fp32, frozen-BN scale folded in) for both GEMMs that read it:
//   forward / weight-gradient layout  f[co][(r*kw + s)*ci_n + ci] = w[co][ci][r][s] * scale[co]
//   data-gradient layout              d[ci][(r*kw + s)*co_n + co] = w[co][ci][kh-1-r][kw-1-s] * scale[co]   (rotated kernel)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, int co_n, int ci_n,
                                        int kh, int kw, __nv_bfloat16* __restrict__ fh, __nv_bfloat16* __restrict__ fl,
                                        __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl) {
  const long long total = static_cast<long long>(co_n) * ci_n * kh * kw;
  const int taps = kh * kw;
  const bool dgrad = blockIdx.y == 1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int co, ci, tap;
    if (!dgrad) {
      ci = static_cast<int>(i % ci_n);
      tap = static_cast<int>((i / ci_n) % taps);
      co = static_cast<int>(i / (static_cast<long long>(ci_n) * taps));
    } else {
      co = static_cast<int>(i % co_n);
      tap = taps - 1 - static_cast<int>((i / co_n) % taps);          // 180-degree rotation
      ci = static_cast<int>(i / (static_cast<long long>(co_n) * taps));
    }
    float v = w[(static_cast<long long>(co) * ci_n + ci) * taps + tap];
    if (scale != nullptr) v *= scale[co];
    __nv_bfloat16 h, l;
    split2(v, h, l);
    if (!dgrad) {
      fh[i] = h;
      fl[i] = l;
    } else {
      dh[i] = h;
      dl[i] = l;
    }
  }
}

// dW[co][ci][r][s] = g[co][(r*kw + s)*ci_n + ci] * scale[co]: the weight-gradient GEMM's output back in the parameter's layout
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ scale, int co_n, int ci_n,
                                         int taps, float* __restrict__ dw, const bool accumulate) {
  const long long total = static_cast<long long>(co_n) * ci_n * taps;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int ci = static_cast<int>((i / taps) % ci_n);
    const int co = static_cast<int>(i / (static_cast<long long>(taps) * ci_n));
    float v = g[(static_cast<long long>(co) * taps + tap) * ci_n + ci];
    if (scale != nullptr) v *= scale[co];
    // accumulate: a red.add -- the query trunk's and the support trunk's backward run on two streams and may write the
    // same parameter's gradient at the same time (two commutative contributions onto a zeroed arena: deterministic)
    if (accumulate) atomicAdd(dw + i, v); else dw[i] = v;
  }
}

inline bool al(const void* q, int a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; }

}  // namespace
}  // namespace dana

using namespace dana;

extern "C" {

int dana_grad_prepare(const float* grad, const float* relu_out, int64_t pixels, int channels, float* out_f32,
                      void* out_hi, void* out_lo, void* t_hi, void* t_lo, int64_t t_pitch, void* stream) {
  if (!grad || pixels <= 0 || channels <= 0) return DANA_EINVAL;
  if (!out_f32 && !out_hi && !t_hi) return DANA_EINVAL;
  if ((out_lo && !out_hi) || (t_lo && !t_hi)) return DANA_EINVAL;
  if (t_hi && (t_pitch < pixels || (t_pitch % 8))) return DANA_EINVAL;
  if (!al(t_hi, 16) || !al(t_lo, 16)) return DANA_EINVAL;
  const bool vec = (channels % 4) == 0 && al(grad, 16) && al(relu_out, 16) && al(out_f32, 16) && al(out_hi, 8) && al(out_lo, 8);
  const dim3 grid(static_cast<unsigned>((pixels + kTP - 1) / kTP), static_cast<unsigned>((channels + kTC - 1) / kTC));
  grad_prepare_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad, relu_out, pixels, channels, out_f32, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo),
      static_cast<__nv_bfloat16*>(t_hi), static_cast<__nv_bfloat16*>(t_lo), t_pitch, vec);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

int dana_im2col_t(const void* x_hi, const void* x_lo, int batch, int height, int width, int channels, int64_t stride_n,
                  int64_t stride_y, int64_t stride_x, int ksize, int conv_stride, void* t_hi, void* t_lo, int64_t t_pitch,
                  void* stream) {
  if (!x_hi || !t_hi || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || (channels % 8)) return DANA_EINVAL;
  if (!((ksize == 1 && conv_stride >= 1) || (ksize == 3 && conv_stride == 1))) return DANA_ENOTSUP;
  if ((stride_n % 8) || (stride_y % 8) || (stride_x % 8)) return DANA_EINVAL;
  if (!al(x_hi, 16) || !al(x_lo, 16) || !al(t_hi, 16) || !al(t_lo, 16) || (x_lo != nullptr) != (t_lo != nullptr)) return DANA_EINVAL;
  Im2colT a;
  a.xh = static_cast<const __nv_bfloat16*>(x_hi);
  a.xl = static_cast<const __nv_bfloat16*>(x_lo);
  a.sn = stride_n, a.sy = stride_y, a.sx = stride_x;
  a.n = batch, a.h = height, a.w = width, a.c = channels;
  a.stride = conv_stride;
  a.pad = ksize == 3 ? 1 : 0;
  a.taps_s = ksize;
  a.oh = ksize == 3 ? height : (height - 1) / conv_stride + 1;
  a.ow = ksize == 3 ? width : (width - 1) / conv_stride + 1;
  a.oh_t = static_cast<__nv_bfloat16*>(t_hi);
  a.ol_t = static_cast<__nv_bfloat16*>(t_lo);
  a.t_pitch = t_pitch;
  const long long pixels = static_cast<long long>(batch) * a.oh * a.ow;
  if (t_pitch < pixels || (t_pitch % 8)) return DANA_EINVAL;
  const dim3 grid(static_cast<unsigned>((pixels + kTP - 1) / kTP), static_cast<unsigned>((channels + kTC - 1) / kTC),
                  static_cast<unsigned>(ksize * ksize));
  im2col_t_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

int dana_pack_conv_weight(const float* weight, const float* scale, int out_channels, int in_channels, int kh, int kw,
                          void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* dgrad_lo, void* stream) {
  if (!weight || !fwd_hi || !fwd_lo || out_channels <= 0 || in_channels <= 0 || kh <= 0 || kw <= 0) return DANA_EINVAL;
  if ((dgrad_hi == nullptr) != (dgrad_lo == nullptr)) return DANA_EINVAL;
  const long long total = static_cast<long long>(out_channels) * in_channels * kh * kw;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks), dgrad_hi ? 2 : 1);
  pack_conv_weight_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weight, scale, out_channels, in_channels, kh, kw, static_cast<__nv_bfloat16*>(fwd_hi),
      static_cast<__nv_bfloat16*>(fwd_lo), static_cast<__nv_bfloat16*>(dgrad_hi), static_cast<__nv_bfloat16*>(dgrad_lo));
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

int dana_unpack_conv_wgrad(const float* wgrad, const float* scale, int out_channels, int in_channels, int taps,
                           float* weight_grad, int accumulate, void* stream) {
  if (!wgrad || !weight_grad || out_channels <= 0 || in_channels <= 0 || taps <= 0) return DANA_EINVAL;
  const long long total = static_cast<long long>(out_channels) * in_channels * taps;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  unpack_conv_wgrad_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      wgrad, scale, out_channels, in_channels, taps, weight_grad, accumulate != 0);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

// The whole backward of one convolution (+ frozen-BN fold + ReLU) in one call: mask / split / transpose of the incoming
// gradient, data-gradient GEMM, transposing im2col of the saved input, weight-gradient GEMM, unpack (+ accumulate).
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t dana_conv_backward_workspace_bytes(int batch, int height, int width, int in_channels, int out_channels, int ksize,
                                           int stride) {
  if (batch <= 0 || height <= 0 || width <= 0 || in_channels <= 0 || out_channels <= 0 || ksize <= 0 || stride <= 0) return 256;
  const int64_t oh = ksize == 3 ? height : (height - 1) / stride + 1, ow = ksize == 3 ? width : (width - 1) / stride + 1;
  const int64_t pixels = static_cast<int64_t>(batch) * oh * ow, pitch = (pixels + 7) / 8 * 8, taps = ksize * ksize;
  return 2 * align256(2 * pixels * out_channels) + 2 * align256(2 * out_channels * pitch) +
         2 * align256(2 * taps * in_channels * pitch) + align256(4 * out_channels * taps * in_channels) + 256;
}

int dana_conv_backward(const dana_conv_bwd_args* a, void* stream) {
  if (!a || !a->grad_out) return DANA_EINVAL;
  if (a->batch <= 0 || a->height <= 0 || a->width <= 0 || a->in_channels <= 0 || a->out_channels <= 0) return DANA_EINVAL;
  if (!((a->ksize == 1 && a->stride >= 1) || (a->ksize == 3 && a->stride == 1))) return DANA_ENOTSUP;
  const int n = a->batch, ci = a->in_channels, co = a->out_channels, ks = a->ksize, taps = ks * ks;
  const int oh = ks == 3 ? a->height : (a->height - 1) / a->stride + 1, ow = ks == 3 ? a->width : (a->width - 1) / a->stride + 1;
  const int64_t pixels = static_cast<int64_t>(n) * oh * ow, pitch = (pixels + 7) / 8 * 8;
  const bool need_dx = a->dx != nullptr, need_dw = a->dw != nullptr;
  if (need_dx && (!a->wd_hi || !a->wd_lo)) return DANA_EINVAL;
  if (need_dw && (!a->x_hi || !a->x_lo)) return DANA_EINVAL;
  if (a->workspace_bytes < dana_conv_backward_workspace_bytes(n, a->height, a->width, ci, co, ks, a->stride) ||
      (reinterpret_cast<uintptr_t>(a->workspace) & 255))
    return DANA_EINVAL;
  char* w = static_cast<char*>(a->workspace);
  auto take = [&](int64_t bytes) { char* q = w; w += align256(bytes); return q; };
  void* gp_hi = take(2 * pixels * co);
  void* gp_lo = take(2 * pixels * co);
  void* gt_hi = take(2 * co * pitch);
  void* gt_lo = take(2 * co * pitch);
  void* xt_hi = take(2 * static_cast<int64_t>(taps) * ci * pitch);
  void* xt_lo = take(2 * static_cast<int64_t>(taps) * ci * pitch);
  float* dwk = reinterpret_cast<float*>(take(4LL * co * taps * ci));
  int rc = dana_grad_prepare(a->grad_out, a->relu_out, pixels, co, a->dres, need_dx ? gp_hi : nullptr,
                             need_dx ? gp_lo : nullptr, need_dw ? gt_hi : nullptr, need_dw ? gt_lo : nullptr, pitch, stream);
  if (rc != DANA_OK && (need_dx || need_dw || a->dres)) return rc;
  dana_conv_gemm_args g;
  if (need_dx) {
    // dx[p][ci] = sum_{tap,co} g'[p - tap][co] W[co][ci][tap]: the forward kernel on the rotated weight planes; a strided
    // 1x1 wrote only every stride-th input pixel, so its data-gradient lands on those pixels of a zeroed map
    if (a->stride > 1) {
      cudaError_t e = cudaMemsetAsync(a->dx, 0, 4LL * n * a->height * a->width * ci, static_cast<cudaStream_t>(stream));
      if (e != cudaSuccess) return cuda_fail(e);
    }
    memset(&g, 0, sizeof(g));
    g.a_hi = gp_hi, g.a_lo = gp_lo;
    g.a_c = co, g.a_w = ow, g.a_h = oh, g.a_n = n;
    g.a_sx = co, g.a_sy = static_cast<int64_t>(ow) * co, g.a_sn = static_cast<int64_t>(oh) * ow * co;
    g.taps_r = g.taps_s = ks;
    g.pad_y = g.pad_x = ks == 3 ? 1 : 0;
    g.b_hi = a->wd_hi, g.b_lo = a->wd_lo;
    g.b_pitch = static_cast<int64_t>(taps) * co;
    g.n_out = ci;
    g.out_w = ow, g.out_h = oh, g.out_n = n;
    g.o_sx = static_cast<int64_t>(a->stride) * ci;
    g.o_sy = static_cast<int64_t>(a->stride) * a->width * ci;
    g.o_sn = static_cast<int64_t>(a->height) * a->width * ci;
    g.out_f32 = a->dx;
    g.alpha = 1.0f;
    g.workspace = a->gemm_workspace, g.workspace_bytes = a->gemm_workspace_bytes, g.sk_epoch = a->sk_epoch;
    rc = dana_conv_gemm(&g, stream);
    if (rc != DANA_OK) return rc;
  }
  if (need_dw) {
    rc = dana_im2col_t(a->x_hi, a->x_lo, n, a->height, a->width, ci, a->x_sn, a->x_sy, a->x_sx, ks, a->stride, xt_hi, xt_lo,
                       pitch, stream);
    if (rc != DANA_OK) return rc;
    // dW'[co][tap*ci_n + ci] = sum_p g'^T[co][p] * x^T[tap*ci_n + ci][p]: rows = co, K = pixels
    memset(&g, 0, sizeof(g));
    g.a_hi = gt_hi, g.a_lo = gt_lo;
    g.a_c = pixels, g.a_w = co, g.a_h = 1, g.a_n = 1;
    g.a_sx = pitch, g.a_sy = pitch * co, g.a_sn = pitch * co;
    g.taps_r = g.taps_s = 1;
    g.b_hi = xt_hi, g.b_lo = xt_lo;
    g.b_pitch = pitch;
    g.n_out = taps * ci;
    g.out_w = co, g.out_h = 1, g.out_n = 1;
    g.o_sx = static_cast<int64_t>(taps) * ci, g.o_sy = g.o_sx * co, g.o_sn = g.o_sy;
    g.out_f32 = dwk;
    g.alpha = 1.0f;
    g.workspace = a->gemm_workspace, g.workspace_bytes = a->gemm_workspace_bytes;
    g.sk_epoch = a->sk_epoch ? a->sk_epoch + 1 : 0;
    rc = dana_conv_gemm(&g, stream);
    if (rc != DANA_OK) return rc;
    rc = dana_unpack_conv_wgrad(dwk, a->scale, co, ci, taps, a->dw, a->dw_accumulate, stream);
    if (rc != DANA_OK) return rc;
  }
  return DANA_OK;
}

int dana_sgd_momentum(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                      float weight_decay, float grad_scale, void* stream) {
  if (!param || !grad || !momentum_buf || n <= 0) return DANA_EINVAL;
  if (!al(param, 16) || !al(grad, 16) || !al(momentum_buf, 16)) return DANA_EINVAL;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  sgd_momentum_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, momentum_buf, n, lr, momentum, weight_decay, grad_scale);
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // extern "C"
