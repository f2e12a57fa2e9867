// RoIAlign forward / backward (lib/model/csrc/cpu/ROIAlign_cpu.cpp:18-219 sampling rules:
// no coordinate rounding, no half-pixel shift, malformed RoIs forced to 1x1, adaptive
// ceil(roi/pooled) sampling grid when sampling_ratio <= 0, samples outside [-1, H] x [-1, W] are zero).
//
// Forward design.  Bilinear sampling on a regular grid followed by an average is separable:
//     out[ph][pw] = (1/count) * sum_y sum_x  wy[ph][y] * wx[pw][x] * F[y][x]
// with wy[ph][y] the summed row weights of the bin's samples (wx likewise).  One CTA per RoI
// (x channel group) builds the two small weight tables in shared memory once, then every thread
// owns a few consecutive channels of the NHWC map: all loads of a warp are one contiguous
// 128..512-byte run, each feature value is read O(1) times instead of 4 x grid^2 times, and the
// loop bounds are warp-uniform.  Sample coordinates are evaluated with explicit round-to-nearest
// intrinsics in the reference's expression order so floor()/validity decisions agree bit for bit.
#pragma once
#include <stdlib.h>

#include "api_common.cuh"
#include "tc_common.cuh"

namespace dana {

struct RoiGeom {
  int batch_ind;
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* roi, float spatial_scale, int pooled_h, int pooled_w,
                                                int sampling_ratio) {
  RoiGeom g;
  g.batch_ind = static_cast<int>(roi[0]);
  g.start_w = __fmul_rn(roi[1], spatial_scale);
  g.start_h = __fmul_rn(roi[2], spatial_scale);
  const float end_w = __fmul_rn(roi[3], spatial_scale);
  const float end_h = __fmul_rn(roi[4], spatial_scale);
  const float roi_w = fmaxf(__fsub_rn(end_w, g.start_w), 1.0f);
  const float roi_h = fmaxf(__fsub_rn(end_h, g.start_h), 1.0f);
  g.bin_h = __fdiv_rn(roi_h, static_cast<float>(pooled_h));
  g.bin_w = __fdiv_rn(roi_w, static_cast<float>(pooled_w));
  g.grid_h = (sampling_ratio > 0) ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(roi_h, static_cast<float>(pooled_h))));
  g.grid_w = (sampling_ratio > 0) ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(roi_w, static_cast<float>(pooled_w))));
  return g;
}

// sample coordinate of (bin p, sub-sample i): start + p*bin + (i + .5)*bin/grid, reference order
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int i, int grid) {
  const float a = __fadd_rn(start, __fmul_rn(static_cast<float>(p), bin));
  const float b = __fdiv_rn(__fmul_rn(static_cast<float>(i) + 0.5f, bin), static_cast<float>(grid));
  return __fadd_rn(a, b);
}

// 1-D half of bilinear_interpolate: returns false when the sample is outside [-1, size]
__device__ __forceinline__ bool axis_taps(float v, int size, int& low, int& high, float& wl, float& wh) {
  if (v < -1.0f || v > static_cast<float>(size)) return false;
  if (v <= 0.0f) v = 0.0f;
  low = static_cast<int>(v);
  if (low >= size - 1) {
    high = low = size - 1;
    v = static_cast<float>(low);
  } else {
    high = low + 1;
  }
  wh = v - static_cast<float>(low);  // weight of `high`
  wl = 1.0f - wh;                    // weight of `low`
  return true;
}

// Builds w[p][0..size) and the non-zero range [lo[p], hi[p]] for one axis; called by `pooled` threads.
__device__ __forceinline__ void build_axis_table(float* w, int* lo, int* hi, int p, int size, float start, float bin,
                                                 int grid) {
  float* row = w + p * size;
  for (int i = 0; i < size; ++i) row[i] = 0.0f;
  int mn = size, mx = -1;
  for (int i = 0; i < grid; ++i) {
    const float v = sample_coord(start, bin, p, i, grid);
    int l, h;
    float wl, wh;
    if (!axis_taps(v, size, l, h, wl, wh)) continue;
    row[l] += wl;
    row[h] += wh;
    mn = min(mn, l);
    mx = max(mx, h);
  }
  lo[p] = mn;
  hi[p] = mx;
}

constexpr int kRoiMaxPooled = 16;

// VEC channels per thread.  NCHW_OUT: output [R][C][ph][pw] staged through shared memory so the
// global write is one contiguous run; otherwise output [R][ph][pw][C] (fp32 and/or bf16 hi/lo).
// grid: (num_rois, channel_groups), block: cg_channels / VEC threads
template <int VEC, bool NCHW_OUT>
__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(const float* __restrict__ feat /*NHWC*/, const float* __restrict__ rois, int channels, int height,
                     int width, int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio,
                     float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ float s_dyn[];
  float* s_wy = s_dyn;                          // [pooled_h][height]
  float* s_wx = s_wy + pooled_h * height;       // [pooled_w][width]
  float* s_stage = s_wx + pooled_w * width;     // NCHW_OUT: [blockDim.x*VEC][pooled_h*pooled_w]
  __shared__ int s_ylo[kRoiMaxPooled], s_yhi[kRoiMaxPooled], s_xlo[kRoiMaxPooled], s_xhi[kRoiMaxPooled];
  __shared__ RoiGeom s_g;

  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) s_g = roi_geometry(rois + static_cast<long long>(r) * 5, spatial_scale, pooled_h, pooled_w, sampling_ratio);
  __syncthreads();
  const RoiGeom g = s_g;
  if (tid < pooled_h) build_axis_table(s_wy, s_ylo, s_yhi, tid, height, g.start_h, g.bin_h, g.grid_h);
  if (tid >= 32 && tid < 32 + pooled_w)
    build_axis_table(s_wx, s_xlo, s_xhi, tid - 32, width, g.start_w, g.bin_w, g.grid_w);
  __syncthreads();

  const int cg_channels = blockDim.x * VEC;
  const int c0 = blockIdx.y * cg_channels + tid * VEC;
  const bool c_ok = c0 < channels;
  const float inv_count_den = static_cast<float>(g.grid_h * g.grid_w);
  const float* fbase = feat + static_cast<long long>(g.batch_ind) * height * width * channels + c0;
  const int bins = pooled_h * pooled_w;

  for (int ph = 0; ph < pooled_h; ++ph) {
    float acc[kRoiMaxPooled > 8 ? 8 : kRoiMaxPooled][VEC];
#pragma unroll
    for (int pw = 0; pw < 8; ++pw)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[pw][v] = 0.0f;
    const int ylo = s_ylo[ph], yhi = s_yhi[ph];
    for (int y = ylo; y <= yhi; ++y) {
      const float wy = s_wy[ph * height + y];
      if (wy == 0.0f) continue;
      const float* frow = fbase + static_cast<long long>(y) * width * channels;
#pragma unroll
      for (int pw = 0; pw < 8; ++pw) {
        if (pw < pooled_w) {
          float rs[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) rs[v] = 0.0f;
          const int xlo = s_xlo[pw], xhi = s_xhi[pw];
          for (int x = xlo; x <= xhi; ++x) {
            const float wx = s_wx[pw * width + x];
            if (c_ok) {
              if (VEC == 4) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(frow + static_cast<long long>(x) * channels));
                rs[0] += wx * f.x;
                rs[1 % VEC] += wx * f.y;
                rs[2 % VEC] += wx * f.z;
                rs[3 % VEC] += wx * f.w;
              } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) rs[v] += wx * __ldg(frow + static_cast<long long>(x) * channels + v);
              }
            }
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[pw][v] += wy * rs[v];
        }
      }
    }
    // write this row of bins
#pragma unroll
    for (int pw = 0; pw < 8; ++pw) {
      if (pw < pooled_w) {
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) o[v] = acc[pw][v] / inv_count_den;
        if (NCHW_OUT) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) s_stage[(tid * VEC + v) * bins + ph * pooled_w + pw] = o[v];
        } else if (c_ok) {
          const long long off = (static_cast<long long>(r) * bins + ph * pooled_w + pw) * channels + c0;
          if (out != nullptr) {
            if (VEC == 4) {
              *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
            } else {
#pragma unroll
              for (int v = 0; v < VEC; ++v) out[off + v] = o[v];
            }
          }
          if (out_hi != nullptr) {
            __nv_bfloat16 h[VEC], l[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) split_bf16(o[v], h[v], l[v]);
            if (VEC == 4) {
              *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(pack_bf16x2(h[0], h[1 % VEC]), pack_bf16x2(h[2 % VEC], h[3 % VEC]));
              if (out_lo != nullptr)
                *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(pack_bf16x2(l[0], l[1 % VEC]), pack_bf16x2(l[2 % VEC], l[3 % VEC]));
            } else {
#pragma unroll
              for (int v = 0; v < VEC; ++v) {
                out_hi[off + v] = h[v];
                if (out_lo != nullptr) out_lo[off + v] = l[v];
              }
            }
          }
        }
      }
    }
  }
  if (NCHW_OUT) {
    __syncthreads();
    // [R][C][bins]: this CTA's channels are one contiguous run of cg_channels*bins floats
    const int cbase = blockIdx.y * cg_channels;
    const int nch = min(cg_channels, channels - cbase);
    float* dst = out + (static_cast<long long>(r) * channels + cbase) * bins;
    const int total = nch * bins;
    for (int i = tid; i < total; i += blockDim.x) dst[i] = s_stage[i];
  }
}

// ---------------------------------------------------------------------------------------------
// 7x7 RoIAlign, the kernel of the pipeline and of the reference-layout operator (NHWC fp32 map in).
//
// Same separable math as above; organised so that the kernel is bound by the output write, not by the
// gather.  The gather cost of an RoI is its taps, (sum_ph ny) x (sum_pw nx) 16-byte loads per 4 channels in
// the plain per-bin form -- large RoIs dominate, and at one thread-private dependent load per tap the old
// kernel was latency-bound (20 % of HBM peak).  Here:
//   * grid (RoI, 512-channel slab), 128 threads x 4 consecutive channels: a warp load is one 512-byte run;
//   * the x pass of a map row is fully unrolled over (7 bins x T taps), T = the RoI's widest bin rounded up
//     to a compiled size, zero weights in the padding taps (the tap window of a narrower bin is shifted left
//     so that every address stays inside the row): 7*T independent 16-byte loads in flight per thread;
//   * the y pass ROLLS over the window rows: a row that two adjacent bins share is read once and added to
//     both (two live accumulator sets) whenever no row is shared by three bins -- which holds for every RoI
//     taller than ~14 feature rows, i.e. the expensive ones; small RoIs take the per-bin order;
//   * compact per-bin weight tables (<= 16 taps) built once per CTA in the reference's expression order;
//   * fused outputs: fp32, bf16 hi/lo pair, pair of (value + positional encoding[bin]) (dana.py:259), or
//     (MODE 1) the reference's [R][C][7][7] layout staged through shared memory and written as one
//     contiguous 100 KB run per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kRoiTaps = 16;
constexpr int kRoi7Threads = 128;

__device__ __forceinline__ void build_axis_compact(float* w /*[kRoiTaps]*/, int* lo_out, int* n_out, int p, int size,
                                                   float start, float bin, int grid) {
#pragma unroll
  for (int i = 0; i < kRoiTaps; ++i) w[i] = 0.0f;
  int lo = -1, hi = -1;
  for (int i = 0; i < grid; ++i) {
    const float v = sample_coord(start, bin, p, i, grid);
    int l, h;
    float wl, wh;
    if (!axis_taps(v, size, l, h, wl, wh)) continue;
    if (lo < 0) lo = l;          // sample coordinates are non-decreasing in i: the first valid tap is the lowest
    if (l - lo < kRoiTaps) w[l - lo] += wl;
    if (h - lo < kRoiTaps) w[h - lo] += wh;
    hi = h;
  }
  *lo_out = lo < 0 ? 0 : lo;
  // a bin of more than kRoiTaps - 2 map pixels does not fit the table; the kernel sends such RoIs to the direct path
  // before the tables are built (bin > 14), so the marker kRoiTaps + 1 below is a guard that never fires
  *n_out = lo < 0 ? 0 : (hi - lo + 1 > kRoiTaps ? kRoiTaps + 1 : hi - lo + 1);
}

struct Roi7Tables {
  float wy[8][kRoiTaps];      // row 7 is never used with a non-zero count (keeps the look-ahead read in bounds)
  float wx[7][kRoiTaps];
  float2 wx2[7][kRoiTaps];    // (w, w): operand of the packed FMA
  int ylo[7], ny[7], xlo[7], nx[7];
  RoiGeom g;
};

__device__ __forceinline__ void store4_pair(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, const float (&f)[4]) {
  uint32_t ph[2], pl[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - __uint_as_float(ph[e] << 16),
                                                    f[2 * e + 1] - __uint_as_float(ph[e] & 0xFFFF0000u));
    pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(ph[0], ph[1]);
  if (lo != nullptr) *reinterpret_cast<uint2*>(lo + off) = make_uint2(pl[0], pl[1]);
}

struct Roi7Out {
  float* out;                 // MODE 0: [R][49][C] fp32 (optional); MODE 1: [R][C][49] fp32
  __nv_bfloat16 *hi, *lo;     // MODE 0: bf16 pair (optional)
  const float* pe;            // [49][C]
  __nv_bfloat16 *qhi, *qlo;   // MODE 0: pair of value + pe[bin] (optional)
  __half* h16;                // MODE 0: one fp16 plane (optional): the layer4 input of the mixed-precision mode
};

// Accumulators are float2 pairs and every multiply-add is the packed fma.rn.f32x2 (two fp32 FMAs per issue slot:
// the kernel is issue-bound once its loads overlap); the x weights are stored duplicated (w, w) for that purpose.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// x pass of one map row for this thread's 4 channels: rs[pw] = sum_k wx[pw][k] * F[row][xlo[pw] + k]
template <int T, int CH>
__device__ __forceinline__ void roi7_row_pass(const float* __restrict__ rowp, int channels, const Roi7Tables& s,
                                              float2 (&rs)[7][2]) {
  const int cs = CH ? CH : channels;
#pragma unroll
  for (int pw = 0; pw < 7; ++pw) {
    const float* px = rowp + static_cast<long long>(s.xlo[pw]) * cs;
    float2 a0 = f2(0.f, 0.f), a1 = f2(0.f, 0.f);
    if constexpr (T > 0) {
      float4 f[T];
#pragma unroll
      for (int k = 0; k < T; ++k) f[k] = __ldg(reinterpret_cast<const float4*>(px + static_cast<long long>(k) * cs));
#pragma unroll
      for (int k = 0; k < T; ++k) {
        const float2 w = s.wx2[pw][k];
        a0 = __ffma2_rn(w, f2(f[k].x, f[k].y), a0);
        a1 = __ffma2_rn(w, f2(f[k].z, f[k].w), a1);
      }
    } else {
      const int n = s.nx[pw];
      for (int k = 0; k < n; ++k, px += cs) {
        const float2 w = s.wx2[pw][k];
        const float4 f = __ldg(reinterpret_cast<const float4*>(px));
        a0 = __ffma2_rn(w, f2(f.x, f.y), a0);
        a1 = __ffma2_rn(w, f2(f.z, f.w), a1);
      }
    }
    rs[pw][0] = a0;
    rs[pw][1] = a1;
  }
}

// writes bin row `ph`: value = (w0 * a + w1 * b) / count   (general order: w0 = 1, w1 = 0)
template <int MODE>
__device__ __forceinline__ void roi7_emit(int r, int ph, int c0, int channels, float inv_count, const float2 (&a)[7][2],
                                          const float2 (&b)[7][2], float w0, float w1, const Roi7Out& o, float* s_stage,
                                          int tid) {
  const float2 k0 = f2(w0 * inv_count, w0 * inv_count), k1 = f2(w1 * inv_count, w1 * inv_count);
#pragma unroll
  for (int pw = 0; pw < 7; ++pw) {
    const float2 v0 = __ffma2_rn(b[pw][0], k1, __fmul2_rn(a[pw][0], k0));
    const float2 v1 = __ffma2_rn(b[pw][1], k1, __fmul2_rn(a[pw][1], k0));
    const float v[4] = {v0.x, v0.y, v1.x, v1.y};
    const int bin = ph * 7 + pw;
    if constexpr (MODE == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s_stage[(tid * 4 + e) * 49 + bin] = v[e];
    } else {
      const long long off = (static_cast<long long>(r) * 49 + bin) * channels + c0;
      if (o.out != nullptr) *reinterpret_cast<float4*>(o.out + off) = make_float4(v[0], v[1], v[2], v[3]);
      if (o.hi != nullptr) store4_pair(o.hi, o.lo, off, v);
      if (o.h16 != nullptr) {
        const __half2 h0 = __floats2half2_rn(fminf(fmaxf(v[0], -65504.f), 65504.f), fminf(fmaxf(v[1], -65504.f), 65504.f));
        const __half2 h1 = __floats2half2_rn(fminf(fmaxf(v[2], -65504.f), 65504.f), fminf(fmaxf(v[3], -65504.f), 65504.f));
        *reinterpret_cast<uint2*>(o.h16 + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0),
                                                            *reinterpret_cast<const uint32_t*>(&h1));
      }
      if (o.qhi != nullptr) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(o.pe + static_cast<long long>(bin) * channels + c0));
        const float q[4] = {v[0] + p4.x, v[1] + p4.y, v[2] + p4.z, v[3] + p4.w};
        store4_pair(o.qhi, o.qlo, off, q);
      }
    }
  }
}

// The y loop: bins in order, two register sets.  Two orders, chosen per RoI (uniform):
//   rolling   the sets accumulate the current and the next bin; a map row shared by both is read once and added to
//             both.  A bin starts at the first row not yet counted for it, or at the first row of the NEXT bin if
//             that lies lower (rows shared by three bins, RoIs 112..168 px tall, are read twice instead of three
//             times); window rows of an RoI taller than ~170 px are each read exactly once;
//   two-row   (every bin has <= 2 row taps, i.e. grid_h == 1, RoIs up to 112 px tall): the sets CACHE the x passes of
//             two map rows (tagged with their row index); a bin is w0 * row(l) + w1 * row(l + 1), and since l advances
//             by 0 or 1 from bin to bin every window row is passed once (3..8 passes instead of 14).
// ONE copy of the loop nest and ONE call site of the row pass, with the compiled tap count selected by a uniform
// switch: the per-CTA instruction footprint stays a few KB (an earlier version instantiated the whole nest per tap
// count and spent 40 % of its stall samples on instruction fetch, ncu `stalled_no_instructions`).
template <int CH, int MODE>
__device__ __forceinline__ void roi7_gather(const float* __restrict__ fbase, int channels, int width, const Roi7Tables& s,
                                            int taps, int order, int r, int c0, float inv_count, const Roi7Out& o,
                                            float* s_stage, int tid) {
  const bool tworow = order == 2;
  const long long row_pitch = static_cast<long long>(width) * (CH ? CH : channels);
  float2 ca[7][2], na[7][2];
#pragma unroll
  for (int pw = 0; pw < 7; ++pw) ca[pw][0] = ca[pw][1] = na[pw][0] = na[pw][1] = f2(0.f, 0.f);
  int y_cont = 0;             // rolling order: map rows below y_cont are already accumulated in `ca`
  int tag0 = -1, tag1 = -1;   // two-row order: map rows whose x pass sits in ca / na
#pragma unroll 1
  for (int cur = 0; cur < 7; ++cur) {
    const int ylo_c = s.ylo[cur];
    const int nyc = s.ny[cur];
    const int yend = ylo_c + nyc;
    int nlo = 0x3fffffff, nn = 0;
    if (!tworow && cur < 6 && s.ny[cur + 1] > 0) {
      nlo = s.ylo[cur + 1];
      nn = s.ny[cur + 1];
    }
    // rolling: continue below the last row already counted for this bin, but not past the first row of the next bin
    // (a row that the bin before last also used is read again, for the next bin only)
    int y = tworow ? ylo_c : max(ylo_c, min(y_cont, nlo));
    if (tworow && nyc > 0 && tag0 != ylo_c && tag1 == ylo_c) {   // the row cached in `na` becomes this bin's first row
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        ca[pw][0] = na[pw][0];
        ca[pw][1] = na[pw][1];
      }
      tag0 = ylo_c;
      tag1 = -1;
    }
#pragma unroll 1
    for (; y < yend; ++y) {
      float wa = 0.0f, wb = 0.0f;
      const bool first = (y == ylo_c);
      if (tworow) {
        if (first ? (tag0 == y) : (tag1 == y)) continue;      // already cached
      } else {
        wa = (y >= y_cont) ? s.wy[cur][y - ylo_c] : 0.0f;
        const int d = y - nlo;
        wb = (d >= 0 && d < nn) ? s.wy[cur + 1][d] : 0.0f;
        if (wa == 0.0f && wb == 0.0f) continue;
      }
      float2 rs[7][2];
      const float* rowp = fbase + static_cast<long long>(y) * row_pitch;
      switch (taps) {
        case 2: roi7_row_pass<2, CH>(rowp, channels, s, rs); break;
        case 3: roi7_row_pass<3, CH>(rowp, channels, s, rs); break;
        case 4: roi7_row_pass<4, CH>(rowp, channels, s, rs); break;
        case 6: roi7_row_pass<6, CH>(rowp, channels, s, rs); break;
        case 8: roi7_row_pass<8, CH>(rowp, channels, s, rs); break;
        case 12: roi7_row_pass<12, CH>(rowp, channels, s, rs); break;
        case 16: roi7_row_pass<16, CH>(rowp, channels, s, rs); break;
        default: roi7_row_pass<0, CH>(rowp, channels, s, rs); break;
      }
      if (tworow) {
        if (first) {
#pragma unroll
          for (int pw = 0; pw < 7; ++pw) {
            ca[pw][0] = rs[pw][0];
            ca[pw][1] = rs[pw][1];
          }
          tag0 = y;
        } else {
#pragma unroll
          for (int pw = 0; pw < 7; ++pw) {
            na[pw][0] = rs[pw][0];
            na[pw][1] = rs[pw][1];
          }
          tag1 = y;
        }
        continue;
      }
      const float2 wa2 = f2(wa, wa);
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        ca[pw][0] = __ffma2_rn(wa2, rs[pw][0], ca[pw][0]);
        ca[pw][1] = __ffma2_rn(wa2, rs[pw][1], ca[pw][1]);
      }
      if (wb != 0.0f) {
        const float2 wb2 = f2(wb, wb);
#pragma unroll
        for (int pw = 0; pw < 7; ++pw) {
          na[pw][0] = __ffma2_rn(wb2, rs[pw][0], na[pw][0]);
          na[pw][1] = __ffma2_rn(wb2, rs[pw][1], na[pw][1]);
        }
      }
    }
    if (tworow) {
      const float w0 = nyc > 0 ? s.wy[cur][0] : 0.0f;
      const float w1 = nyc > 1 ? s.wy[cur][1] : 0.0f;
      roi7_emit<MODE>(r, cur, c0, channels, inv_count, ca, na, w0, w1, o, s_stage, tid);
    } else {
      roi7_emit<MODE>(r, cur, c0, channels, inv_count, ca, na, 1.0f, 0.0f, o, s_stage, tid);
      y_cont = max(y_cont, yend);
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        ca[pw][0] = na[pw][0];
        ca[pw][1] = na[pw][1];
        na[pw][0] = na[pw][1] = f2(0.f, 0.f);
      }
    }
  }
}

// Direct form for RoIs whose bins do not fit the tap tables (only RoIs several times larger than the map, which the
// proposal layer never emits but the operator boundary accepts): every grid sample is evaluated as in
// ROIAlign_cpu.cpp:56-109, four taps each.  Correct for any geometry, not tuned.  The kernel branches to it right
// after the geometry, before any table state is live (branching later, next to the gather, cost the common path
// 30 % in register spills; a second launch for these RoIs cost 7 %).
template <int MODE>
__device__ __forceinline__ void roi7_direct(const float* __restrict__ fbase, int channels, int height, int width,
                                         const RoiGeom& g, int r, int c0, float inv_count, const Roi7Out& o,
                                         float* s_stage, int tid) {
  const long long cs = channels;
  for (int ph = 0; ph < 7; ++ph) {
    float2 acc[7][2];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) acc[pw][0] = acc[pw][1] = f2(0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      int yl, yh;
      float wyl, wyh;
      if (!axis_taps(y, height, yl, yh, wyl, wyh)) continue;
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const float x = sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
          int xl, xh;
          float wxl, wxh;
          if (!axis_taps(x, width, xl, xh, wxl, wxh)) continue;
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(fbase + (static_cast<long long>(yl) * width + xl) * cs));
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(fbase + (static_cast<long long>(yl) * width + xh) * cs));
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(fbase + (static_cast<long long>(yh) * width + xl) * cs));
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(fbase + (static_cast<long long>(yh) * width + xh) * cs));
          const float w1 = wyl * wxl, w2 = wyl * wxh, w3 = wyh * wxl, w4 = wyh * wxh;
          acc[pw][0].x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
          acc[pw][0].y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
          acc[pw][1].x += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
          acc[pw][1].y += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
        }
      }
    }
    roi7_emit<MODE>(r, ph, c0, channels, inv_count, acc, acc, 1.0f, 0.0f, o, s_stage, tid);
  }
}

// MODE 1: [R][C][49] -- this CTA's channels are one contiguous run of the output
__device__ __forceinline__ void roi7_copy_out(const Roi7Out& o, int r, int channels, const float* s_stage, int tid) {
  const int cbase = blockIdx.y * kRoi7Threads * 4;
  const int nch = min(kRoi7Threads * 4, channels - cbase);
  float* dst = o.out + (static_cast<long long>(r) * channels + cbase) * 49;
  const int total = nch * 49;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (int i = tid; i < total / 4; i += kRoi7Threads)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_stage)[i];
  } else {
    for (int i = tid; i < total; i += kRoi7Threads) dst[i] = s_stage[i];
  }
}

// grid (num_rois, ceil(C / 512)), block 128.  MODE 0: NHWC outputs; MODE 1: [R][C][49] via shared memory.
template <int MODE, int CH, int OCC = (MODE == 1 ? 2 : 4)>
__global__ void __launch_bounds__(kRoi7Threads, OCC)
roi_align7_kernel(const float* __restrict__ feat /*NHWC*/, const float* __restrict__ rois, int channels, int height,
                  int width, float spatial_scale, int sampling_ratio, Roi7Out o) {
  extern __shared__ float s_stage[];   // MODE 1: [512][49]
  __shared__ Roi7Tables s;
  __shared__ int s_t, s_rolling;
  const int r = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s.g = roi_geometry(rois + static_cast<long long>(r) * 5, spatial_scale, 7, 7, sampling_ratio);
  __syncthreads();
  const RoiGeom g = s.g;
  if (g.bin_w > 14.0f || g.bin_h > 14.0f) {
    // bins that may not fit the 16-tap tables (an RoI much larger than the map): direct sampling, decided before any
    // of the table machinery is live
    const int c0d = (blockIdx.y * kRoi7Threads + tid) * 4;
    const bool okd = c0d < channels;
    const float* fb = feat + static_cast<long long>(g.batch_ind) * height * width * channels + (okd ? c0d : 0);
    if (okd || MODE == 1)
      roi7_direct<MODE>(fb, channels, height, width, g, r, c0d, 1.0f / static_cast<float>(g.grid_h * g.grid_w), o, s_stage, tid);
    if constexpr (MODE == 1) {
      __syncthreads();
      roi7_copy_out(o, r, channels, s_stage, tid);
    }
    return;
  }
  if (tid < 7) build_axis_compact(s.wy[tid], &s.ylo[tid], &s.ny[tid], tid, height, g.start_h, g.bin_h, g.grid_h);
  if (tid >= 32 && tid < 39)
    build_axis_compact(s.wx[tid - 32], &s.xlo[tid - 32], &s.nx[tid - 32], tid - 32, width, g.start_w, g.bin_w, g.grid_w);
  __syncthreads();
  // compiled tap count T >= widest bin; 0 = dynamic loops (map narrower than the compiled size)
  int maxn = 0, maxny = 0;
#pragma unroll
  for (int pw = 0; pw < 7; ++pw) {
    maxn = max(maxn, s.nx[pw]);
    maxny = max(maxny, s.ny[pw]);
  }
  const bool overflow = maxn > kRoiTaps || maxny > kRoiTaps;   // uniform: read from shared memory by every thread
  int t = maxn <= 2 ? 2 : maxn <= 3 ? 3 : maxn <= 4 ? 4 : maxn <= 6 ? 6 : maxn <= 8 ? 8 : maxn <= 12 ? 12 : 16;
  if (t > width) t = 0;
  if (overflow) t = 0;
  if (t > 0 && tid >= 32 && tid < 39) {
    // shift this bin's tap window left so that xlo + t <= width: padding taps read valid pixels with zero weight
    const int pw = tid - 32;
    const int lo = s.xlo[pw];
    const int base = min(lo, width - t);
    const int sh = lo - base;
    if (sh > 0) {
      for (int k = kRoiTaps - 1; k >= 0; --k) s.wx[pw][k] = (k >= sh) ? s.wx[pw][k - sh] : 0.0f;
      s.xlo[pw] = base;
    }
  }
  if (tid >= 32 && tid < 39) {
#pragma unroll
    for (int k = 0; k < kRoiTaps; ++k) s.wx2[tid - 32][k] = make_float2(s.wx[tid - 32][k], s.wx[tid - 32][k]);
  }
  if (tid == 0) {
    bool small = true;          // two-row order: every bin has at most two row taps
    for (int p = 0; p < 7; ++p)
      if (s.ny[p] > 2) small = false;
    s_rolling = small ? 2 : 1;
    s_t = t;
  }
  __syncthreads();
  const int c0 = (blockIdx.y * kRoi7Threads + tid) * 4;
  const bool c_ok = c0 < channels;
  const float inv_count = 1.0f / static_cast<float>(g.grid_h * g.grid_w);
  const float* fbase = feat + static_cast<long long>(g.batch_ind) * height * width * channels + (c_ok ? c0 : 0);
  if ((c_ok || MODE == 1) && !overflow) {   // (overflow cannot occur past the early branch: bins <= 14 px -> <= 16 taps)
    roi7_gather<CH, MODE>(fbase, channels, width, s, s_t, s_rolling, r, c0, inv_count, o, s_stage, tid);
  }
  if constexpr (MODE == 1) {
    __syncthreads();
    roi7_copy_out(o, r, channels, s_stage, tid);
  }
}

template <int MODE>
inline int roi_align7_launch(const float* feat_nhwc, const float* rois, int num_rois, int channels, int height, int width,
                             float spatial_scale, int sampling_ratio, const Roi7Out& o, cudaStream_t stream) {
  const int groups = (channels + kRoi7Threads * 4 - 1) / (kRoi7Threads * 4);
  const size_t smem = MODE == 1 ? sizeof(float) * kRoi7Threads * 4 * 49 : 0;
  if (MODE == 1) {
    static bool configured_dev[kMaxDevices] = {};
    bool& configured = configured_dev[current_device()];
    if (!configured) {
      DANA_CUDA_CHECK(cudaFuncSetAttribute(roi_align7_kernel<MODE, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
      DANA_CUDA_CHECK(cudaFuncSetAttribute(roi_align7_kernel<MODE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
      configured = true;
    }
  }
  if (channels == 1024) {
    // (a 5-CTA/SM register allocation -- 96 registers, ~150 bytes of spills -- was measured slower: 0.145 vs 0.117 ms)
    roi_align7_kernel<MODE, 1024><<<dim3(num_rois, groups), kRoi7Threads, smem, stream>>>(
        feat_nhwc, rois, channels, height, width, spatial_scale, sampling_ratio, o);
  } else {
    roi_align7_kernel<MODE, 0><<<dim3(num_rois, groups), kRoi7Threads, smem, stream>>>(
        feat_nhwc, rois, channels, height, width, spatial_scale, sampling_ratio, o);
  }
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

// (bins that do not fit the 16-tap tables -- maps larger than 98 pixels, RoIs larger than the map -- are handled
// per RoI by the direct path)
inline bool roi_align7_supported(int channels, int /*height*/, int /*width*/, int /*sampling_ratio*/) {
  return channels % 4 == 0;
}

inline int roi_align_head_run(const float* feat_nhwc, const float* rois, int num_rois, int batch, int channels,
                              int height, int width, float spatial_scale, int sampling_ratio, float* out, void* out_hi,
                              void* out_lo, const float* pe, void* qpe_hi, void* qpe_lo, void* out_f16,
                              cudaStream_t stream) {
  if (num_rois == 0) return DANA_OK;
  if (!feat_nhwc || !rois || num_rois < 0 || batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return DANA_EINVAL;
  if (!out && !out_hi && !qpe_hi && !out_f16) return DANA_EINVAL;
  if (qpe_hi && !pe) return DANA_EINVAL;
  if (!roi_align7_supported(channels, height, width, sampling_ratio)) return DANA_ENOTSUP;
  Roi7Out o{out, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo), pe,
            static_cast<__nv_bfloat16*>(qpe_hi), static_cast<__nv_bfloat16*>(qpe_lo), static_cast<__half*>(out_f16)};
  return roi_align7_launch<0>(feat_nhwc, rois, num_rois, channels, height, width, spatial_scale, sampling_ratio, o, stream);
}

// NCHW -> NHWC transpose of one feature map batch: [B][C][HW] -> [B][HW][C], 32x32 tiles
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int channels, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* src = in + static_cast<long long>(b) * channels * hw;
  float* dst = out + static_cast<long long>(b) * channels * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < channels && p < hw) ? src[static_cast<long long>(c) * hw + p] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < hw && c < channels) dst[static_cast<long long>(p) * channels + c] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the 7x7 RoIAlign (semantics of lib/model/csrc/cuda/ROIAlign_cuda.cu:178-254: every sample scatters its
// bin gradient to its four taps), as the TRANSPOSE of the forward's separable form:
//   dF[y][x] += (1/count) * sum_ph wy[ph][y] * sum_pw wx[pw][x] * dOut[ph][pw]
// NHWC on both sides.  Grid (RoI, 512-channel slab), 128 threads x 4 consecutive channels, the same compact per-bin tap
// tables as the forward.  A map pixel of the RoI window receives ONE 16-byte vector reduction (red.global.add.v4.f32)
// per bin row / bin column pair that touches it -- (sum ny) x (sum nx) per thread, coalesced 512-byte runs per warp --
// instead of grid_h * grid_w * 4 scalar atomics per bin and channel on an NCHW map (a 600-px RoI: ~2 k vector
// reductions against ~28 k scalar ones per 4 channels).  The x contraction of a bin row is done once per (ph, x) in
// registers and reused by all its map rows.  RoIs whose bins exceed the tables take the direct per-sample path.
// (Float atomics: the summation order, hence the last bits, vary run to run -- as in the reference.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}

__global__ void __launch_bounds__(kRoi7Threads, 4)
roi_align7_bwd_kernel(const float* __restrict__ grad_out /*[R][49][C]*/, const float* __restrict__ rois, int channels,
                      int height, int width, float spatial_scale, int sampling_ratio, float* __restrict__ grad_in /*NHWC*/) {
  __shared__ Roi7Tables s;
  const int r = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s.g = roi_geometry(rois + static_cast<long long>(r) * 5, spatial_scale, 7, 7, sampling_ratio);
  __syncthreads();
  const RoiGeom g = s.g;
  const int c0 = (blockIdx.y * kRoi7Threads + tid) * 4;
  const bool c_ok = c0 < channels;
  const float inv_count = 1.0f / static_cast<float>(g.grid_h * g.grid_w);
  float* gbase = grad_in + static_cast<long long>(g.batch_ind) * height * width * channels + (c_ok ? c0 : 0);
  const float* go = grad_out + static_cast<long long>(r) * 49 * channels + (c_ok ? c0 : 0);
  if (g.bin_w > 14.0f || g.bin_h > 14.0f) {
    // bins wider than the tap tables (an RoI much larger than the map): per-sample scatter, still NHWC / vectorised
    if (!c_ok) return;
    for (int ph = 0; ph < 7; ++ph) {
      for (int pw = 0; pw < 7; ++pw) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(go + static_cast<long long>(ph * 7 + pw) * channels));
        for (int iy = 0; iy < g.grid_h; ++iy) {
          const float y = sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
          int yl, yh;
          float wyl, wyh;
          if (!axis_taps(y, height, yl, yh, wyl, wyh)) continue;
          for (int ix = 0; ix < g.grid_w; ++ix) {
            const float x = sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
            int xl, xh;
            float wxl, wxh;
            if (!axis_taps(x, width, xl, xh, wxl, wxh)) continue;
            const float w00 = wyl * wxl * inv_count, w01 = wyl * wxh * inv_count, w10 = wyh * wxl * inv_count,
                        w11 = wyh * wxh * inv_count;
            red_add4(gbase + (static_cast<long long>(yl) * width + xl) * channels, d.x * w00, d.y * w00, d.z * w00, d.w * w00);
            red_add4(gbase + (static_cast<long long>(yl) * width + xh) * channels, d.x * w01, d.y * w01, d.z * w01, d.w * w01);
            red_add4(gbase + (static_cast<long long>(yh) * width + xl) * channels, d.x * w10, d.y * w10, d.z * w10, d.w * w10);
            red_add4(gbase + (static_cast<long long>(yh) * width + xh) * channels, d.x * w11, d.y * w11, d.z * w11, d.w * w11);
          }
        }
      }
    }
    return;
  }
  if (tid < 7) build_axis_compact(s.wy[tid], &s.ylo[tid], &s.ny[tid], tid, height, g.start_h, g.bin_h, g.grid_h);
  if (tid >= 32 && tid < 39)
    build_axis_compact(s.wx[tid - 32], &s.xlo[tid - 32], &s.nx[tid - 32], tid - 32, width, g.start_w, g.bin_w, g.grid_w);
  __syncthreads();
  if (!c_ok) return;
  for (int ph = 0; ph < 7; ++ph) {
    const int ny = s.ny[ph];
    if (ny == 0) continue;
    float4 d[7];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) {
      d[pw] = __ldg(reinterpret_cast<const float4*>(go + static_cast<long long>(ph * 7 + pw) * channels));
      d[pw].x *= inv_count; d[pw].y *= inv_count; d[pw].z *= inv_count; d[pw].w *= inv_count;
    }
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) {
      const int nx = s.nx[pw], x0 = s.xlo[pw];
      for (int j = 0; j < nx; ++j) {
        const float wx = s.wx[pw][j];
        if (wx == 0.0f) continue;
        const float4 e = make_float4(d[pw].x * wx, d[pw].y * wx, d[pw].z * wx, d[pw].w * wx);   // x-contracted, reused over y
        float* col = gbase + static_cast<long long>(x0 + j) * channels;
        for (int k = 0; k < ny; ++k) {
          const float wy = s.wy[ph][k];
          if (wy == 0.0f) continue;
          red_add4(col + static_cast<long long>(s.ylo[ph] + k) * width * channels, e.x * wy, e.y * wy, e.z * wy, e.w * wy);
        }
      }
    }
  }
}

// [R][C][49] -> [R][49][C] (and back for the map: see nhwc_to_nchw_kernel) -- the reference-layout boundary of the backward
__global__ void rc49_to_r49c_kernel(const float* __restrict__ in, float* __restrict__ out, int channels) {
  __shared__ float tile[32][50];
  const int r = blockIdx.x, c0 = blockIdx.y * 32;
  const float* src = in + (static_cast<long long>(r) * channels + c0) * 49;
  const int nch = min(32, channels - c0);
  for (int i = threadIdx.x; i < nch * 49; i += blockDim.x) tile[i / 49][i % 49] = src[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * 32; i += blockDim.x) {
    const int bin = i >> 5, c = i & 31;
    if (c < nch) out[(static_cast<long long>(r) * 49 + bin) * channels + c0 + c] = tile[c][bin];
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int channels, int hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* src = in + static_cast<long long>(b) * channels * hw;
  float* dst = out + static_cast<long long>(b) * channels * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && c < channels) ? src[static_cast<long long>(p) * channels + c] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < channels && p < hw) dst[static_cast<long long>(c) * hw + p] = tile[threadIdx.x][i];
  }
}

// Generic pooled sizes (not used by any shipped configuration): per-element scatter on the reference's NCHW layout.
__global__ void roi_align_bwd_generic_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois,
                                             long long nthreads, int channels, int height, int width, int pooled_h,
                                             int pooled_w, float spatial_scale, int sampling_ratio,
                                             float* __restrict__ grad_in) {
  for (long long index = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; index < nthreads;
       index += static_cast<long long>(blockDim.x) * gridDim.x) {
    const int pw = static_cast<int>(index % pooled_w);
    const int ph = static_cast<int>((index / pooled_w) % pooled_h);
    const int c = static_cast<int>((index / pooled_w / pooled_h) % channels);
    const int n = static_cast<int>(index / pooled_w / pooled_h / channels);
    const RoiGeom g = roi_geometry(rois + static_cast<long long>(n) * 5, spatial_scale, pooled_h, pooled_w, sampling_ratio);
    float* gin = grad_in + (static_cast<long long>(g.batch_ind) * channels + c) * height * width;
    const float go = grad_out[index] / static_cast<float>(g.grid_h * g.grid_w);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h);
      int yl, yh;
      float wyl, wyh;
      if (!axis_taps(y, height, yl, yh, wyl, wyh)) continue;
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w);
        int xl, xh;
        float wxl, wxh;
        if (!axis_taps(x, width, xl, xh, wxl, wxh)) continue;
        atomicAdd(gin + yl * width + xl, go * (wyl * wxl));
        atomicAdd(gin + yl * width + xh, go * (wyl * wxh));
        atomicAdd(gin + yh * width + xl, go * (wyh * wxl));
        atomicAdd(gin + yh * width + xh, go * (wyh * wxh));
      }
    }
  }
}

inline int roi_align_forward_run(const float* input, const float* rois, int num_rois, int batch, int channels,
                                 int height, int width, int pooled_h, int pooled_w, float spatial_scale,
                                 int sampling_ratio, int layout, float* out, void* out_hi, void* out_lo,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (num_rois == 0) return DANA_OK;
  if (!input || !rois || num_rois < 0 || batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return DANA_EINVAL;
  if (pooled_h <= 0 || pooled_w <= 0 || pooled_h > 8 || pooled_w > 8) return DANA_ENOTSUP;
  const size_t table_bytes = sizeof(float) * (static_cast<size_t>(pooled_h) * height + static_cast<size_t>(pooled_w) * width);
  if (layout == 0) {
    if (!out || !workspace) return DANA_EINVAL;
    const int64_t need = 4LL * batch * channels * height * width;
    if (workspace_bytes < need) return DANA_EINVAL;
    float* nhwc = static_cast<float*>(workspace);
    const int hw = height * width;
    nchw_to_nhwc_kernel<<<dim3((hw + 31) / 32, (channels + 31) / 32, batch), dim3(32, 8), 0, stream>>>(input, nhwc,
                                                                                                      channels, hw);
    if (pooled_h == 7 && pooled_w == 7 && roi_align7_supported(channels, height, width, sampling_ratio)) {
      Roi7Out o{out, nullptr, nullptr, nullptr, nullptr, nullptr};
      return roi_align7_launch<1>(nhwc, rois, num_rois, channels, height, width, spatial_scale, sampling_ratio, o, stream);
    }
    const int threads = 128;
    const size_t smem = table_bytes + sizeof(float) * threads * pooled_h * pooled_w;
    static size_t configured_dev[kMaxDevices] = {};
    size_t& configured = configured_dev[current_device()];
    if (smem > 48 * 1024 && smem > configured) {
      DANA_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
      configured = smem;
    }
    roi_align_fwd_kernel<1, true><<<dim3(num_rois, (channels + threads - 1) / threads), threads, smem, stream>>>(
        nhwc, rois, channels, height, width, pooled_h, pooled_w, spatial_scale, sampling_ratio, out, nullptr, nullptr);
  } else if (layout == 1) {
    if (!out && !out_hi) return DANA_EINVAL;
    if (pooled_h == 7 && pooled_w == 7 && roi_align7_supported(channels, height, width, sampling_ratio)) {
      Roi7Out o{out, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo), nullptr, nullptr, nullptr, nullptr};
      return roi_align7_launch<0>(input, rois, num_rois, channels, height, width, spatial_scale, sampling_ratio, o, stream);
    }
    if (channels % 4 != 0) return DANA_ENOTSUP;
    // >= 64 threads: warp 0 builds the y table, warp 1 the x table
    const int threads = (channels / 4 >= 256) ? 256 : (((channels / 4 + 31) / 32) * 32 < 64 ? 64 : ((channels / 4 + 31) / 32) * 32);
    const int cg = threads * 4;
    static size_t configured_dev[kMaxDevices] = {};
    size_t& configured = configured_dev[current_device()];
    if (table_bytes > 48 * 1024 && table_bytes > configured) {
      DANA_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(table_bytes)));
      configured = table_bytes;
    }
    roi_align_fwd_kernel<4, false><<<dim3(num_rois, (channels + cg - 1) / cg), threads, table_bytes, stream>>>(
        input, rois, channels, height, width, pooled_h, pooled_w, spatial_scale, sampling_ratio, out,
        static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo));
  } else {
    return DANA_EINVAL;
  }
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

inline int64_t roi_align_backward_workspace(int num_rois, int batch, int channels, int height, int width, int pooled_h,
                                            int pooled_w, int layout) {
  if (layout != 0 || pooled_h != 7 || pooled_w != 7 || channels % 4 != 0) return 256;
  return 4LL * num_rois * channels * 49 + 256 + 4LL * batch * channels * height * width + 256;
}

// layout 0: grad_out [R][C][ph][pw], grad_input [B][C][H][W] (the reference's operator); 1: [R][ph*pw][C] / [B][H][W][C].
inline int roi_align_backward_run(const float* grad_out, const float* rois, int num_rois, int batch, int channels,
                                  int height, int width, int pooled_h, int pooled_w, float spatial_scale,
                                  int sampling_ratio, int layout, float* grad_input, void* workspace,
                                  int64_t workspace_bytes, cudaStream_t stream) {
  if (!grad_input || batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || (layout != 0 && layout != 1)) return DANA_EINVAL;
  const int64_t map_bytes = 4LL * batch * channels * height * width;
  const bool fast = pooled_h == 7 && pooled_w == 7 && channels % 4 == 0;
  if (layout == 1 && !fast) return DANA_ENOTSUP;
  if (num_rois == 0) {
    DANA_CUDA_CHECK(cudaMemsetAsync(grad_input, 0, map_bytes, stream));
    return DANA_OK;
  }
  if (!grad_out || !rois || num_rois < 0) return DANA_EINVAL;
  if (!fast) {
    DANA_CUDA_CHECK(cudaMemsetAsync(grad_input, 0, map_bytes, stream));
    const long long nthreads = static_cast<long long>(num_rois) * channels * pooled_h * pooled_w;
    const long long blocks = (nthreads + 255) / 256;
    roi_align_bwd_generic_kernel<<<static_cast<int>(blocks > 148 * 32 ? 148 * 32 : blocks), 256, 0, stream>>>(
        grad_out, rois, nthreads, channels, height, width, pooled_h, pooled_w, spatial_scale, sampling_ratio, grad_input);
    DANA_LAUNCH_CHECK();
    return DANA_OK;
  }
  const float* go = grad_out;
  float* gi = grad_input;
  if (layout == 0) {
    if (!workspace || workspace_bytes < roi_align_backward_workspace(num_rois, batch, channels, height, width, 7, 7, 0))
      return DANA_EINVAL;
    float* go_t = static_cast<float*>(workspace);
    const int64_t off = (4LL * num_rois * channels * 49 + 255) / 256 * 256;
    gi = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + off);
    rc49_to_r49c_kernel<<<dim3(num_rois, (channels + 31) / 32), 256, 0, stream>>>(grad_out, go_t, channels);
    go = go_t;
  }
  DANA_CUDA_CHECK(cudaMemsetAsync(gi, 0, map_bytes, stream));
  const int groups = (channels + kRoi7Threads * 4 - 1) / (kRoi7Threads * 4);
  roi_align7_bwd_kernel<<<dim3(num_rois, groups), kRoi7Threads, 0, stream>>>(go, rois, channels, height, width,
                                                                            spatial_scale, sampling_ratio, gi);
  if (layout == 0) {
    const int hw = height * width;
    nhwc_to_nchw_kernel<<<dim3((hw + 31) / 32, (channels + 31) / 32, batch), dim3(32, 8), 0, stream>>>(gi, grad_input,
                                                                                                      channels, hw);
  }
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

}  // namespace dana
