// Host launcher of the tcgen05 implicit-GEMM kernel: validates the request, picks the
// output tile, encodes the TMA tensor maps and launches one persistent CTA per SM.
#pragma once
#include <stdlib.h>

#include "api_common.cuh"
#include "conv_gemm.cuh"

namespace dana {

struct TileChoice {
  int bw, bh, bn;
};

// Pick (bw, bh, bn) with bw*bh*bn == 128 (powers of two) minimising padded work.
inline TileChoice choose_tile(int w, int h, int n, bool batched_b) {
  TileChoice best{128, 1, 1};
  long long best_cost = -1;
  for (int bw = 1; bw <= 128; bw <<= 1) {
    for (int bh = 1; bw * bh <= 128; bh <<= 1) {
      const int bn = 128 / (bw * bh);
      if (batched_b && bn != 1) continue;
      const long long tx = (w + bw - 1) / bw, ty = (h + bh - 1) / bh, tn = (n + bn - 1) / bn;
      const long long cost = tx * ty * tn;
      // prefer wider rows on ties (longer contiguous TMA runs)
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && bw > best.bw)) {
        best_cost = cost;
        best = TileChoice{bw, bh, bn};
      }
    }
  }
  return best;
}

template <int BLOCK_N, int NSPLIT, int EPI, int CM, int KT = 64, bool DX3 = false>
inline int launch_conv_gemm_v(const ConvGemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<BLOCK_N, NSPLIT, KT, DX3>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    DANA_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, NSPLIT, EPI, CM, KT, DX3>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  // Programmatic dependent launch (opt-in, DANA_PDL=1): consecutive GEMM launches let the next grid be scheduled while
  // this one drains; the kernel runs its prologue (barrier init, TMEM allocation, descriptor prefetch) and then waits
  // on `griddepcontrol.wait` before touching global memory, so stream order is preserved.  Measured on the benchmark
  // step (CUDA-graph replay): 7.82 ms with it, 7.66 ms without -- persistent CTAs fill every SM's shared memory, so
  // the dependent grid cannot start its prologue early and only the bookkeeping is paid.  Off by default.
  static int pdl = -1;
  if (pdl < 0) {
    const char* env = getenv("DANA_PDL");
    pdl = (env != nullptr && atoi(env) == 1) ? 1 : 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CM > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CM;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DANA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, NSPLIT, EPI, CM, KT, DX3>, p));
  DANA_LAUNCH_CHECK();
  return DANA_OK;
}

// The fast epilogue needs: bf16 pair output only, full 32-column chunks, 16-byte aligned pointers and
// strides that are multiples of 8 elements (so every 8-channel group is one aligned 16-byte access).
inline bool fast_epilogue_ok(const ConvGemmParams& p, int nsplit, bool io_f16 = false) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (p.out_f32 != nullptr || p.res_f32 != nullptr || p.out_hi == nullptr) return false;
  if ((p.n_out % 32) != 0) return false;
  if (!al16(p.out_hi) || (p.so_x % 8) || (p.so_y % 8) || (p.so_n % 8)) return false;
  if (io_f16) {   // single fp16 plane out (and residual), whatever the operand planes are
    if (p.out_lo != nullptr || p.res_lo != nullptr) return false;
    if (p.res_hi != nullptr && (!al16(p.res_hi) || (p.sr_x % 8) || (p.sr_y % 8) || (p.sr_n % 8))) return false;
    return true;
  }
  if (nsplit == 2 && (p.out_lo == nullptr || !al16(p.out_lo))) return false;
  if (nsplit == 1 && p.out_lo != nullptr) return false;
  if (p.res_hi != nullptr) {
    if (!al16(p.res_hi) || (p.sr_x % 8) || (p.sr_y % 8) || (p.sr_n % 8)) return false;
    if (nsplit == 2 && (p.res_lo == nullptr || !al16(p.res_lo))) return false;
    if (nsplit == 1 && p.res_lo != nullptr) return false;
  }
  return true;
}

template <int BLOCK_N, int NSPLIT, int CM, int KT = 64>
inline int launch_conv_gemm(const ConvGemmParams& p, int grid, cudaStream_t stream, bool io_f16 = false) {
  if (io_f16) {
    // fp16 single-plane output: the row-contiguous epilogue only (every use of the path has n_out % 32 == 0 and
    // aligned planes); single-plane operands at every tile width, split operands (the P.V contraction feeding the
    // fp16 RPN input) at 64 / 128 columns -- the launcher never picks 256-wide split tiles for an fp16 output
    if constexpr (CM == 1 && KT == 64 && (NSPLIT == 1 || BLOCK_N <= 128)) {
      if (p.sm_ns == 0 && fast_epilogue_ok(p, NSPLIT, true)) return launch_conv_gemm_v<BLOCK_N, NSPLIT, 3, 1>(p, grid, stream);
    }
    if constexpr (CM == 2 && KT == 64 && NSPLIT == 1 && BLOCK_N == 256) {   // experiment: DANA_CLUSTER=4 (fp16 planes)
      if (p.sm_ns == 0 && fast_epilogue_ok(p, NSPLIT, true)) return launch_conv_gemm_v<256, 1, 3, 2>(p, grid, stream);
    }
    return DANA_ENOTSUP;
  }
  if constexpr (KT == 32) {   // only the plain-epilogue, non-cluster variants are instantiated at KT = 32
    if (fast_epilogue_ok(p, NSPLIT)) return launch_conv_gemm_v<BLOCK_N, NSPLIT, 1, 1, 32>(p, grid, stream);
    return launch_conv_gemm_v<BLOCK_N, NSPLIT, 0, 1, 32>(p, grid, stream);
  }
  if (p.sm_ns > 0) {
    if constexpr (CM == 1 && BLOCK_N != 128) return launch_conv_gemm_v<BLOCK_N, NSPLIT, 2, 1>(p, grid, stream);
    return DANA_ENOTSUP;
  }
  if (fast_epilogue_ok(p, NSPLIT)) return launch_conv_gemm_v<BLOCK_N, NSPLIT, 1, CM>(p, grid, stream);
  return launch_conv_gemm_v<BLOCK_N, NSPLIT, 0, CM>(p, grid, stream);
}

inline int conv_gemm_dispatch(const dana_conv_gemm_args* a, cudaStream_t stream) {
  if (a == nullptr || a->a_hi == nullptr || a->b_hi == nullptr) return DANA_EINVAL;
  if ((a->a_lo == nullptr) != (a->b_lo == nullptr)) return DANA_EINVAL;
  if (a->out_hi == nullptr && a->out_f32 == nullptr) return DANA_EINVAL;
  if (a->n_out <= 0 || a->a_c <= 0 || a->out_w <= 0 || a->out_h <= 0 || a->out_n <= 0) return DANA_EINVAL;
  const int taps = a->taps_r * a->taps_s;
  if (taps < 1 || taps > 64) return DANA_EINVAL;
  if (taps > 1 && (a->a_c % 64) != 0) return DANA_EINVAL;
  // TMA alignment rules: 16-byte base and strides
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->a_hi) || !al16(a->b_hi) || (a->a_lo && !al16(a->a_lo)) || (a->b_lo && !al16(a->b_lo)))
    return DANA_EINVAL;
  if ((a->a_sx % 8) || (a->a_sy % 8) || (a->a_sn % 8) || (a->b_pitch % 8) || (a->b_batch_stride % 8))
    return DANA_EINVAL;
  const bool batched = a->b_batch_stride != 0;
  const bool softmax = a->softmax_ns > 0;
  const bool io_f16 = a->io_f16 != 0;
  if (io_f16 && (softmax || a->out_lo != nullptr || a->res_lo != nullptr || a->out_f32 != nullptr ||
                 a->res_f32 != nullptr || a->out_hi == nullptr))
    return DANA_EINVAL;
  if (a->ab_f16 && a->a_lo != nullptr) return DANA_EINVAL;   // fp16 operands are single planes
  if (softmax) {
    auto al16b = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (a->softmax_ns > 512 || a->softmax_pitch < a->softmax_ns || (a->softmax_pitch % 8) != 0) return DANA_EINVAL;
    if (a->softmax_ns > 256 && ((a->a_c % 32) != 0 || taps != 1)) return DANA_ENOTSUP;   // wide tile: 32-wide K stages
    if (a->out_hi == nullptr || a->out_f32 != nullptr || a->scale || a->bias || a->res_hi || a->res_f32) return DANA_EINVAL;
    if (!al16b(a->out_hi) || (a->out_lo && !al16b(a->out_lo)) || (a->o_sx % 8) || (a->o_sy % 8) || (a->o_sn % 8))
      return DANA_EINVAL;
    if ((a->a_lo != nullptr) != (a->out_lo != nullptr)) return DANA_EINVAL;
  }

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  TileChoice tc{a->tile_w, a->tile_h, a->tile_n};
  // 3x3 stride-1 pad-1 convolutions with split operands can take the window-per-column-shift schedule (ConvGemmCfg:
  // DX3; output tiles of 8 x 16 or 16 x 8 pixels, whichever pads the map less).  Measured on B200 (round 2,
  // tools/gemm_bench.py / bench.py): 64 channels (layer1) 0.0605 -> 0.0537 ms and +0.7 % on the step; 128 / 256 channels
  // LOSE (0.0455 -> 0.0477, 0.0395 -> 0.0433 ms: only two 68 KB stages fit, and these layers were never bound by the
  // L2 -> SM bytes the schedule saves -- ncu shows the tensor pipe, at 40 % efficiency for 64-column MMAs, as the
  // busiest unit).  Default: 64-channel layers only; DANA_DX3=<max channels> overrides (0 = off).
  bool dx3 = false;
  {
    static int dx3_env = -1;
    if (dx3_env < 0) {
      const char* env = getenv("DANA_DX3");   // largest input-channel count that takes the schedule (0 = off)
      dx3_env = (env != nullptr) ? atoi(env) : 64;
    }
    dx3 = a->a_c <= dx3_env && a->taps_r == 3 && a->taps_s == 3 && a->pad_y == 1 && a->pad_x == 1 && a->a_lo != nullptr && !batched &&
          !softmax && !io_f16 && a->bias_sn == 0 && (a->a_c == 64 || a->a_c == 128 || a->a_c == 256) &&
          (a->n_out == 64 || a->n_out % 128 == 0) && a->tile_w <= 0 && a->out_w == a->a_w && a->out_h == a->a_h &&
          a->out_f32 == nullptr && a->res_f32 == nullptr && a->out_hi != nullptr;   // the schedule only has the fast epilogue
    if (dx3) {
      auto cost = [&](int bw) {
        const int bh = 128 / bw;
        return static_cast<long long>((a->out_w + bw - 1) / bw) * ((a->out_h + bh - 1) / bh);
      };
      tc = (cost(16) < cost(8)) ? TileChoice{16, 8, 1} : TileChoice{8, 16, 1};
    }
  }
  if (tc.bw <= 0 || tc.bh <= 0 || tc.bn <= 0)
    tc = choose_tile(a->out_w, a->out_h, a->out_n, batched || a->bias_sn != 0);
  if (tc.bw * tc.bh * tc.bn != 128 || (batched && tc.bn != 1)) return DANA_EINVAL;
  p.bw = tc.bw;
  p.bh = tc.bh;
  p.bn = tc.bn;
  auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  if ((tc.bw & (tc.bw - 1)) || (tc.bh & (tc.bh - 1))) return DANA_EINVAL;   // power-of-two tile extents
  p.lbw = ilog2(tc.bw);
  p.lbh = ilog2(tc.bh);
  p.tiles_x = (a->out_w + tc.bw - 1) / tc.bw;
  p.tiles_y = (a->out_h + tc.bh - 1) / tc.bh;
  p.tiles_n = (a->out_n + tc.bn - 1) / tc.bn;
  p.taps_r = a->taps_r;
  p.taps_s = a->taps_s;
  p.pad_y = a->pad_y;
  p.pad_x = a->pad_x;
  p.c_in = static_cast<int>(a->a_c);
  p.n_out = a->n_out;
  p.out_w = a->out_w;
  p.out_h = a->out_h;
  p.out_n = a->out_n;
  p.so_x = a->o_sx;
  p.so_y = a->o_sy;
  p.so_n = a->o_sn;
  p.sr_x = a->r_sx;
  p.sr_y = a->r_sy;
  p.sr_n = a->r_sn;
  p.b_batched = batched ? 1 : 0;
  p.bias_sn = a->bias_sn;
  if (a->bias_sn != 0 && tc.bn != 1) return DANA_EINVAL;
  p.relu = a->relu;
  p.alpha = a->alpha;
  p.scale = a->scale;
  p.bias = a->bias;
  p.res_hi = static_cast<const __nv_bfloat16*>(a->res_hi);
  p.res_lo = static_cast<const __nv_bfloat16*>(a->res_lo);
  p.res_f32 = a->res_f32;
  p.out_hi = static_cast<__nv_bfloat16*>(a->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);
  p.out_f32 = a->out_f32;
  p.ab_f16 = a->ab_f16 ? 1 : 0;
  {
    // cross-tile residual look-ahead (conv_gemm.cuh, epilogue).  Measured (round 2, tools/gemm_bench.py): no gain --
    // layer1 conv3+res 0.0946 ms with it, 0.0920 without; the step 701.6 vs 706.1 images/s -- so it is an opt-in
    // (DANA_RES_CROSS=1): the residual stall ncu shows is not a matter of issue distance within one warp
    static int rc = -1;
    if (rc < 0) {
      const char* env = getenv("DANA_RES_CROSS");
      rc = (env != nullptr && atoi(env) == 1) ? 1 : 0;
    }
    p.res_cross = rc;
  }

  // BLOCK_N: the widest tile that the output fills (wide tiles read A once per 256 columns and keep the MMA off the
  // shared-memory read limit); wave quantisation is handled by stream-K below, not by shrinking tiles
  const long long sp_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * p.tiles_n;
  const int sms = sm_count();
  int block_n = a->n_out > 128 ? 256 : (a->n_out > 64 ? 128 : 64);
  // Split precision: 256-wide tiles only pay off for long K (tensor-bound: RPN 3x3 0.34 vs 0.41 ms, layer4 3x3 0.18 vs
  // 0.20); up to K = 2304 the 128-wide tile is 13..22 % faster on every trunk layer (three 64 KB stages instead of
  // two 96 KB ones, half-size epilogues, finer wave granularity) -- profiles/r01_gemm_blockn.txt.
  if (a->a_lo != nullptr && block_n == 256 && (io_f16 || static_cast<long long>(taps) * a->a_c <= 2304)) block_n = 128;
  // few M-tiles and a short K loop (the q/k projections of a single episode): narrower tiles spread the work over
  // more SMs at no extra cost -- stream-K's partial-tile exchange costs more than such a launch (measured: q-proj
  // 1900x256x1024 took 50 us as 15 stream-K'd 256-wide tiles)
  if (static_cast<long long>(taps) * a->a_c <= 2048) {
    while (block_n > 64 && sp_tiles * ((a->n_out + block_n - 1) / block_n) * 2 <= sms) block_n >>= 1;
  }
  // long K with few output tiles (the weight-gradient GEMMs of the training step: M = output channels, K = pixels; the
  // layers of a single 800x1333 episode): stream-K alone leaves every CTA a short K range plus a 128 x 256 fp32 partial to
  // exchange; narrower tiles give it whole tiles to spread first.  Measured on the res101 training step: 36.8 ms with
  // this rule off, 33.5 / 32.9 ms with every launch forced to 128 / 64 columns (tools/train_bench.py, DANA_BLOCK_N).
  if (a->a_lo != nullptr) {
    while (block_n > 64 && sp_tiles * ((a->n_out + block_n - 1) / block_n) < sms) block_n >>= 1;
  }
  if (softmax) block_n = a->softmax_ns <= 64 ? 64 : (a->softmax_ns <= 256 ? 256 : 512);   // one N-tile per shot segment
  const bool wide = softmax && block_n == 512;
  const int sm_half = wide ? ((a->softmax_ns + 31) / 32 * 32) / 2 : 0;   // % 16 == 0, <= 256
  if (dx3) block_n = a->n_out == 64 ? 64 : 128;
  {
    const char* env = getenv("DANA_BLOCK_N");
    if (env != nullptr && !softmax && !dx3) {
      const int v = atoi(env);
      if (v == 64 || v == 128 || v == 256) block_n = v;
    }
  }
  p.tiles_co = softmax ? (a->n_out + a->softmax_ns - 1) / a->softmax_ns : (a->n_out + block_n - 1) / block_n;
  p.sm_ns = softmax ? a->softmax_ns : 0;
  p.sm_pitch = softmax ? a->softmax_pitch : 0;
  p.sm_half = sm_half;
  // cluster of 2 CTAs along M sharing (multicasting) the weight tile: only where weights are shared across
  // tiles (not the per-image attention operands), the K loop is long enough to matter and tiles are 256 wide
  int cm = 1;
  {
    const long long k_total_ = static_cast<long long>(taps) * a->a_c;
    const int nsplit_ = (a->a_lo != nullptr) ? 2 : 1;
    // measured on B200 (tools/gemm_bench.py): no gain -- the big layers lose to wave quantisation, not to L2
    // bandwidth -- so the multicast path is opt-in (DANA_CLUSTER=2) and stream-K below is the default remedy
    const char* env = getenv("DANA_CLUSTER");
    if (env != nullptr && atoi(env) == 2 && block_n == 256 && !batched && !io_f16 && k_total_ >= 256 && sp_tiles >= 2) cm = 2;
    // experiment (DANA_CLUSTER=4): the fp16-plane layers issue a third of the MMAs per operand byte, so their weight
    // re-streaming weighs more: CTA pairs multicasting the weight tile
    if (env != nullptr && atoi(env) == 4 && block_n == 256 && nsplit_ == 1 && io_f16 && !batched && !softmax && k_total_ >= 256 &&
        sp_tiles >= 2)
      cm = 2;
    // experiment (DANA_CLUSTER=3): the 3x3 convolutions of the 64 / 128-channel stages re-stream their weights for every
    // 128-pixel tile and are bound by L2 -> SM traffic; a CTA pair sharing the weight tile halves that half of it
    if (env != nullptr && atoi(env) == 3 && nsplit_ == 2 && block_n <= 128 && taps == 9 && !batched && !io_f16 && !softmax &&
        sp_tiles >= 2 * sms)
      cm = 2;
  }

  // K extent of a pipeline stage (see ConvGemmCfg): 32 for the long-K / wide-N launches of the split mode
  const int nsplit = (a->a_lo != nullptr) ? 2 : 1;
  int ktile = 64;
  {
    const long long k_total_ = static_cast<long long>(taps) * a->a_c;
    if (nsplit == 2 && block_n == 256 && cm == 1 && !softmax && !io_f16 && (a->a_c % 32) == 0 && k_total_ >= 256 &&
        (a->n_out >= 512 || k_total_ >= 2048))
      ktile = 32;
    const char* env = getenv("DANA_KTILE");
    if (env != nullptr && nsplit == 2 && block_n == 256 && cm == 1 && !softmax && !io_f16 && (a->a_c % 32) == 0)
      ktile = atoi(env) == 32 ? 32 : 64;
    // (128-wide tiles stay at KT = 64: six 32 KB stages instead of three 64 KB ones measured 7 % slower on the step)
    if (dx3) ktile = (block_n == 64) ? 64 : 32;   // two stages of 88 KB / 68 KB (window box + three weight tiles, two planes)
    if (wide) ktile = 32;                         // two stages of (8 + 32) KB per plane
  }
  p.c_blocks = static_cast<int>((a->a_c + ktile - 1) / ktile);
  // tensor maps
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(a->a_c), static_cast<uint64_t>(a->a_w),
                              static_cast<uint64_t>(a->a_h), static_cast<uint64_t>(a->a_n)};
    const uint64_t str[3] = {static_cast<uint64_t>(a->a_sx) * 2, static_cast<uint64_t>(a->a_sy) * 2,
                             static_cast<uint64_t>(a->a_sn) * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(ktile), static_cast<uint32_t>(tc.bw),
                             static_cast<uint32_t>(dx3 ? tc.bh + 2 : tc.bh), static_cast<uint32_t>(tc.bn)};
    p.dx_box_bytes = dx3 ? tc.bw * (tc.bh + 2) * ktile * 2 : 0;
    int rc = encode_bf16_map(&p.tm_a_hi, a->a_hi, 4, dims, str, box, ktile == 32);
    if (rc != DANA_OK) return rc;
    if (nsplit == 2) {
      rc = encode_bf16_map(&p.tm_a_lo, a->a_lo, 4, dims, str, box, ktile == 32);
      if (rc != DANA_OK) return rc;
    }
  }
  {
    const uint64_t k_total = static_cast<uint64_t>(taps) * static_cast<uint64_t>(a->a_c);
    const uint64_t nb = batched ? static_cast<uint64_t>(a->out_n) : 1;
    const uint64_t dims[3] = {k_total, static_cast<uint64_t>(a->n_out), nb};
    const uint64_t bs = batched ? static_cast<uint64_t>(a->b_batch_stride)
                                : static_cast<uint64_t>(a->b_pitch) * static_cast<uint64_t>(a->n_out);
    const uint64_t str[2] = {static_cast<uint64_t>(a->b_pitch) * 2, ((bs * 2 + 15) / 16) * 16};
    const uint32_t box[3] = {static_cast<uint32_t>(ktile), static_cast<uint32_t>(wide ? sm_half : block_n / cm), 1};
    int rc = encode_bf16_map(&p.tm_b_hi, a->b_hi, 3, dims, str, box, ktile == 32);
    if (rc != DANA_OK) return rc;
    if (nsplit == 2) {
      rc = encode_bf16_map(&p.tm_b_lo, a->b_lo, 3, dims, str, box, ktile == 32);
      if (rc != DANA_OK) return rc;
    }
  }

  const long long num_work = ((sp_tiles + cm - 1) / cm) * p.tiles_co;
  const long long max_clusters = sms / cm;
  int grid = static_cast<int>((num_work < max_clusters ? num_work : max_clusters) * cm);
  // tile order (see the kernel's producer): N fast when the launch takes several waves of tiles, so that the N-tiles of
  // an M-tile share its activation reads in L2; a launch whose tiles are all in flight at once (stream-K over a few
  // long-K tiles, e.g. the P.V contraction of a single query) reads less with N slow (measured: 300 MB vs 586 MB)
  p.n_fast = (num_work > max_clusters) ? 1 : 0;
  {
    // L2 prefetch of the next tile's activation boxes (opt-in, DANA_A_PREFETCH=1; multi-wave launches only).
    // Measured: no layer gains, the 1x1 layers lose 3-20 % and the step drops from 584 to 554 images/s -- the extra
    // L2 requests compete with the loads they were meant to help.  Off by default.
    static int a_pf = -1;
    if (a_pf < 0) {
      const char* env = getenv("DANA_A_PREFETCH");
      a_pf = (env != nullptr && atoi(env) == 1) ? 1 : 0;
    }
    p.a_prefetch = (a_pf && num_work > max_clusters && !softmax) ? 1 : 0;
  }
  // stream-K when whole-tile scheduling would leave SMs idle (partial last round or fewer tiles than SMs)
  p.sk_epoch = 0;
  {
    const long long num_kb = static_cast<long long>(dx3 ? 3 : taps) * p.c_blocks;   // DX3: a k-block is a column shift (3 taps)
    const long long total_it = num_work * num_kb;
    const long long rounds = (num_work + sms - 1) / sms;
    const double eff = static_cast<double>(num_work) / static_cast<double>(rounds * sms);
    const int64_t need = 4096 + static_cast<int64_t>(sms) * 128 * block_n * 4;
    const char* env = getenv("DANA_STREAMK");
    const bool allowed = (env == nullptr) || atoi(env) != 0;
    // cost model (microseconds): a k-block costs ~0.8 us in x3 / ~0.27 us in bf16 at BLOCK_N = 256; stream-K pays
    // ~4 us for publishing / collecting partial tiles
    const double kb_us = (nsplit == 2 ? 0.8 : 0.27) * block_n / 256.0 * ktile / 64.0 * (dx3 ? 3.0 : 1.0);
    const double plain_us = static_cast<double>(rounds * num_kb) * kb_us;
    const double sk_us = static_cast<double>((total_it + sms - 1) / sms) * kb_us + 4.0;
    if (allowed && !softmax && cm == 1 && a->workspace != nullptr && a->workspace_bytes >= need && a->sk_epoch != 0 && eff < 0.9 &&
        sk_us < 0.85 * plain_us && total_it >= sms && sms <= 1024 && num_work * 4 >= 16) {
      p.sk_epoch = 1;   // flags are reset by the consumer, a constant "ready" value is enough
      p.sk_flags = static_cast<int*>(a->workspace);
      p.sk_partials = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + 4096);
      // at most ~4 CTAs per tile: collecting many partial tiles costs more than it balances
      const long long cap = num_work * 4;
      grid = static_cast<int>(cap < sms ? cap : sms);
    }
  }
  if (wide) {
    return nsplit == 2 ? launch_conv_gemm_v<512, 2, 4, 1, 32>(p, grid, stream) : launch_conv_gemm_v<512, 1, 4, 1, 32>(p, grid, stream);
  }
  if (dx3) {
    if (!fast_epilogue_ok(p, 2)) return DANA_ENOTSUP;   // checked before choosing the schedule would be nicer; trunk layers always pass
    if (block_n == 64) return launch_conv_gemm_v<64, 2, 1, 1, 64, true>(p, grid, stream);
    return launch_conv_gemm_v<128, 2, 1, 1, 32, true>(p, grid, stream);
  }
  if (nsplit == 1) {
    if (block_n == 256) return cm == 2 ? launch_conv_gemm<256, 1, 2>(p, grid, stream, io_f16) : launch_conv_gemm<256, 1, 1>(p, grid, stream, io_f16);
    if (block_n == 128) return launch_conv_gemm<128, 1, 1>(p, grid, stream, io_f16);
    return launch_conv_gemm<64, 1, 1>(p, grid, stream, io_f16);
  }
  if (block_n == 256 && ktile == 32) return launch_conv_gemm<256, 2, 1, 32>(p, grid, stream, io_f16);
  if (block_n == 256) return cm == 2 ? launch_conv_gemm<256, 2, 2>(p, grid, stream, io_f16) : launch_conv_gemm<256, 2, 1>(p, grid, stream, io_f16);
  if (block_n == 128) return cm == 2 ? launch_conv_gemm<128, 2, 2>(p, grid, stream, io_f16) : launch_conv_gemm<128, 2, 1>(p, grid, stream, io_f16);
  return cm == 2 ? launch_conv_gemm<64, 2, 2>(p, grid, stream, io_f16) : launch_conv_gemm<64, 2, 1>(p, grid, stream, io_f16);
}

}  // namespace dana
