// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// One kernel serves every GEMM-shaped op of the DAnA forward path:
//   * 1x1 convolutions (stride folded into the TMA view), 3x3 stride-1 pad-1 convolutions
//     (nine shifted TMA boxes; out-of-bounds zero fill is the padding)      [resnet.py:71-76, rpn.py:28]
//   * Linear layers and the attention contractions (W = rows, H = 1)          [dana.py:124,140,142,147]
//
// Activations are NHWC bf16.  In "x3" mode every operand is a (hi, lo) bf16 pair and each k-step
// issues hi*hi + hi*lo + lo*hi, which restores fp32-equivalent products with fp32 TMEM accumulation
// (the reference runs fp32 end to end; SURVEY.md section 7 hard part 1).
//
// Structure: persistent CTAs (one per SM), 6 warps.
//   warp 0      TMA producer: walks (tap, 64-channel block) k-blocks through a smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, double-buffered accumulators
//   warps 2..5  epilogue: tcgen05.ld -> scale/bias (folded BN), residual, ReLU -> bf16 hi/lo or fp32
#pragma once
#include "tc_common.cuh"

namespace dana {

struct ConvGemmParams {
  CUtensorMap tm_a_hi, tm_a_lo;  // 4-D (c, x, y, n), box (64, bw, bh, bn)
  CUtensorMap tm_b_hi, tm_b_lo;  // 3-D (k, co, batch), box (64, BLOCK_N, 1)
  int tiles_x, tiles_y, tiles_n, tiles_co;
  int bw, bh, bn;
  int taps_r, taps_s, pad_y, pad_x;
  int c_blocks;   // ceil(C_in / 64)
  int c_in;       // K extent per tap
  int n_out;      // output channels
  int out_w, out_h, out_n;
  long long so_x, so_y, so_n;  // output strides (elements)
  long long sr_x, sr_y, sr_n;  // residual strides (elements)
  int b_batched;
  long long bias_sn;  // per-n bias stride (elements), 0 = shared
  int relu;
  float alpha;
  const float* scale;
  const float* bias;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  const float* res_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;
};

constexpr int kGemmThreads = 192;
constexpr int kTileM = 128;
constexpr int kTileK = 64;
constexpr int kATileBytes = kTileM * kTileK * 2;  // 16 KB per plane

template <int BLOCK_N, int NSPLIT>
struct ConvGemmCfg {
  static constexpr int kBTileBytes = BLOCK_N * kTileK * 2;
  static constexpr int kStageBytes = NSPLIT * (kATileBytes + kBTileBytes);
  static constexpr int kBudget = 200 * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // two accumulators
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 2 * BLOCK_N * 4 /*scale,bias*/ + 256;
};

template <int BLOCK_N, int NSPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BLOCK_N, NSPLIT>;
  constexpr int kStages = Cfg::kStages;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(kStages >= 2, "pipeline too shallow");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;                                   // kStages * kStageBytes
  float* s_scale = reinterpret_cast<float*>(tiles + kStages * Cfg::kStageBytes);
  float* s_bias = s_scale + BLOCK_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + BLOCK_N);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_kb = p.taps_r * p.taps_s * p.c_blocks;
  const int sp_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int num_tiles = sp_tiles * p.tiles_co;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a_hi);
    tma_prefetch_desc(&p.tm_b_hi);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&p.tm_a_lo);
      tma_prefetch_desc(&p.tm_b_lo);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int co_t = t % p.tiles_co;
        const int sp = t / p.tiles_co;
        const int x0 = (sp % p.tiles_x) * p.bw;
        const int y0 = ((sp / p.tiles_x) % p.tiles_y) * p.bh;
        const int n0 = (sp / (p.tiles_x * p.tiles_y)) * p.bn;
        const int bcoord = p.b_batched ? n0 : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.c_blocks;
          const int cb = kb - tap * p.c_blocks;
          const int r = tap / p.taps_s;
          const int s = tap - r * p.taps_s;
          mbar_wait(&empty_bar[stage], phase ^ 1, 101);
          uint8_t* st = tiles + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int ka = cb * kTileK;
          const int kbk = tap * p.c_in + cb * kTileK;
          tma_load_4d(st, &p.tm_a_hi, &full_bar[stage], ka, x0 + s - p.pad_x, y0 + r - p.pad_y, n0);
          tma_load_3d(st + NSPLIT * kATileBytes, &p.tm_b_hi, &full_bar[stage], kbk, co_t * BLOCK_N, bcoord);
          if (NSPLIT == 2) {
            tma_load_4d(st + kATileBytes, &p.tm_a_lo, &full_bar[stage], ka, x0 + s - p.pad_x, y0 + r - p.pad_y, n0);
            tma_load_3d(st + 2 * kATileBytes + Cfg::kBTileBytes, &p.tm_b_lo, &full_bar[stage], kbk, co_t * BLOCK_N,
                        bcoord);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1, 102);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 103);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint32_t b_hi = a_hi + NSPLIT * kATileBytes;
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {
            const uint32_t koff = k * 32;  // 16 bf16 = 32 B inside the 128-B swizzle row
            const uint64_t da = umma_desc_sw128(a_hi + koff);
            const uint64_t db = umma_desc_sw128(b_hi + koff);
            if (NSPLIT == 2) {
              const uint64_t dal = umma_desc_sw128(a_hi + kATileBytes + koff);
              const uint64_t dbl = umma_desc_sw128(b_hi + Cfg::kBTileBytes + koff);
              // small cross terms first, leading term last
              umma_bf16(d_addr, dal, db, idesc, (kb | k) != 0);
              umma_bf16(d_addr, da, dbl, idesc, 1);
              umma_bf16(d_addr, da, db, idesc, 1);
            } else {
              umma_bf16(d_addr, da, db, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&acc_full[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;         // tile row
    const int et = threadIdx.x - 64;     // 0..127
    const int bw_i = m % p.bw;
    const int bh_i = (m / p.bw) % p.bh;
    const int bn_i = m / (p.bw * p.bh);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int co_t = t % p.tiles_co;
      const int sp = t / p.tiles_co;
      const int x = (sp % p.tiles_x) * p.bw + bw_i;
      const int y = ((sp / p.tiles_x) % p.tiles_y) * p.bh + bh_i;
      const int n = (sp / (p.tiles_x * p.tiles_y)) * p.bn + bn_i;
      const bool row_ok = (x < p.out_w) && (y < p.out_h) && (n < p.out_n);
      const int co0 = co_t * BLOCK_N;

      // stage per-channel scale / bias for this tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < BLOCK_N; i += 128) {
        const int co = co0 + i;
        s_scale[i] = (p.scale != nullptr && co < p.n_out) ? __ldg(p.scale + co) : 1.0f;
        s_bias[i] = (p.bias != nullptr && co < p.n_out)
                        ? __ldg(p.bias + static_cast<long long>(sp / (p.tiles_x * p.tiles_y)) * p.bn * p.bias_sn + co)
                        : 0.0f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&acc_full[acc], acc_phase, 104);
      tc_fence_after();
      const long long o_off = static_cast<long long>(n) * p.so_n + static_cast<long long>(y) * p.so_y +
                              static_cast<long long>(x) * p.so_x;
      const long long r_off = static_cast<long long>(n) * p.sr_n + static_cast<long long>(y) * p.sr_y +
                              static_cast<long long>(x) * p.sr_x;
      const bool have_res = (p.res_hi != nullptr) || (p.res_f32 != nullptr);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * BLOCK_N + c * 32);
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        const int cbase = co0 + c * 32;
        if (row_ok && cbase < p.n_out) {
          const bool full = (cbase + 32 <= p.n_out);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha * s_scale[c * 32 + j] + s_bias[c * 32 + j];
          if (have_res) {
            if (p.res_f32 != nullptr) {
              const float* rp = p.res_f32 + r_off + cbase;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || cbase + j < p.n_out) f[j] += __ldg(rp + j);
            } else {
              const __nv_bfloat16* rh = p.res_hi + r_off + cbase;
              const bool vec = full && ((reinterpret_cast<uintptr_t>(rh) & 15) == 0);
              if (vec) {
                const uint4* rh4 = reinterpret_cast<const uint4*>(rh);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const uint4 w = __ldg(rh4 + g);
                  const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    f[g * 8 + e * 2] += __uint_as_float(ws[e] << 16);
                    f[g * 8 + e * 2 + 1] += __uint_as_float(ws[e] & 0xFFFF0000u);
                  }
                }
                if (p.res_lo != nullptr) {
                  const uint4* rl4 = reinterpret_cast<const uint4*>(p.res_lo + r_off + cbase);
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const uint4 w = __ldg(rl4 + g);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      f[g * 8 + e * 2] += __uint_as_float(ws[e] << 16);
                      f[g * 8 + e * 2 + 1] += __uint_as_float(ws[e] & 0xFFFF0000u);
                    }
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (full || cbase + j < p.n_out) {
                    f[j] += __bfloat162float(rh[j]);
                    if (p.res_lo != nullptr) f[j] += __bfloat162float(p.res_lo[r_off + cbase + j]);
                  }
                }
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (p.out_f32 != nullptr) {
            float* op = p.out_f32 + o_off + cbase;
            if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
              float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
              for (int g = 0; g < 8; ++g) o4[g] = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || cbase + j < p.n_out) op[j] = f[j];
            }
          }
          if (p.out_hi != nullptr) {
            __nv_bfloat16* oh = p.out_hi + o_off + cbase;
            __nv_bfloat16* ol = (p.out_lo != nullptr) ? p.out_lo + o_off + cbase : nullptr;
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(f[2 * j], h0, l0);
              split_bf16(f[2 * j + 1], h1, l1);
              ph[j] = pack_bf16x2(h0, h1);
              pl[j] = pack_bf16x2(l0, l1);
            }
            if (full && ((reinterpret_cast<uintptr_t>(oh) & 15) == 0)) {
              uint4* o4 = reinterpret_cast<uint4*>(oh);
#pragma unroll
              for (int g = 0; g < 4; ++g) o4[g] = make_uint4(ph[g * 4], ph[g * 4 + 1], ph[g * 4 + 2], ph[g * 4 + 3]);
              if (ol != nullptr) {
                uint4* l4 = reinterpret_cast<uint4*>(ol);
#pragma unroll
                for (int g = 0; g < 4; ++g) l4[g] = make_uint4(pl[g * 4], pl[g * 4 + 1], pl[g * 4 + 2], pl[g * 4 + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (full || cbase + j < p.n_out) {
                  const uint32_t wh = ph[j >> 1], wl = pl[j >> 1];
                  oh[j] = __ushort_as_bfloat16(static_cast<unsigned short>((j & 1) ? (wh >> 16) : (wh & 0xFFFF)));
                  if (ol != nullptr)
                    ol[j] = __ushort_as_bfloat16(static_cast<unsigned short>((j & 1) ? (wl >> 16) : (wl & 0xFFFF)));
                }
              }
            }
          }
        }
      }
      // all tcgen05.ld of this warp have completed (wait::ld above): release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace dana
