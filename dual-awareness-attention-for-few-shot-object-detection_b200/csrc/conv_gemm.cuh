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
// Structure: persistent CTAs (one per SM), 10 warps.
//   warp 0      TMA producer: walks (tap, 64-channel block) k-blocks through a smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, double-buffered accumulators
//   warps 2..9  epilogue: tcgen05.ld -> scale/bias (folded BN), residual, ReLU -> bf16 hi/lo or fp32
#pragma once
#include "tc_common.cuh"

namespace dana {

struct ConvGemmParams {
  CUtensorMap tm_a_hi, tm_a_lo;  // 4-D (c, x, y, n), box (64, bw, bh, bn)
  CUtensorMap tm_b_hi, tm_b_lo;  // 3-D (k, co, batch), box (64, BLOCK_N, 1)
  int tiles_x, tiles_y, tiles_n, tiles_co;
  int bw, bh, bn;
  int lbw, lbh;   // log2(bw), log2(bh): the tile extents are powers of two
  int taps_r, taps_s, pad_y, pad_x;
  int c_blocks;   // ceil(C_in / K-tile)
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
  // stream-K (see the kernel comment): per-CTA fp32 partial-accumulator slots and ready flags
  float* sk_partials;   // [grid][128 * BLOCK_N]
  int* sk_flags;        // [grid]
  int sk_epoch;         // value a flag takes when this launch's partial is published; 0 = stream-K off
  int sm_ns, sm_pitch;  // SOFTMAX epilogue: keys per segment, column pitch of a segment in the output
  int sm_half;          // wide softmax tile (EPI 4): columns per MMA = rows per key box = ceil32(sm_ns) / 2
  int n_fast;           // tile order: 1 = the N-tiles of an M-tile are consecutive work items, 0 = N is the slow index
  int a_prefetch;       // producer prefetches the next tile's activation boxes into L2
  int ab_f16;           // operands are IEEE fp16 planes (kind::f16 with F16 formats) instead of bf16
  int dx_box_bytes;     // DX3: bytes of one window box per plane, bw * (bh + 2) rows of the K-tile
  int res_cross;        // epilogue: load the next tile's first residual chunk during this tile's last chunk
};

constexpr int kGemmThreads = 320;
constexpr int kTileM = 128;
// KT = K extent of one pipeline stage: 64 bf16 (128-byte swizzle rows) by default.  The split mode at BLOCK_N = 256
// has two planes per operand, i.e. only two 96 KB stages; its long-K launches use KT = 32 (64-byte swizzle rows,
// four 48 KB stages), which covers the TMA latency better (measured: RPN 3x3 conv 0.406 -> 0.357 ms).  Narrow tiles
// already have >= 3 stages and lose with the halved boxes (twice the TMA / barrier traffic), so they stay at 64.
// DX3: 3x3 stride-1 pad-1 convolution with the input window loaded ONCE PER COLUMN SHIFT instead of once per tap.
// The output tile is bw x bh pixels (bw = 8 or 16, bw * bh = 128, row m = y * bw + x).  For dx in {0,1,2} one TMA box
// of bw x (bh + 2) pixels starting at (x0 + dx - 1, y0 - 1) lands in shared memory as (bh + 2) * bw rows; the A
// operand of tap (dy, dx) is that box from row dy * bw on -- 128 consecutive rows whose start is a whole number of
// 8-row swizzle atoms, so the UMMA descriptor needs nothing special.  A pipeline stage holds one such box and the
// three weight tiles of its taps: 3 window loads per channel block instead of 9, i.e. 1.1x instead of 9x the
// activation bytes through L2 -> SM (ncu on layer1's 3x3: 530 MB pulled through TMA for 77 MB of operands).
template <int BLOCK_N, int NSPLIT, int KT = 64, bool DX3 = false>
struct ConvGemmCfg {
  static constexpr int kTileK = KT;
  static constexpr int kARows = DX3 ? 160 : kTileM;          // 8 x 18 = 144 or 16 x 10 = 160 window rows
  static constexpr int kBTaps = DX3 ? 3 : 1;
  static constexpr int kATileBytes = kARows * kTileK * 2;   // per plane
  static constexpr int kBTileBytes = BLOCK_N * kTileK * 2;  // per plane and tap
  static constexpr int kBGroupBytes = kBTaps * kBTileBytes;
  static constexpr int kStageBytes = NSPLIT * (kATileBytes + kBGroupBytes);
  static constexpr int kBudget = (BLOCK_N > 256 ? 184 : 200) * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  // two accumulators; the 512-column tile (wide softmax epilogue) owns all of TMEM with a single one
  static constexpr int kAccBufs = BLOCK_N > 256 ? 1 : 2;
  static constexpr int kTmemCols = BLOCK_N > 256 ? 512 : (2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
  static constexpr int kXposeBytes = 8 * 4096;  // one 32x32 fp32 transposition buffer per epilogue warp
  // softmax epilogue: (row max, row sum) exchange, 8 warps x 64 floats; it overlays the scale/bias arrays (unused
  // by that epilogue) when they are large enough -- the <256, 2> budget has no 2 KB to spare
  // (the softmax epilogue exists at BLOCK_N = 64 and 256 only; the <128, 2> budget is full as well)
  static constexpr int kXchgBytes = (BLOCK_N <= 64) ? 8 * 64 * 4 : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kXposeBytes + 2 * BLOCK_N * 4 /*scale,bias*/ + kXchgBytes + 256;
};

// FAST epilogue: host-verified n_out % 32 == 0, bf16 pair output (no fp32 output / residual), every
// pointer 16-byte aligned and every stride a multiple of 8 elements -> no per-element predicates.
// Work distribution.  Classic persistent scheduling hands out whole 128 x BLOCK_N tiles; with 150 tiles on 148
// SMs that is two rounds at 51 % utilisation, with 75 tiles half the SMs idle.  With stream-K (sk_epoch != 0) the
// linearised (tile, k-block) iteration space is cut into gridDim.x equal contiguous ranges instead.  A CTA whose
// range ends inside a tile dumps its raw fp32 accumulators into its workspace slot and raises its flag; the CTA
// that owns the tile's last k-block adds the slots of the (lower-numbered) CTAs that covered the tile's head and
// runs the normal epilogue.  Every CTA publishes at most one partial, waits only on lower-numbered CTAs, and
// all CTAs are co-resident (grid <= SM count), so the wait cannot deadlock.
struct SkRange {
  long long it, it_end;
  int num_kb;
  // Next segment of this CTA (tile index, k-block range [kb0, kb1)), walking the range BACKWARDS: the tail
  // segment (a tile this CTA only starts) is computed and published first, the head segment (a tile whose
  // beginning a lower CTA computes) is finished last, so by the time a partial is needed it has long been
  // written -- in the forward order every CTA would wait for its predecessor's whole range (a serial chain).
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (it >= it_end) return false;
    tile = static_cast<int>((it_end - 1) / num_kb);
    const long long tile_start = static_cast<long long>(tile) * num_kb;
    kb1 = static_cast<int>(it_end - tile_start);
    kb0 = (tile_start >= it) ? 0 : static_cast<int>(it - tile_start);
    it_end = tile_start + kb0;
    return true;
  }
};

// CM: thread-block cluster size along M.  The CM CTAs of a cluster work on CM consecutive M-tiles of the same
// N-tile; each loads 1/CM of the weight (B) tile and TMA-multicasts it to all of them, which divides the L2->SM
// weight traffic by CM (the compute-bound layers are L2-bandwidth bound at 128x256 tiles).  A smem slot is
// released cluster-wide: every CTA's MMA commit arrives on the `empty` barrier of all CM CTAs.
// EPI: 0 generic epilogue, 1 FAST, 2 SOFTMAX -- the attention-logits contraction with the row softmax fused in
// (dana.py:142-143,273-274): N-tile t is shot t's segment of sm_ns keys (sm_ns <= BLOCK_N), the epilogue takes
// max / sum over the segment straight from TMEM and writes the normalised probabilities as a bf16 pair at
// column t*sm_pitch (pad columns zeroed); the fp32 logits never leave the SM.
// EPI 3: FAST with single-plane fp16 output / residual (out_hi, res_hi are __half planes, no lo plane) whatever the
// operand planes are: the producer of an fp16-operand layer (the layers that run one MMA per product, see
// DESIGN.md section 3) writes the plane its consumer feeds to the tensor cores.  Values saturate at +-65504.
template <int BLOCK_N, int NSPLIT, int EPI, int CM, int KT = 64, bool DX3 = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  static_assert(!DX3 || (CM == 1 && EPI != 2), "DX3: single-CTA clusters, no softmax epilogue");
  constexpr bool FAST = (EPI == 1 || EPI == 3);
  constexpr bool F16IO = (EPI == 3);
  constexpr bool OUT2 = (NSPLIT == 2) && !F16IO;   // output / residual carry a lo plane
  constexpr bool SOFTMAX = (EPI == 2 || EPI == 4);
  // EPI 4: the same fused attention softmax for segments of 257..512 keys (the reference's 20 x 20 supports: 400).  A
  // segment does not fit one 256-column MMA, nor two accumulators TMEM: the tile is ONE accumulator of up to 512
  // columns written by two MMAs of sm_half = ceil32(ns) / 2 columns per k-step, and the epilogue makes three passes
  // over it straight from TMEM (row max, sum of exponentials, normalised probabilities) instead of holding the
  // segment in registers.  No fp32 logits in HBM, no separate softmax kernel.
  constexpr bool SOFTMAXW = (EPI == 4);
  static_assert(!SOFTMAXW || (BLOCK_N == 512 && KT == 32 && CM == 1 && !DX3), "wide softmax: 512-column tile, 32-wide K stages");
  constexpr int kAccBufs = ConvGemmCfg<BLOCK_N, NSPLIT, KT, DX3>::kAccBufs;
  using Cfg = ConvGemmCfg<BLOCK_N, NSPLIT, KT, DX3>;
  constexpr int kTileK = Cfg::kTileK;
  constexpr int kATileBytes = Cfg::kATileBytes;
  constexpr int kStages = Cfg::kStages;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && (BLOCK_N <= 256 || SOFTMAXW), "BLOCK_N");
  static_assert(kStages >= 2, "pipeline too shallow");
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "shared-memory budget");

  // No alignment slack: the budget of the <256, 2> configuration is within 1 KB of the 227 KB limit.  The dynamic
  // window starts at offset 0 of the CTA's shared memory (the kernel has no static __shared__), which is
  // 1024-byte aligned as SWIZZLE_128B needs; checked below.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tiles = smem_raw;                               // kStages * kStageBytes
  float* s_xpose = reinterpret_cast<float*>(tiles + kStages * Cfg::kStageBytes);   // [8 warps][32][32]
  float* s_scale = s_xpose + Cfg::kXposeBytes / 4;
  float* s_bias = s_scale + BLOCK_N;
  float* s_xchg = (Cfg::kXchgBytes == 0) ? s_scale : s_bias + BLOCK_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + BLOCK_N + Cfg::kXchgBytes / 4);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_kb = (DX3 ? 3 : p.taps_r * p.taps_s) * p.c_blocks;
  const int sp_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  // work item w -> (N-tile, group of CM consecutive M-tiles); this CTA takes M-tile group*CM + rank, which may
  // lie past the end (phantom tile: TMA zero-fills, the epilogue masks every row) so that all CTAs of a cluster
  // run the same pipeline schedule
  const int cm_rank = (CM > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CM;
  const int num_clusters = gridDim.x / CM;
  const int m_groups = (sp_tiles + CM - 1) / CM;
  const int num_work = m_groups * p.tiles_co;
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << CM) - 1u);
  const bool stream_k = (CM == 1) && (p.sk_epoch != 0);
  const long long total_it = static_cast<long long>(num_work) * num_kb;
  auto my_range = [&]() {
    SkRange r;
    r.num_kb = num_kb;
    if (stream_k) {
      r.it = total_it * blockIdx.x / gridDim.x;
      r.it_end = total_it * (blockIdx.x + 1) / gridDim.x;
    } else {
      r.it = 0;
      r.it_end = 0;
    }
    return r;
  };

  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) {
    atomicExch(&g_device_error, 105);
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a_hi);
    tma_prefetch_desc(&p.tm_b_hi);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&p.tm_a_lo);
      tma_prefetch_desc(&p.tm_b_lo);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CM);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CM > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (see the launcher): everything above touched only shared memory, TMEM and the
  // kernel parameters.  Wait for the preceding grids (their global writes are visible afterwards), then let the
  // next launch in the stream be scheduled as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SkRange rng = my_range();
      int wk = cluster_id - num_clusters, kb_lo = 0, kb_hi = num_kb;
      while (stream_k ? rng.next(wk, kb_lo, kb_hi) : ((wk += num_clusters) < num_work)) {
        // N-tile is the FAST index: the tiles_co N-tiles of one M-tile run on neighbouring CTAs at the same time, so
        // the activation tile (and, for 3x3 convs, its nine shifted views) is fetched from DRAM once and re-read from
        // L2; the weights are small and L2-resident either way.  (With N slow, every pass over M re-streamed the
        // whole activation: ncu showed 543 MB of DRAM traffic for the 350 MB layer4 conv3, 559 MB for the 136 MB RPN conv.)
        // (launches that fit one wave -- every tile in flight at once under stream-K -- keep N slow: n_fast = 0)
        const int mg = p.n_fast ? wk / p.tiles_co : wk % m_groups;
        const int co_t = p.n_fast ? wk - mg * p.tiles_co : wk / m_groups;
        const int sp = mg * CM + cm_rank;
        const int x0 = (sp % p.tiles_x) * p.bw;
        const int y0 = ((sp / p.tiles_x) % p.tiles_y) * p.bh;
        const int n0 = (sp / (p.tiles_x * p.tiles_y)) * p.bn;
        const int bcoord = p.b_batched ? n0 : 0;
        if (p.a_prefetch && !stream_k) {
          // Experiment kept as an opt-in (see the launcher: measured slower).  The stages hold at most ~190 KB in flight
          // per SM; when a stage contains DRAM misses its latency is 2-3 us and the main loop of the narrow layers runs
          // at 190 KB / 3.4 us (ncu: the epilogue waits on the accumulator barrier, L2 at 30 %).  This asks L2 for the
          // NEXT tile's activation boxes (centre tap only: the other taps are the same pixels shifted) a tile ahead.
          const int wn = wk + num_clusters;
          if (wn < num_work) {
            const int mgn = p.n_fast ? wn / p.tiles_co : wn % m_groups;
            const int spn = mgn * CM + cm_rank;
            const int xn = (spn % p.tiles_x) * p.bw;
            const int yn = ((spn / p.tiles_x) % p.tiles_y) * p.bh;
            const int nn = (spn / (p.tiles_x * p.tiles_y)) * p.bn;
            const bool same_a = p.n_fast && (wn / p.tiles_co == wk / p.tiles_co);   // next item: same M-tile, other N-tile
            if (!same_a) {
              const int taps = p.taps_r * p.taps_s;
              const int t0 = (taps == 9) ? 4 : 0, t1 = (taps == 9) ? 5 : taps;
              for (int tap = t0; tap < t1; ++tap) {
                const int r = tap / p.taps_s, s2 = tap - r * p.taps_s;
                for (int cb = 0; cb < p.c_blocks; ++cb) {
                  tma_prefetch_4d(&p.tm_a_hi, cb * kTileK, xn + s2 - p.pad_x, yn + r - p.pad_y, nn);
                  if (NSPLIT == 2) tma_prefetch_4d(&p.tm_a_lo, cb * kTileK, xn + s2 - p.pad_x, yn + r - p.pad_y, nn);
                }
              }
            }
          }
        }
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          const int tap = kb / p.c_blocks;
          const int cb = kb - tap * p.c_blocks;
          mbar_wait(&empty_bar[stage], phase ^ 1, 101);
          uint8_t* st = tiles + stage * Cfg::kStageBytes;
          const int ka = cb * kTileK;
          if constexpr (DX3) {
            // `tap` is the column shift dx: one window box (all three row shifts) + the weight tiles of taps (dy, dx)
            mbar_arrive_expect_tx(&full_bar[stage], NSPLIT * (p.dx_box_bytes + Cfg::kBGroupBytes));
            tma_load_4d(st, &p.tm_a_hi, &full_bar[stage], ka, x0 + tap - 1, y0 - 1, n0);
            if (NSPLIT == 2) tma_load_4d(st + kATileBytes, &p.tm_a_lo, &full_bar[stage], ka, x0 + tap - 1, y0 - 1, n0);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int kbk = (dy * 3 + tap) * p.c_in + ka;
              tma_load_3d(st + NSPLIT * kATileBytes + dy * Cfg::kBTileBytes, &p.tm_b_hi, &full_bar[stage], kbk,
                          co_t * BLOCK_N, 0);
              if (NSPLIT == 2)
                tma_load_3d(st + NSPLIT * kATileBytes + Cfg::kBGroupBytes + dy * Cfg::kBTileBytes, &p.tm_b_lo,
                            &full_bar[stage], kbk, co_t * BLOCK_N, 0);
            }
          } else if constexpr (SOFTMAXW) {
            // A k-block + the segment's keys as two boxes of sm_half rows (a TMA box has at most 256 rows)
            const int hb = p.sm_half * (kTileK * 2);
            mbar_arrive_expect_tx(&full_bar[stage], NSPLIT * (kATileBytes + 2 * hb));
            tma_load_4d(st, &p.tm_a_hi, &full_bar[stage], ka, x0, y0, n0);
            if (NSPLIT == 2) tma_load_4d(st + kATileBytes, &p.tm_a_lo, &full_bar[stage], ka, x0, y0, n0);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              tma_load_3d(st + NSPLIT * kATileBytes + hf * hb, &p.tm_b_hi, &full_bar[stage], ka,
                          co_t * p.sm_ns + hf * p.sm_half, bcoord);
              if (NSPLIT == 2)
                tma_load_3d(st + NSPLIT * kATileBytes + Cfg::kBGroupBytes + hf * hb, &p.tm_b_lo, &full_bar[stage], ka,
                            co_t * p.sm_ns + hf * p.sm_half, bcoord);
            }
          } else {
          const int r = tap / p.taps_s;
          const int s = tap - r * p.taps_s;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int kbk = tap * p.c_in + cb * kTileK;
          tma_load_4d(st, &p.tm_a_hi, &full_bar[stage], ka, x0 + s - p.pad_x, y0 + r - p.pad_y, n0);
          if (CM == 1) {
            tma_load_3d(st + NSPLIT * kATileBytes, &p.tm_b_hi, &full_bar[stage], kbk, SOFTMAX ? co_t * p.sm_ns : co_t * BLOCK_N, bcoord);
          } else {
            tma_load_3d_mc(st + NSPLIT * kATileBytes + cm_rank * (Cfg::kBTileBytes / CM), &p.tm_b_hi, &full_bar[stage],
                           kbk, co_t * BLOCK_N + cm_rank * (BLOCK_N / CM), 0, kMcMask);
          }
          if (NSPLIT == 2) {
            tma_load_4d(st + kATileBytes, &p.tm_a_lo, &full_bar[stage], ka, x0 + s - p.pad_x, y0 + r - p.pad_y, n0);
            if (CM == 1) {
              tma_load_3d(st + 2 * kATileBytes + Cfg::kBTileBytes, &p.tm_b_lo, &full_bar[stage], kbk,
                          SOFTMAX ? co_t * p.sm_ns : co_t * BLOCK_N, bcoord);
            } else {
              tma_load_3d_mc(st + 2 * kATileBytes + Cfg::kBTileBytes + cm_rank * (Cfg::kBTileBytes / CM), &p.tm_b_lo,
                             &full_bar[stage], kbk, co_t * BLOCK_N + cm_rank * (BLOCK_N / CM), 0, kMcMask);
            }
          }
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
    // The whole warp walks the loop on warp-uniform values (so the shared-memory descriptors sit in uniform
    // registers and a k-step costs one add per operand); one elected lane issues the MMAs and the commits.
    {
      const int mma_n = SOFTMAXW ? p.sm_half : BLOCK_N;
      const uint32_t idesc = p.ab_f16 ? umma_idesc_f16(kTileM, mma_n) : umma_idesc_bf16(kTileM, mma_n);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tiles_u = __shfl_sync(0xffffffffu, smem_u32(tiles), 0);
      // descriptor of a tile = constant high word + (address >> 4); a 16-element k-step is 32 B = +2
      constexpr uint64_t kDescHi = (kTileK == 64)
          ? ((static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61))
          : ((static_cast<uint64_t>(512 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(4) << 61));
      auto desc_of = [&](uint32_t addr) -> uint64_t {
        return kDescHi | static_cast<uint64_t>(((addr & 0x3FFFFu) >> 4) | (1u << 16));
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SkRange rng = my_range();
      int wk = cluster_id - num_clusters, kb_lo = 0, kb_hi = num_kb;
      while (stream_k ? rng.next(wk, kb_lo, kb_hi) : ((wk += num_clusters) < num_work)) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1, 102);
        tc_fence_after();
        const uint32_t d_addr = tmem_u + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full_bar[stage], phase, 103);
          tc_fence_after();
          const uint32_t a_hi = tiles_u + static_cast<uint32_t>(stage * Cfg::kStageBytes);
          const uint32_t b_hi = a_hi + NSPLIT * kATileBytes;
          if (elect_one()) {
#pragma unroll
            for (int dy = 0; dy < Cfg::kBTaps; ++dy) {
              // DX3: tap (dy, dx) reads the window box from row dy * bw on (a whole number of 8-row swizzle atoms)
              const uint32_t a_t = a_hi + (DX3 ? static_cast<uint32_t>(dy * p.bw * (kTileK * 2)) : 0u);
              const uint32_t b_t = b_hi + static_cast<uint32_t>(dy * Cfg::kBTileBytes);
#pragma unroll
              for (int hf = 0; hf < (SOFTMAXW ? 2 : 1); ++hf) {
                // wide softmax tile: columns [hf * sm_half, (hf + 1) * sm_half) from the hf-th box of keys
                const uint32_t d_h = d_addr + (SOFTMAXW ? static_cast<uint32_t>(hf * p.sm_half) : 0u);
                const uint32_t b_h = b_t + (SOFTMAXW ? static_cast<uint32_t>(hf * p.sm_half * (kTileK * 2)) : 0u);
                const uint64_t da0 = desc_of(a_t), db0 = desc_of(b_h);
                const uint64_t dal0 = desc_of(a_t + kATileBytes), dbl0 = desc_of(b_h + Cfg::kBGroupBytes);
#pragma unroll
                for (int k = 0; k < kTileK / 16; ++k) {
                  const uint32_t first = static_cast<uint32_t>((kb - kb_lo) | k | dy);
                  if (NSPLIT == 2) {
                    // small cross terms first, leading term last
                    umma_bf16(d_h, dal0 + 2 * k, db0 + 2 * k, idesc, first);
                    umma_bf16(d_h, da0 + 2 * k, dbl0 + 2 * k, idesc, 1);
                    umma_bf16(d_h, da0 + 2 * k, db0 + 2 * k, idesc, 1);
                  } else {
                    umma_bf16(d_h, da0 + 2 * k, db0 + 2 * k, idesc, first);
                  }
                }
              }
            }
            if (CM == 1) {
              umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            } else {
              umma_commit_mc(&empty_bar[stage], kMcMask);  // ... in every CTA of the cluster
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&acc_full[acc]);  // accumulator ready for the epilogue
        __syncwarp();
        if (++acc == kAccBufs) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter; each owns half of the tile's 32-column chunks.  A chunk goes
    //   TMEM --tcgen05.ld--> registers (one row per lane) --scale/bias--> 32x32 fp32 smem tile (16-byte XOR
    //   swizzle, conflict free both ways) --> registers (8 consecutive channels per lane, 4 lanes per row)
    // so that every global access of the warp is row-contiguous: the residual read and the bf16 hi/lo (or
    // fp32) store move 64..128 contiguous bytes per row instead of 16 bytes at 32 different rows.
    // The residual of the next chunk is requested before the current one is processed.
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;                // which half of the chunks
    const int et = threadIdx.x - 64;                 // 0..255
    float* tb = s_xpose + (warp - 2) * 1024;         // this warp's transposition tile
    constexpr int kChunks = BLOCK_N / 32;
    constexpr int kCpw = kChunks >= 2 ? kChunks / 2 : 1;   // chunks per warp
    const int c_begin = half * kCpw;
    const int sub = lane >> 2;                       // row within an 8-row group (phase B)
    const int cg = (lane & 3) * 8;                   // first channel of this lane's group (phase B)
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    const bool res_pair = (p.res_hi != nullptr);
    const bool res_lo = (p.res_lo != nullptr);
    const bool res_vec = res_pair && al16(p.res_hi) && (!res_lo || al16(p.res_lo)) && (p.sr_x % 8 == 0) &&
                         (p.sr_y % 8 == 0) && (p.sr_n % 8 == 0);
    const bool out_vec = (p.out_hi == nullptr || (al16(p.out_hi) && (p.out_lo == nullptr || al16(p.out_lo)) &&
                                                  (p.so_x % 8 == 0) && (p.so_y % 8 == 0) && (p.so_n % 8 == 0))) &&
                         (p.out_f32 == nullptr || (al16(p.out_f32) && (p.so_x % 4 == 0) && (p.so_y % 4 == 0) &&
                                                   (p.so_n % 4 == 0))) &&
                         (p.res_f32 == nullptr || (al16(p.res_f32) && (p.sr_x % 4 == 0) && (p.sr_y % 4 == 0) &&
                                                   (p.sr_n % 4 == 0)));
    int acc = 0;
    uint32_t acc_phase = 0;
    int staged_co = -1, staged_n = -1;
    // residual look-ahead registers; they live across tiles: the loads of a tile's FIRST chunk are issued while the
    // previous tile's last chunk is processed (pre_valid), so that no chunk meets its residual with zero lead -- with
    // one or two chunks per warp and tile that was every chunk / every other chunk (ncu: long_scoreboard on the first
    // use of the residual words was the top stall of the residual layers)
    uint4 rh[4], rl[4];
    bool pre_valid = false;
    const bool cross_tile = (p.res_cross != 0);
    SkRange rng = my_range();
    int wk = cluster_id - num_clusters, kb_lo = 0, kb_hi = num_kb;
    while (stream_k ? rng.next(wk, kb_lo, kb_hi) : ((wk += num_clusters) < num_work)) {
      if (stream_k && kb_hi < num_kb) {
        // ---- this CTA's range ends inside the tile: publish the raw accumulators and move on
        mbar_wait(&acc_full[acc], acc_phase, 104);
        tc_fence_after();
        float* slot = p.sk_partials + static_cast<long long>(blockIdx.x) * (128 * BLOCK_N);
#pragma unroll
        for (int ci = 0; ci < kCpw; ++ci) {
          const int c = c_begin + ci;
          if (c >= kChunks) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N + c * 32), v);
          tmem_ld_wait();
          float* dst = slot + (c * 4 + q) * 1024 + lane;   // [chunk][quarter][column j][lane]: coalesced per j
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j * 32] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          __threadfence();
          atomicExch(p.sk_flags + blockIdx.x, p.sk_epoch);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == kAccBufs) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      // contributors to the head of this tile (stream-K, range started inside the tile): CTAs below this one
      int sk_first = blockIdx.x, sk_last = blockIdx.x;   // [sk_first, sk_last) = contributing CTA ids
      if (stream_k && kb_lo > 0) {
        const long long tile_start = static_cast<long long>(wk) * num_kb;
        int cprev = static_cast<int>(blockIdx.x) - 1;
        while (cprev >= 0) {
          const long long b0 = total_it * cprev / gridDim.x, b1 = total_it * (cprev + 1) / gridDim.x;
          if (b1 <= tile_start) break;
          sk_first = cprev;
          if (b0 <= tile_start) break;
          --cprev;
        }
        for (int cc = sk_first + static_cast<int>(lane); cc < sk_last; cc += 32) {
          const long long t0 = clock64();
          while (atomicAdd(p.sk_flags + cc, 0) != p.sk_epoch) {
            if (clock64() - t0 > 8000000000LL) {
              atomicExch(&g_device_error, 106);
              __trap();
            }
          }
        }
        __syncwarp();
        __threadfence();
      }
      // tile coordinates: integer divisions by runtime values are ~25 instructions each, so lane 0 does them
      // once and the warp shares the results
      int co_t = 0, tx0 = 0, ty0 = 0, tn0 = 0;
      if (lane == 0) {
        const int mg = p.n_fast ? wk / p.tiles_co : wk % m_groups;   // tile order: see the producer
        co_t = p.n_fast ? wk - mg * p.tiles_co : wk / m_groups;
        const int sp = mg * CM + cm_rank;
        const int txy = p.tiles_x * p.tiles_y;
        const int tn = sp / txy;
        const int rem = sp - tn * txy;
        const int ty = rem / p.tiles_x;
        tx0 = (rem - ty * p.tiles_x) * p.bw;
        ty0 = ty * p.bh;
        tn0 = tn * p.bn;
      }
      co_t = __shfl_sync(0xffffffffu, co_t, 0);
      tx0 = __shfl_sync(0xffffffffu, tx0, 0);
      ty0 = __shfl_sync(0xffffffffu, ty0, 0);
      tn0 = __shfl_sync(0xffffffffu, tn0, 0);
      const int co0 = co_t * BLOCK_N;
      bool have_next = false;                       // next work item of this CTA (FAST + residual + whole-tile schedule)
      int n_co = 0, n_x0 = 0, n_y0 = 0, n_n0 = 0;
      if constexpr (FAST) {
        // Residual stream: the layers that carry a residual are the HBM-bound ones (K = 64..512), and their
        // epilogue stalled on the residual loads (ncu: long_scoreboard on the first use of the prefetched words --
        // one chunk of look-ahead does not cover DRAM latency).  Pull the NEXT tile's residual rows into L2 now,
        // a whole tile ahead: 2 threads per row, 128-byte lines, no registers held.
        if (p.res_hi != nullptr && !stream_k) {
          const int wn = wk + num_clusters;
          if (wn < num_work) {
            int nco = 0, nx0 = 0, ny0 = 0, nn0 = 0;
            have_next = true;
            if (lane == 0) {
              const int mg = p.n_fast ? wn / p.tiles_co : wn % m_groups;
              nco = p.n_fast ? wn - mg * p.tiles_co : wn / m_groups;
              const int sp = mg * CM + cm_rank;
              const int txy = p.tiles_x * p.tiles_y;
              const int tn = sp / txy;
              const int rem = sp - tn * txy;
              const int ty = rem / p.tiles_x;
              nx0 = (rem - ty * p.tiles_x) * p.bw;
              ny0 = ty * p.bh;
              nn0 = tn * p.bn;
            }
            nco = __shfl_sync(0xffffffffu, nco, 0);
            nx0 = __shfl_sync(0xffffffffu, nx0, 0);
            ny0 = __shfl_sync(0xffffffffu, ny0, 0);
            nn0 = __shfl_sync(0xffffffffu, nn0, 0);
            n_co = nco; n_x0 = nx0; n_y0 = ny0; n_n0 = nn0;
            const int m = et >> 1;
            const int x = nx0 + (m & (p.bw - 1));
            const int y = ny0 + ((m >> p.lbw) & (p.bh - 1));
            const int n = nn0 + (m >> (p.lbw + p.lbh));
            if (x < p.out_w && y < p.out_h && n < p.out_n) {
              const long long off = static_cast<long long>(n) * p.sr_n + static_cast<long long>(y) * p.sr_y +
                                    static_cast<long long>(x) * p.sr_x + nco * BLOCK_N + (et & 1) * (BLOCK_N / 2);
              const int cols = min(BLOCK_N / 2, p.n_out - nco * BLOCK_N - (et & 1) * (BLOCK_N / 2));
              for (int c = 0; c < cols; c += 64) {   // 64 bf16 = one 128-byte line
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_hi + off + c));
                if (OUT2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res_lo + off + c));
              }
            }
          }
        }
      }
      // the four rows this lane serves in phase B
      long long o_off[4], r_off[4];
      bool row_ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = q * 32 + 8 * i + sub;
        const int x = tx0 + (m & (p.bw - 1));
        const int y = ty0 + ((m >> p.lbw) & (p.bh - 1));
        const int n = tn0 + (m >> (p.lbw + p.lbh));
        row_ok[i] = (x < p.out_w) && (y < p.out_h) && (n < p.out_n);
        o_off[i] = static_cast<long long>(n) * p.so_n + static_cast<long long>(y) * p.so_y +
                   static_cast<long long>(x) * p.so_x;
        r_off[i] = static_cast<long long>(n) * p.sr_n + static_cast<long long>(y) * p.sr_y +
                   static_cast<long long>(x) * p.sr_x;
      }

      // FAST path: per-row base pointers (first chunk of this warp, this lane's 8-channel group); chunk ci is
      // then a compile-time offset of ci*32 elements, the lo plane a uniform delta
      const __nv_bfloat16* rhp[4];
      __nv_bfloat16* ohp[4];
      const long long d_res = (FAST && res_pair && OUT2) ? (p.res_lo - p.res_hi) : 0;
      const long long d_out = (FAST && OUT2) ? (p.out_lo - p.out_hi) : 0;
      if constexpr (FAST) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rhp[i] = res_pair ? p.res_hi + r_off[i] + co0 + c_begin * 32 + cg : nullptr;
          ohp[i] = p.out_hi + o_off[i] + co0 + c_begin * 32 + cg;
        }
      }
      auto prefetch = [&](int c) {
        if constexpr (FAST) {
          if (res_pair && co0 + c * 32 < p.n_out) {
            const int ofs = (c - c_begin) * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (row_ok[i]) {
                rh[i] = __ldg(reinterpret_cast<const uint4*>(rhp[i] + ofs));
                if (OUT2) rl[i] = __ldg(reinterpret_cast<const uint4*>(rhp[i] + d_res + ofs));
              }
            }
          }
          return;
        }
        const int col = co0 + c * 32 + cg;
        if (res_vec && col + 8 <= p.n_out) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (row_ok[i]) {
              rh[i] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + r_off[i] + col));
              if (res_lo) rl[i] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + r_off[i] + col));
            }
          }
        }
      };
      if (c_begin < kChunks && !pre_valid) prefetch(c_begin);
      pre_valid = false;
      // next tile's first chunk of this warp (issued from the last chunk of this tile, see above)
      auto prefetch_next = [&]() {
        if constexpr (FAST) {
          if (!(cross_tile && have_next && res_pair) || n_co * BLOCK_N + c_begin * 32 >= p.n_out || c_begin >= kChunks) return;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = q * 32 + 8 * i + sub;
            const int x = n_x0 + (m & (p.bw - 1));
            const int y = n_y0 + ((m >> p.lbw) & (p.bh - 1));
            const int n = n_n0 + (m >> (p.lbw + p.lbh));
            if (x < p.out_w && y < p.out_h && n < p.out_n) {
              const __nv_bfloat16* rp = p.res_hi + static_cast<long long>(n) * p.sr_n + static_cast<long long>(y) * p.sr_y +
                                        static_cast<long long>(x) * p.sr_x + n_co * BLOCK_N + c_begin * 32 + cg;
              rh[i] = __ldg(reinterpret_cast<const uint4*>(rp));
              if (OUT2) rl[i] = __ldg(reinterpret_cast<const uint4*>(rp + d_res));
            }
          }
          pre_valid = true;
        }
      };

      // stage per-channel scale (x alpha) / bias -- only when the N-tile (or, with per-image bias, the image)
      // changes, which with N-major tile order is once or twice per CTA
      if (!SOFTMAX && (co_t != staged_co || (p.bias_sn != 0 && tn0 != staged_n))) {
        staged_co = co_t;
        staged_n = tn0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < BLOCK_N; i += 256) {
          const int co = co0 + i;
          s_scale[i] = ((p.scale != nullptr && co < p.n_out) ? __ldg(p.scale + co) : 1.0f) * p.alpha;
          s_bias[i] = (p.bias != nullptr && co < p.n_out)
                          ? __ldg(p.bias + static_cast<long long>(tn0) * p.bias_sn + co)
                          : 0.0f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }

      mbar_wait(&acc_full[acc], acc_phase, 104);
      tc_fence_after();
      if constexpr (SOFTMAXW) {
        // Three passes over the single wide accumulator, 32 columns at a time (the two warps of a TMEM lane quarter
        // split the chunks): row max -> exchange -> sum of exponentials -> exchange -> normalised probabilities.
        // Nothing but one chunk lives in registers, the logits stay in TMEM until the last pass has read them.
        const int ns = p.sm_ns;
        const int used = (ns + 31) >> 5;                       // chunks that hold keys (<= 16)
        const int n_first = (used + 1) >> 1;
        const int my_c0 = half ? n_first : 0;
        const int my_n = half ? used - n_first : n_first;
        const uint32_t trow0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* xs_mine = s_xchg + (warp - 2) * 64;
        const float* xs_peer = s_xchg + ((warp - 2) ^ 4) * 64;
        const float sc = p.alpha * 1.4426950408889634f;
        uint32_t v[32];
        float mx = -INFINITY;
#pragma unroll 1
        for (int ci = 0; ci < my_n; ++ci) {
          const int c = my_c0 + ci;
          tmem_ld32(trow0 + c * 32, v);
          tmem_ld_wait();
          const int nv = ns - c * 32;
          if (nv >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
        }
        xs_mine[lane] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        mx = fmaxf(mx, xs_peer[lane]);
        const float mb = mx * sc;
        float sum = 0.0f;
#pragma unroll 1
        for (int ci = 0; ci < my_n; ++ci) {
          const int c = my_c0 + ci;
          tmem_ld32(trow0 + c * 32, v);
          tmem_ld_wait();
          const int nv = ns - c * 32;
          float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            const float t = fmaf(__uint_as_float(v[j]), sc, -mb);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
            if (j >= nv) e = 0.0f;
            part[j & 3] += e;
          }
          sum += (part[0] + part[1]) + (part[2] + part[3]);
        }
        xs_mine[32 + lane] = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        sum += xs_peer[32 + lane];
        const float inv = 1.0f / sum;
        const long long seg_col = static_cast<long long>(co_t) * p.sm_pitch;
#pragma unroll 1
        for (int ci = 0; ci < my_n; ++ci) {
          const int c = my_c0 + ci;
          tmem_ld32(trow0 + c * 32, v);
          tmem_ld_wait();
          const int nv = ns - c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e;
            const float t = fmaf(__uint_as_float(v[j]), sc, -mb);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
            v[j] = __float_as_uint(j < nv ? e * inv : 0.0f);
          }
          uint4* trow = reinterpret_cast<uint4*>(tb + lane * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) trow[g ^ (lane & 7)] = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          __syncwarp();
          const int col = c * 32 + cg;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + sub;
            const float4* tr = reinterpret_cast<const float4*>(tb + r * 32);
            const float4 a = tr[(2 * (lane & 3)) ^ (r & 7)];
            const float4 b = tr[(2 * (lane & 3) + 1) ^ (r & 7)];
            if (row_ok[i] && col < p.sm_pitch) {
              const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - __uint_as_float(ph[e] << 16),
                                                                f[2 * e + 1] - __uint_as_float(ph[e] & 0xFFFF0000u));
                pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              *reinterpret_cast<uint4*>(p.out_hi + o_off[i] + seg_col + col) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (NSPLIT == 2)
                *reinterpret_cast<uint4*>(p.out_lo + o_off[i] + seg_col + col) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
          }
          __syncwarp();
        }
        // every tcgen05.ld of this warp has completed: hand the (only) accumulator back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == kAccBufs) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      if constexpr (SOFTMAX && !SOFTMAXW) {
        static_assert(!SOFTMAX || Cfg::kXchgBytes > 0 || 2 * BLOCK_N * 4 >= 8 * 64 * 4, "no room for the softmax exchange");
        // The segment's 32-column chunks are split between the two warps of the TMEM lane quarter.  Each warp
        // pulls its chunks into registers with one tcgen05.ld burst and hands the accumulator back to the MMA
        // warp at once; row max and row sum are combined with the partner through shared memory (64-thread
        // named barrier per quarter); one ex2.approx per element, scale folded into the exponent.
        const int ns = p.sm_ns;
        const int used = (ns + 31) >> 5;                       // chunks that hold keys (<= BLOCK_N / 32)
        const int n_first = (used + 1) >> 1;
        const int my_c0 = half ? n_first : 0;
        const int my_n = half ? used - n_first : n_first;
        constexpr int kMaxC = (kChunks + 1) / 2;
        const uint32_t trow0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        uint32_t v[kMaxC][32];
#pragma unroll
        for (int ci = 0; ci < kMaxC; ++ci)
          if (ci < my_n) tmem_ld32(trow0 + (my_c0 + ci) * 32, v[ci]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);           // the logits live in registers from here on
        if (++acc == kAccBufs) {
          acc = 0;
          acc_phase ^= 1;
        }
        float* xs_mine = s_xchg + (warp - 2) * 64;
        const float* xs_peer = s_xchg + ((warp - 2) ^ 4) * 64;
        float mx = -INFINITY;
        // only the segment's last chunk can be partial: full chunks run without per-element predicates
#pragma unroll
        for (int ci = 0; ci < kMaxC; ++ci)
          if (ci < my_n) {
            const int nv = ns - (my_c0 + ci) * 32;   // valid columns of this chunk (>= 1)
            if (nv >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[ci][j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) mx = fmaxf(mx, __uint_as_float(v[ci][j]));
            }
          }
        xs_mine[lane] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        mx = fmaxf(mx, xs_peer[lane]);
        const float sc = p.alpha * 1.4426950408889634f;
        const float mb = mx * sc;
        float sum = 0.0f;
#pragma unroll
        for (int ci = 0; ci < kMaxC; ++ci)
          if (ci < my_n) {
            const int nv = ns - (my_c0 + ci) * 32;
            float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // four independent sum chains
            if (nv >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float e;
                const float t = fmaf(__uint_as_float(v[ci][j]), sc, -mb);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
                part[j & 3] += e;
                v[ci][j] = __float_as_uint(e);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float e;
                const float t = fmaf(__uint_as_float(v[ci][j]), sc, -mb);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
                if (j >= nv) e = 0.0f;
                part[j & 3] += e;
                v[ci][j] = __float_as_uint(e);
              }
            }
            sum += (part[0] + part[1]) + (part[2] + part[3]);
          }
        xs_mine[32 + lane] = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        sum += xs_peer[32 + lane];
        const float inv = 1.0f / sum;
        const long long seg_col = static_cast<long long>(co_t) * p.sm_pitch;
#pragma unroll
        for (int ci = 0; ci < kMaxC; ++ci) {
          if (ci >= my_n) break;
          const int c = my_c0 + ci;
          float4* trow = reinterpret_cast<float4*>(tb + lane * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            trow[g ^ (lane & 7)] = make_float4(__uint_as_float(v[ci][g * 4 + 0]) * inv, __uint_as_float(v[ci][g * 4 + 1]) * inv,
                                               __uint_as_float(v[ci][g * 4 + 2]) * inv, __uint_as_float(v[ci][g * 4 + 3]) * inv);
          __syncwarp();
          const int col = c * 32 + cg;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + sub;
            const float4* tr = reinterpret_cast<const float4*>(tb + r * 32);
            const float4 a = tr[(2 * (lane & 3)) ^ (r & 7)];
            const float4 b = tr[(2 * (lane & 3) + 1) ^ (r & 7)];
            if (row_ok[i] && col < p.sm_pitch) {
              const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - __uint_as_float(ph[e] << 16),
                                                                f[2 * e + 1] - __uint_as_float(ph[e] & 0xFFFF0000u));
                pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              *reinterpret_cast<uint4*>(p.out_hi + o_off[i] + seg_col + col) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (NSPLIT == 2)
                *reinterpret_cast<uint4*>(p.out_lo + o_off[i] + seg_col + col) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
          }
          __syncwarp();
        }
        continue;
      }
#pragma unroll
      for (int ci = 0; ci < kCpw; ++ci) {
        const int c = c_begin + ci;
        if (c >= kChunks) break;
        // the last N-tile may be partial; FAST launches have n_out % 32 == 0, so a chunk is either whole or absent
        // (the generic path predicates per element)
        if (FAST && co0 + c * 32 >= p.n_out) break;
        // ---- phase A: accumulator row -> scale/bias -> swizzled smem tile
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * BLOCK_N + c * 32);
        tmem_ld32(taddr, v);
        uint4 ch[4], cl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ch[i] = rh[i];
          cl[i] = rl[i];
        }
        if (ci + 1 < kCpw && c + 1 < kChunks && co0 + (c + 1) * 32 < p.n_out) {
          prefetch(c + 1);
        } else {
          prefetch_next();
        }
        tmem_ld_wait();
        for (int cc = sk_first; cc < sk_last; ++cc) {   // stream-K: add the partial sums of the tile's head
          const float* src = p.sk_partials + static_cast<long long>(cc) * (128 * BLOCK_N) + (c * 4 + q) * 1024 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(src + j * 32));
        }
        {
          // raw accumulators into the swizzled tile; scale / bias are applied in phase B, where a lane needs
          // only the 8 values of its own channel group
          uint4* trow = reinterpret_cast<uint4*>(tb + lane * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) trow[g ^ (lane & 7)] = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
        __syncwarp();
        // ---- phase B: 8 consecutive channels of 4 rows per lane, row-contiguous global accesses
        const int col = co0 + c * 32 + cg;
        const bool col_full = (col + 8 <= p.n_out);
        const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + c * 32 + cg);
        const float4 sc1 = *reinterpret_cast<const float4*>(s_scale + c * 32 + cg + 4);
        const float4 bi0 = *reinterpret_cast<const float4*>(s_bias + c * 32 + cg);
        const float4 bi1 = *reinterpret_cast<const float4*>(s_bias + c * 32 + cg + 4);
        if constexpr (FAST) {
          const float floor_v = p.relu ? 0.0f : (F16IO ? -65504.0f : -INFINITY);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + sub;
            const float4* trow = reinterpret_cast<const float4*>(tb + r * 32);
            const float4 a = trow[(2 * (lane & 3)) ^ (r & 7)];
            const float4 b = trow[(2 * (lane & 3) + 1) ^ (r & 7)];
            float f[8] = {a.x * sc0.x + bi0.x, a.y * sc0.y + bi0.y, a.z * sc0.z + bi0.z, a.w * sc0.w + bi0.w,
                          b.x * sc1.x + bi1.x, b.y * sc1.y + bi1.y, b.z * sc1.z + bi1.z, b.w * sc1.w + bi1.w};
            if constexpr (F16IO) {
              if (res_pair) {
                const uint32_t wh[4] = {ch[i].x, ch[i].y, ch[i].z, ch[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&wh[e]));
                  f[2 * e] += rf.x;
                  f[2 * e + 1] += rf.y;
                }
              }
              uint32_t ph[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(fminf(fmaxf(f[2 * e], floor_v), 65504.0f),
                                                     fminf(fmaxf(f[2 * e + 1], floor_v), 65504.0f));
                ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              if (row_ok[i]) *reinterpret_cast<uint4*>(ohp[i] + ci * 32) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            } else {
            if (res_pair) {
              const uint32_t wh[4] = {ch[i].x, ch[i].y, ch[i].z, ch[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[2 * e] += __uint_as_float(wh[e] << 16);
                f[2 * e + 1] += __uint_as_float(wh[e] & 0xFFFF0000u);
              }
              if (NSPLIT == 2) {
                const uint32_t wl[4] = {cl[i].x, cl[i].y, cl[i].z, cl[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  f[2 * e] += __uint_as_float(wl[e] << 16);
                  f[2 * e + 1] += __uint_as_float(wl[e] & 0xFFFF0000u);
                }
              }
            }
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float f0 = fmaxf(f[2 * e], floor_v), f1 = fmaxf(f[2 * e + 1], floor_v);
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
              ph[e] = *reinterpret_cast<const uint32_t*>(&h2);
              if (NSPLIT == 2) {
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(f0 - __uint_as_float(ph[e] << 16),
                                                                f1 - __uint_as_float(ph[e] & 0xFFFF0000u));
                pl[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
            }
            if (row_ok[i]) {
              *reinterpret_cast<uint4*>(ohp[i] + ci * 32) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (NSPLIT == 2) *reinterpret_cast<uint4*>(ohp[i] + d_out + ci * 32) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            }  // !F16IO
          }
        } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + sub;
          const float4* trow = reinterpret_cast<const float4*>(tb + r * 32);
          const float4 a = trow[(2 * (lane & 3)) ^ (r & 7)];
          const float4 b = trow[(2 * (lane & 3) + 1) ^ (r & 7)];
          if (!row_ok[i] || col >= p.n_out) continue;
          float f[8] = {a.x * sc0.x + bi0.x, a.y * sc0.y + bi0.y, a.z * sc0.z + bi0.z, a.w * sc0.w + bi0.w,
                          b.x * sc1.x + bi1.x, b.y * sc1.y + bi1.y, b.z * sc1.z + bi1.z, b.w * sc1.w + bi1.w};
          if (p.res_f32 != nullptr) {
            const float* rp = p.res_f32 + r_off[i] + col;
            if (col_full && out_vec) {
              const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
              const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp) + 1);
              f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
              f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (col + k < p.n_out) f[k] += __ldg(rp + k);
            }
          } else if (res_pair) {
            if (res_vec && col_full) {
              const uint32_t wh[4] = {ch[i].x, ch[i].y, ch[i].z, ch[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[2 * e] += __uint_as_float(wh[e] << 16);
                f[2 * e + 1] += __uint_as_float(wh[e] & 0xFFFF0000u);
              }
              if (res_lo) {
                const uint32_t wl[4] = {cl[i].x, cl[i].y, cl[i].z, cl[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  f[2 * e] += __uint_as_float(wl[e] << 16);
                  f[2 * e + 1] += __uint_as_float(wl[e] & 0xFFFF0000u);
                }
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (col + k < p.n_out) {
                  f[k] += __bfloat162float(p.res_hi[r_off[i] + col + k]);
                  if (res_lo) f[k] += __bfloat162float(p.res_lo[r_off[i] + col + k]);
                }
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
          }
          if (p.out_f32 != nullptr) {
            float* op = p.out_f32 + o_off[i] + col;
            if (col_full && out_vec) {
              reinterpret_cast<float4*>(op)[0] = make_float4(f[0], f[1], f[2], f[3]);
              reinterpret_cast<float4*>(op)[1] = make_float4(f[4], f[5], f[6], f[7]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (col + k < p.n_out) op[k] = f[k];
            }
          }
          if (p.out_hi != nullptr) {
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(f[2 * e], h0, l0);
              split_bf16(f[2 * e + 1], h1, l1);
              ph[e] = pack_bf16x2(h0, h1);
              pl[e] = pack_bf16x2(l0, l1);
            }
            __nv_bfloat16* oh = p.out_hi + o_off[i] + col;
            __nv_bfloat16* ol = (p.out_lo != nullptr) ? p.out_lo + o_off[i] + col : nullptr;
            if (col_full && out_vec) {
              *reinterpret_cast<uint4*>(oh) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              if (ol != nullptr) *reinterpret_cast<uint4*>(ol) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (col + k < p.n_out) {
                  const uint32_t wh = ph[k >> 1], wl = pl[k >> 1];
                  oh[k] = __ushort_as_bfloat16(static_cast<unsigned short>((k & 1) ? (wh >> 16) : (wh & 0xFFFF)));
                  if (ol != nullptr)
                    ol[k] = __ushort_as_bfloat16(static_cast<unsigned short>((k & 1) ? (wl >> 16) : (wl & 0xFFFF)));
                }
              }
            }
          }
        }
        }  // !FAST
        __syncwarp();   // the tile is rewritten by the next chunk's phase A
      }
      if (sk_first < sk_last) {
        // every epilogue warp has folded the partial sums in: hand the contributors' flags back (zero), so the
        // next launch -- or the next replay of a captured graph -- starts from a clean workspace
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < sk_last - sk_first) atomicExch(p.sk_flags + sk_first + et, 0);
      }
      // all tcgen05.ld of this warp have completed (wait::ld above): release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == kAccBufs) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CM > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into / signal this CTA
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace dana
