// Host-side helpers shared by the C-ABI entry points: error bookkeeping and
// TMA tensor-map encoding through the driver entry point (no link-time libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dana_b200.h"

namespace dana {

static thread_local int t_last_cuda_error = 0;

inline int cuda_fail(cudaError_t e) {
  t_last_cuda_error = static_cast<int>(e);
  return DANA_ECUDA;
}
#define DANA_CUDA_CHECK(expr)                       \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) return dana::cuda_fail(_e); \
  } while (0)
#define DANA_LAUNCH_CHECK() DANA_CUDA_CHECK(cudaGetLastError())

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// bf16 tensor map, 128-byte (or 64-byte) swizzle matching the box's inner extent, zero fill out of bounds.
// dims[0] is the contiguous dimension; strides_bytes has rank-1 entries (dims 1..rank-1).
inline int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, bool swizzle64 = false) {
  EncodeTiledFn fn = get_encode_tiled();
  if (fn == nullptr) return DANA_ECUDA;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[5];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), d, s, b,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    t_last_cuda_error = 100000 + static_cast<int>(r);
    return DANA_ECUDA;
  }
  return DANA_OK;
}

// Kernel attributes (opt-in dynamic shared memory) are per device: the "already configured" flags of the launchers are
// indexed by the current device, so a single process driving several GPUs (the per-GPU threads of nn.DataParallel,
// train.py:104-105) configures every one of them.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

inline int sm_count() {
  static int n_dev[kMaxDevices] = {};
  int& n = n_dev[current_device()];
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace dana
