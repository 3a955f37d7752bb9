// Shared helpers for the sm_100a kernels of libgeoformer_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

#include "../../include/geoformer_b200.h"

#define GF_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return gf_set_error(GF_ERR_LAUNCH, cudaGetErrorString(e__)); \
  } while (0)

int gf_set_error(int code, const char* msg);

static inline int gf_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (function, DEVICE) pair, so it is tracked per
// device of the calling thread (one bit per device ordinal), not once per process.  Use at launch sites:
//   GF_SMEM_OPTIN(kernel, bytes);     // returns GF_ERR_LAUNCH from the enclosing function on failure
#include <atomic>
#define GF_SMEM_OPTIN(kernel, bytes)                                                                         \
  do {                                                                                                       \
    static std::atomic<uint64_t> done__{0};                                                                  \
    int dev__ = 0;                                                                                           \
    cudaGetDevice(&dev__);                                                                                   \
    const uint64_t bit__ = 1ull << (dev__ & 63);                                                             \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                                                 \
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess) \
        return gf_set_error(GF_ERR_LAUNCH, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");        \
      done__.fetch_or(bit__, std::memory_order_release);                                                     \
    }                                                                                                        \
  } while (0)

namespace gf {

// TMA descriptor helpers (gemm_tc.cu).  K-major operand map: dims {k, rows, batches}, box {128 B, box_rows, 1},
// 128B swizzle, OOB -> 0.  Output map: dims {n, rows, batches} fp32, box {32, 32, 1}, 128B swizzle.
int make_tmap(CUtensorMap* m, const void* base, int esize, int64_t k, int64_t rows, int64_t batches,
              int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows);
int make_out_tmap(CUtensorMap* m, float* base, int64_t n, int64_t rows, int64_t batches, int64_t ld,
                  int64_t batch_stride);
int num_sms();

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float elu1(float x) {   // elu(x) + 1  (linear_attention.py:11-12)
  return x > 0.f ? x + 1.f : expf(x);              // expm1(x)+1 == exp(x) up to 1 ulp
}

}  // namespace gf
