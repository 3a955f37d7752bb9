// Optional padding masks of the coarse level (data['mask0'] / data['mask1'], MegaDepth-style zero-padded batches).
// Reference: LinearAttention multiplies the feature-mapped Q / K and the values by the 0/1 token masks
// (model/loftr_src/loftr/loftr_module/linear_attention.py:37-43) and CoarseMatching fills the logits of every
// (padded row | padded column) entry with -1e9 before the dual softmax
// (model/loftr_src/loftr/utils/coarse_matching.py:120-124).  Both are pure HBM passes over buffers the projection /
// similarity kernels have just written; they only run when the caller passes masks (the inference wrappers never do).
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

// One warp per row: rows whose mask byte is 0 get the 32-bit words [w0, w0 + words) of the row cleared.  A multiply by a
// 0/1 mask is a clear (the bit pattern of +0 is the same for fp16 and fp32), so one kernel serves both storage types.
__global__ void __launch_bounds__(256)
mask_rows_kernel(uint32_t* __restrict__ buf, int64_t rows, int64_t ld_words, int64_t w0, int words,
                 const uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    if (mask[r] != 0) continue;                    // warp-uniform
    uint32_t* p = buf + r * ld_words + w0;
    for (int i = lane; i < words; i += 32) p[i] = 0u;
  }
}

// sim[b, i, j] = fill unless mask0[b, i] && mask1[b, j].  CTA = 256 consecutive columns of one row (coalesced);
// rows of padded tokens are written without reading the column mask.
__global__ void __launch_bounds__(256)
mask_fill_sim_kernel(float* __restrict__ sim, int l, int s, const uint8_t* __restrict__ mask0,
                     const uint8_t* __restrict__ mask1, float fill) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= s) return;
  const bool row_ok = mask0[(int64_t)b * l + i] != 0;
  if (row_ok && mask1[(int64_t)b * s + j] != 0) return;
  sim[((int64_t)b * l + i) * s + j] = fill;
}

}  // namespace gf

using namespace gf;
#define STREAM ((cudaStream_t)stream)

extern "C" int gf_mask_rows(void* buf, int elem_bytes, int64_t rows, int64_t ld, int64_t col0, int64_t cols,
                            const uint8_t* mask, gf_stream_t stream) {
  if ((elem_bytes != 2 && elem_bytes != 4) || rows < 0 || ld <= 0 || col0 < 0 || cols <= 0 || col0 + cols > ld)
    return gf_set_error(GF_ERR_ARG, "gf_mask_rows: bad shape");
  if (((ld * elem_bytes) | (col0 * elem_bytes) | (cols * elem_bytes)) & 3 || ((uintptr_t)buf & 3))
    return gf_set_error(GF_ERR_ARG, "gf_mask_rows: row stride, column offset and width must be multiples of 4 bytes");
  if (cols * elem_bytes / 4 > 0x7fffffff) return gf_set_error(GF_ERR_ARG, "gf_mask_rows: row too wide");
  if (rows == 0) return GF_OK;
  const int64_t blocks = (rows + 7) / 8;                       // 8 warps (rows) per CTA
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);      // grid-stride beyond 8 CTAs per SM
  mask_rows_kernel<<<grid, 256, 0, STREAM>>>((uint32_t*)buf, rows, ld * elem_bytes / 4, col0 * elem_bytes / 4,
                                             (int)(cols * elem_bytes / 4), mask);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_mask_fill_sim(float* sim, int n, int l, int s, const uint8_t* mask0, const uint8_t* mask1, float fill,
                                gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0 || n > 65535 || l > 65535) return gf_set_error(GF_ERR_ARG, "gf_mask_fill_sim: bad shape");
  mask_fill_sim_kernel<<<dim3(gf_cdiv(s, 256), l, n), 256, 0, STREAM>>>(sim, l, s, mask0, mask1, fill);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
