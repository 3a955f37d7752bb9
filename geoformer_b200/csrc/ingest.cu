// Image ingest: the step right before the hot path.
// eval_tool/immatch/utils/data_io.py:48-62 does cv2.imread(GRAY) -> cv2.resize(im, (wt, ht)) -> to_tensor (/255).
// This kernel reproduces OpenCV's 8-bit INTER_LINEAR resize BIT-EXACTLY on the GPU and fuses the normalisation:
//   fx = (float)((dx + 0.5) * (wo / wt) - 0.5); sx = floor(fx); fx -= sx; clamp (sx < 0 -> 0, fx = 0; sx >= wo-1 -> wo-1, fx = 0)
//   horizontal taps in 11-bit fixed point: a0 = round((1 - fx) * 2048), a1 = round(fx * 2048)   (same for rows: b0, b1)
//   out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2        with r = S[sx] * a0 + S[sx + 1] * a1
// so only the decoded uint8 image (4x smaller than fp32) crosses PCIe and the resized tensor equals the reference's.
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

__device__ __forceinline__ void linear_tap(int d, double scale, int n_src, int& s0, int& s1, int& a0, int& a1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
  a0 = __float2int_rn((1.f - f) * 2048.f);
  a1 = __float2int_rn(f * 2048.f);
  s0 = s;
  s1 = min(s + 1, n_src - 1);
}

__global__ void resize_gray_u8_kernel(const uint8_t* __restrict__ src, int ho, int wo, float* __restrict__ dst, int ht,
                                      int wt, double sx_scale, double sy_scale) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= wt || y >= ht) return;
  int x0, x1, a0, a1, y0, y1, b0, b1;
  linear_tap(x, sx_scale, wo, x0, x1, a0, a1);
  linear_tap(y, sy_scale, ho, y0, y1, b0, b1);
  const int r0 = (int)src[(int64_t)y0 * wo + x0] * a0 + (int)src[(int64_t)y0 * wo + x1] * a1;
  const int r1 = (int)src[(int64_t)y1 * wo + x0] * a0 + (int)src[(int64_t)y1 * wo + x1] * a1;
  int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  v = max(0, min(255, v));
  dst[(int64_t)y * wt + x] = __fdiv_rn((float)v, 255.f);      // torchvision to_tensor: uint8 -> float / 255
}

}  // namespace gf

using namespace gf;

// src: device uint8 [ho, wo] (decoded grayscale image); dst: device fp32 [ht, wt] in [0, 1]
extern "C" int gf_resize_gray_u8(const void* src, int ho, int wo, float* dst, int ht, int wt, gf_stream_t stream) {
  if (ho <= 0 || wo <= 0 || ht <= 0 || wt <= 0) return gf_set_error(GF_ERR_ARG, "gf_resize_gray_u8: bad shape");
  dim3 block(32, 8), grid(gf_cdiv(wt, 32), gf_cdiv(ht, 8));
  resize_gray_u8_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint8_t*)src, ho, wo, dst, ht, wt,
                                                                 (double)wo / wt, (double)ho / ht);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
