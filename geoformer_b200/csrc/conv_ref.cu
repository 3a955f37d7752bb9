// fp32 FFMA reference of the backbone (resnet_fpn.py:15-40, 58-118) for the accurate mode of this library: the
// golden-match parity tests need a backbone whose features agree with the fp32 reference to ~1e-5, which the fp16
// tcgen05 path (conv_tc.cu) cannot give.  Same contract as gf_conv_f16 / gf_upsample_add_f16, NHWC fp32 throughout,
// any ksize in {1, 3, 7} (pad = ksize / 2), stride 1 or 2, any channel counts (the 1-channel stem included).
// Not a product kernel: a plain smem-tiled SGEMM-style implicit GEMM (64 pixels x 64 channels per CTA, 4 x 4 per thread).
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

constexpr int kPx = 64, kCo = 64, kKc = 16;

__global__ void __launch_bounds__(256)
conv_ref_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                const float* __restrict__ residual, float* __restrict__ y, int batch, int h, int w, int ho, int wo,
                int cin, int cout, int ksize, int stride, int act) {
  __shared__ float As[kKc][kPx + 4];
  __shared__ float Bs[kKc][kCo + 4];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int64_t p0 = (int64_t)blockIdx.x * kPx, npix = (int64_t)batch * ho * wo;
  const int co0 = blockIdx.y * kCo, pad = ksize / 2;
  // A loader: thread -> (pixel t / 4, channel quad t % 4)
  const int lp = t / 4, lq = (t % 4) * 4;
  const int64_t pl = p0 + lp;
  int lb = 0, loy = 0, lox = 0;
  const bool pvalid = pl < npix;
  if (pvalid) { lox = (int)(pl % wo); loy = (int)((pl / wo) % ho); lb = (int)(pl / ((int64_t)wo * ho)); }
  // B loader: thread -> (k row t / 16, channel quad t % 16)
  const int bk = t / 16, bj = (t % 16) * 4;
  float acc[4][4] = {};
  for (int tap = 0; tap < ksize * ksize; ++tap) {
    const int iy = loy * stride + tap / ksize - pad, ix = lox * stride + tap % ksize - pad;
    const bool inb = pvalid && iy >= 0 && iy < h && ix >= 0 && ix < w;
    const float* xrow = x + (((int64_t)lb * h + iy) * w + ix) * cin;
    for (int c0 = 0; c0 < cin; c0 += kKc) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + lq + q;
        As[lq + q][lp] = (inb && c < cin) ? __ldg(xrow + c) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + bk, co = co0 + bj + q;
        Bs[bk][bj + q] = (c < cin && co < cout) ? __ldg(wt + ((int64_t)tap * cin + c) * cout + co) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kKc; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t p = p0 + ty * 4 + i;
    if (p >= npix) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= cout) continue;
      float v = acc[i][j] + (bias ? bias[co] : 0.f);
      if (residual) v += residual[p * cout + co];
      if (act == 1) v = fmaxf(v, 0.f);
      else if (act == 2) v = v > 0.f ? v : 0.01f * v;
      y[p * cout + co] = v;
    }
  }
}

// out = lateral + bilinear_upsample(src -> (h, w), align_corners=True), NHWC fp32 (resnet_fpn.py:108-115)
__global__ void upsample_add_ref_kernel(const float* __restrict__ lateral, const float* __restrict__ src,
                                        float* __restrict__ out, int b, int h, int w, int hs, int ws, int c, float ry,
                                        float rx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)b * h * w * c) return;
  const int ch = (int)(idx % c);
  int64_t r = idx / c;
  const int xo = (int)(r % w); r /= w;
  const int yo = (int)(r % h);
  const int n = (int)(r / h);
  const float fy = ry * yo, fx = rx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float* s = src + (int64_t)n * hs * ws * c + ch;
  const float v00 = s[((int64_t)y0 * ws + x0) * c], v01 = s[((int64_t)y0 * ws + x1) * c];
  const float v10 = s[((int64_t)y1 * ws + x0) * c], v11 = s[((int64_t)y1 * ws + x1) * c];
  // same association as ATen's upsample_bilinear2d: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
  out[idx] = lateral[idx] + ((1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11));
}

}  // namespace gf

extern "C" int gf_conv_ref(const float* x, const float* wt, const float* bias, const float* residual, float* y, int batch,
                           int h, int w, int cin, int cout, int ksize, int stride, int act, gf_stream_t stream) {
  if (batch <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || (ksize != 1 && ksize != 3 && ksize != 7) ||
      (stride != 1 && stride != 2) || act < 0 || act > 2)
    return gf_set_error(GF_ERR_ARG, "gf_conv_ref: bad shape");
  const int pad = ksize / 2;
  const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (w + 2 * pad - ksize) / stride + 1;
  const int64_t npix = (int64_t)batch * ho * wo;
  dim3 grid(gf_cdiv(npix, gf::kPx), gf_cdiv(cout, gf::kCo));
  gf::conv_ref_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, wt, bias, residual, y, batch, h, w, ho, wo, cin, cout,
                                                              ksize, stride, act);
  gf::g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_upsample_add_ref(const float* lateral, const float* src, float* out, int batch, int h, int w, int hs,
                                   int ws, int c, gf_stream_t stream) {
  if (batch <= 0 || h <= 1 || w <= 1 || hs <= 0 || ws <= 0 || c <= 0)
    return gf_set_error(GF_ERR_ARG, "gf_upsample_add_ref: bad shape");
  const int64_t total = (int64_t)batch * h * w * c;
  const float ry = (float)(hs - 1) / (float)(h - 1), rx = (float)(ws - 1) / (float)(w - 1);
  gf::upsample_add_ref_kernel<<<gf_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(lateral, src, out, batch, h, w, hs, ws,
                                                                                    c, ry, rx);
  gf::g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
