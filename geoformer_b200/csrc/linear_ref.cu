// Plain fp32 FFMA implementation of the gf_linear_* contract.  It exists so that (a) the tcgen05 path
// can be validated kernel-against-kernel on the GPU and (b) parity runs can use exact-fp32 projections
// ("accurate mode", SURVEY.md hard part 8).  64x64 tile, 256 threads, 4x4 outputs per thread.
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

__global__ void __launch_bounds__(256)
linear_ref_kernel(const float* __restrict__ A, const float* __restrict__ A2, const float* __restrict__ W,
                  float* __restrict__ Y, int64_t M, int N, int K1, int K2, int epi, int act_cols,
                  const float* __restrict__ bias, const float* __restrict__ rowbias, int rowbias_group,
                  const float* __restrict__ residual, const int* __restrict__ m_dev) {
  __shared__ float As[16][64 + 4], Ws[16][64 + 4];
  int64_t m_live = M;
  if (m_dev) { int v = *m_dev; m_live = v < 0 ? 0 : (v < M ? v : M); }
  const int64_t row0 = (int64_t)blockIdx.x * 64;
  if (row0 >= m_live) return;
  const int col0 = blockIdx.y * 64;
  const int K = K1 + K2;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = tid; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      const int64_t gr = row0 + r;
      const int gk = k0 + k;
      float av = 0.f;
      if (gr < m_live && gk < K) av = gk < K1 ? A[gr * K1 + gk] : A2[gr * K2 + (gk - K1)];
      As[k][r] = av;
      const int gc = col0 + r;
      Ws[k][r] = (gc < N && gk < K) ? W[(int64_t)gc * K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][4 * tx]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + 4 * ty + i;
    if (r >= m_live) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + 4 * tx + j;
      if (c >= N) continue;
      float x = acc[i][j];
      if (bias) x += bias[c];
      if (rowbias) x += rowbias[(r / rowbias_group) * N + c];
      if (epi & GF_EPI_RELU) x = fmaxf(x, 0.f);
      if (epi & GF_EPI_TANH) x = tanhf(x);
      if ((epi & GF_EPI_ELU1) && c < act_cols) x = elu1(x);
      if (residual && !(epi & GF_EPI_LN)) x += residual[r * N + c];
      Y[r * N + c] = x;
    }
  }
}

// in-place row LayerNorm (+ residual); one warp per row
__global__ void layernorm_rows_kernel(float* __restrict__ Y, int64_t M, int N, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ residual,
                                      const int* __restrict__ m_dev) {
  int64_t m_live = M;
  if (m_dev) { int v = *m_dev; m_live = v < 0 ? 0 : (v < M ? v : M); }
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_live) return;
  float* p = Y + row * N;
  float s = 0.f;
  for (int c = lane; c < N; c += 32) s += p[c];
  const float mean = warp_sum(s) / (float)N;
  float q = 0.f;
  for (int c = lane; c < N; c += 32) { const float d = p[c] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)N + 1e-5f);
  for (int c = lane; c < N; c += 32) {
    float x = (p[c] - mean) * rstd * gamma[c] + beta[c];
    if (residual) x += residual[row * N + c];
    p[c] = x;
  }
}

}  // namespace gf

using namespace gf;

extern "C" int gf_linear_ref(const float* A, const float* A2, const float* W, float* Y, int64_t M, int N, int K1,
                             int K2, int epi, int act_cols, const float* bias, const float* rowbias,
                             int rowbias_group, const float* gamma, const float* beta, const float* residual,
                             const int* m_dev, gf_stream_t stream) {
  if (M < 0 || N <= 0 || K1 <= 0 || K2 < 0) return gf_set_error(GF_ERR_ARG, "gf_linear_ref: bad shape");
  if ((epi & GF_EPI_LN) && (!gamma || !beta)) return gf_set_error(GF_ERR_ARG, "gf_linear_ref: LN needs gamma/beta");
  if (M == 0) return GF_OK;
  dim3 grid(gf_cdiv(M, 64), gf_cdiv(N, 64));
  linear_ref_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, A2, W, Y, M, N, K1, K2, epi, act_cols, bias, rowbias,
                                                            rowbias_group > 0 ? rowbias_group : 1, residual, m_dev);
  g_launches++;
  if (epi & GF_EPI_LN) {
    layernorm_rows_kernel<<<gf_cdiv(M, 8), 256, 0, (cudaStream_t)stream>>>(Y, M, N, gamma, beta, residual, m_dev);
    g_launches++;
  }
  GF_CHECK_LAUNCH();
  return GF_OK;
}
