// Linear attention (model/loftr_src/loftr/loftr_module/linear_attention.py:33-49) and small
// element-wise helpers.  Q and K arrive already feature-mapped (elu+1 fused into the projection epilogue).
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

constexpr int kChunk = 256;   // source tokens per partial reduction block

// partial[(n,h,chunk)][D*D + D]: KV block then Ksum
__global__ void linattn_partial_kernel(const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
                                       int s, int heads, int dim, float inv_s, float* __restrict__ partial) {
  extern __shared__ float sh[];
  float* Ks = sh;                    // [64][dim]
  float* Vs = sh + 64 * dim;         // [64][dim]
  const int chunk = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int nchunks = gridDim.x;
  const int entries = dim * dim;
  const int s0 = chunk * kChunk, s1 = min(s, s0 + kChunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // up to 4 entries per thread (dim <= 32 with 256 threads)
  float ksum = 0.f;
  for (int t0 = s0; t0 < s1; t0 += 64) {
    const int cnt = min(64, s1 - t0);
    for (int e = threadIdx.x; e < 64 * dim; e += blockDim.x) {
      const int r = e / dim, d = e - r * dim;
      float kv = 0.f, vv = 0.f;
      if (r < cnt) {
        const int64_t tok = (int64_t)n * s + t0 + r;
        kv = K[tok * ldk + h * dim + d];
        vv = V[tok * ldv + h * dim + d] * inv_s;   // values / v_length (linear_attention.py:46)
      }
      Ks[e] = kv; Vs[e] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + q * blockDim.x;
      if (e < entries) {
        const int d = e / dim, v = e - d * dim;
        float a = acc[q];
        for (int r = 0; r < cnt; ++r) a = fmaf(Ks[r * dim + d], Vs[r * dim + v], a);
        acc[q] = a;
      }
    }
    if (threadIdx.x < dim) {
      float a = ksum;
      for (int r = 0; r < cnt; ++r) a += Ks[r * dim + threadIdx.x];
      ksum = a;
    }
    __syncthreads();
  }
  float* out = partial + ((int64_t)(n * heads + h) * nchunks + chunk) * (entries + dim);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = threadIdx.x + q * blockDim.x;
    if (e < entries) out[e] = acc[q];
  }
  if (threadIdx.x < dim) out[entries + threadIdx.x] = ksum;
}

__global__ void linattn_finalize_kernel(const float* __restrict__ partial, int nchunks, int dim, float* __restrict__ KV,
                                        float* __restrict__ Ksum) {
  const int nh = blockIdx.x;
  const int entries = dim * dim;
  for (int e = threadIdx.x; e < entries + dim; e += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < nchunks; ++c) a += partial[((int64_t)nh * nchunks + c) * (entries + dim) + e];
    if (e < entries) KV[(int64_t)nh * entries + e] = a; else Ksum[(int64_t)nh * dim + (e - entries)] = a;
  }
}

// out[n,l,h,v] = (sum_d Q[l,h,d] KV[h,d,v]) * (1 / (Q[l,h,:].Ksum[h,:] + eps)) * S      blockDim = heads*dim
__global__ void linattn_apply_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ KV,
                                     const float* __restrict__ Ksum, float* __restrict__ out, int l, int heads, int dim,
                                     float s_len) {
  extern __shared__ float sh[];
  const int c = heads * dim;
  float* kv = sh;                       // [heads][dim][dim]
  float* ks = kv + heads * dim * dim;   // [c]
  float* qs = ks + c;                   // [32][c]
  const int n = blockIdx.y;
  const int l0 = blockIdx.x * 32;
  const int t = threadIdx.x;
  for (int e = t; e < heads * dim * dim; e += blockDim.x) kv[e] = KV[(int64_t)n * heads * dim * dim + e];
  ks[t] = Ksum[(int64_t)n * c + t];
  const int cnt = min(32, l - l0);
  for (int r = 0; r < cnt; ++r) qs[r * c + t] = Q[((int64_t)n * l + l0 + r) * ldq + t];
  __syncthreads();
  const int h = t / dim, v = t - h * dim;
  const float* kvh = kv + h * dim * dim;
  for (int r = 0; r < cnt; ++r) {
    const float* q = qs + r * c + h * dim;
    float num = 0.f, den = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float qd = q[d];
      num = fmaf(qd, kvh[d * dim + v], num);
      den = fmaf(qd, ks[h * dim + d], den);
    }
    const float z = 1.f / (den + 1e-6f);
    out[((int64_t)n * l + l0 + r) * c + t] = num * z * s_len;
  }
}

// Fine-level: one CTA per 25-token window; blockDim = heads*dim (=128).  Both phases fused.
__global__ void linattn_window_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                                      const float* __restrict__ V, int ldv, float* __restrict__ out, int tokens,
                                      int heads, int dim) {
  extern __shared__ float sh[];
  const int c = heads * dim;
  float* qs = sh;                  // [tokens][c]
  float* ks = qs + tokens * c;
  float* vs = ks + tokens * c;
  float* ksum = vs + tokens * c;   // [c]
  const int64_t w = blockIdx.x;
  const int t = threadIdx.x;
  const float inv_s = 1.f / (float)tokens;
  for (int r = 0; r < tokens; ++r) {
    const int64_t row = w * tokens + r;
    qs[r * c + t] = Q[row * ldq + t];
    ks[r * c + t] = K[row * ldk + t];
    vs[r * c + t] = V[row * ldv + t] / (float)tokens;
  }
  (void)inv_s;
  float a = 0.f;
  for (int r = 0; r < tokens; ++r) a += ks[r * c + t];
  ksum[t] = a;
  __syncthreads();
  const int h = t / dim, v = t - h * dim;
  // KV[h, d, v] for this thread's (h, v): dim values kept in registers (dim <= 16 at the fine level)
  float kvcol[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) kvcol[d] = 0.f;
  for (int r = 0; r < tokens; ++r) {
    const float vv = vs[r * c + t];
#pragma unroll
    for (int d = 0; d < 16; ++d) if (d < dim) kvcol[d] = fmaf(ks[r * c + h * dim + d], vv, kvcol[d]);
  }
  for (int r = 0; r < tokens; ++r) {
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) if (d < dim) {
      const float qd = qs[r * c + h * dim + d];
      num = fmaf(qd, kvcol[d], num);
      den = fmaf(qd, ksum[h * dim + d], den);
    }
    out[(w * tokens + r) * c + t] = num * (1.f / (den + 1e-6f)) * (float)tokens;
  }
}

__global__ void add_posenc_kernel(const float4* __restrict__ x, const float4* __restrict__ pe, float4* __restrict__ out,
                                  int64_t per_sample4, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const float4 a = x[i], b = pe[i % per_sample4];
  out[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__global__ void select_rows_kernel(float4* __restrict__ dst, const float4* __restrict__ src, const int* __restrict__ flag,
                                   int64_t per_sample4, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  if (flag[i / per_sample4] == 0) dst[i] = src[i];
}

__global__ void gather_rows_kernel(const float* __restrict__ feat, int64_t l, int c, const int64_t* __restrict__ b_ids,
                                   const int64_t* __restrict__ tok, int64_t m, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = c >> 2;
  if (i >= m * c4) return;
  const int64_t r = i / c4;
  const int q = (int)(i - r * c4);
  const float4* src = reinterpret_cast<const float4*>(feat + (b_ids[r] * l + tok[r]) * c);
  reinterpret_cast<float4*>(out)[i] = src[q];
}

}  // namespace gf

using namespace gf;
#define STREAM ((cudaStream_t)stream)

extern "C" int64_t gf_linattn_partial_floats(int n, int s, int heads, int dim) {
  return (int64_t)n * heads * gf_cdiv(s, kChunk) * (dim * dim + dim);
}

extern "C" int gf_linattn_reduce(const float* K, int ldk, const float* V, int ldv, int n, int s, int heads, int dim,
                                 float* partial, float* KV, float* Ksum, gf_stream_t stream) {
  if (n <= 0 || s <= 0 || heads <= 0 || dim <= 0 || dim > 32) return gf_set_error(GF_ERR_ARG, "gf_linattn_reduce: dim must be <= 32");
  const int nchunks = gf_cdiv(s, kChunk);
  const size_t smem = 2 * 64 * dim * sizeof(float);
  linattn_partial_kernel<<<dim3(nchunks, heads, n), 256, smem, STREAM>>>(K, ldk, V, ldv, s, heads, dim, 1.f / (float)s, partial);
  linattn_finalize_kernel<<<n * heads, 256, 0, STREAM>>>(partial, nchunks, dim, KV, Ksum);
  g_launches += 2;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_linattn_apply(const float* Q, int ldq, const float* KV, const float* Ksum, float* out, int n, int l,
                                int s, int heads, int dim, gf_stream_t stream) {
  const int c = heads * dim;
  if (n <= 0 || l <= 0 || c > 1024 || (c % 32)) return gf_set_error(GF_ERR_ARG, "gf_linattn_apply: bad shape");
  const size_t smem = (size_t)(heads * dim * dim + c + 32 * c) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(linattn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
  if (smem > 96 * 1024) return gf_set_error(GF_ERR_ARG, "gf_linattn_apply: shared memory");
  linattn_apply_kernel<<<dim3(gf_cdiv(l, 32), n), c, smem, STREAM>>>(Q, ldq, KV, Ksum, out, l, heads, dim, (float)s);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_linattn_window(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* out,
                                 int64_t n_windows, int tokens, int heads, int dim, gf_stream_t stream) {
  const int c = heads * dim;
  if (n_windows < 0 || tokens <= 0 || tokens > 64 || dim > 16 || c > 1024 || (c % 32))
    return gf_set_error(GF_ERR_ARG, "gf_linattn_window: bad shape");
  if (n_windows == 0) return GF_OK;
  const size_t smem = (size_t)(3 * tokens * c + c) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(linattn_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
  if (smem > 96 * 1024) return gf_set_error(GF_ERR_ARG, "gf_linattn_window: shared memory");
  linattn_window_kernel<<<(unsigned)n_windows, c, smem, STREAM>>>(Q, ldq, K, ldk, V, ldv, out, tokens, heads, dim);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_add_posenc(const float* x, const float* pe, float* out, int n, int64_t l, int c, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || c <= 0 || (c % 4)) return gf_set_error(GF_ERR_ARG, "gf_add_posenc: bad shape");
  const int64_t per = l * c / 4, total = per * n;
  add_posenc_kernel<<<gf_cdiv(total, 256), 256, 0, STREAM>>>((const float4*)x, (const float4*)pe, (float4*)out, per, total);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_select_rows(float* dst, const float* src, const int* flag, int n, int64_t l, int c, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || c <= 0 || (c % 4)) return gf_set_error(GF_ERR_ARG, "gf_select_rows: bad shape");
  const int64_t per = l * c / 4, total = per * n;
  select_rows_kernel<<<gf_cdiv(total, 256), 256, 0, STREAM>>>((float4*)dst, (const float4*)src, flag, per, total);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_gather_rows(const float* feat, int64_t l, int c, const int64_t* b_ids, const int64_t* tok_ids,
                              int64_t m, float* out, gf_stream_t stream) {
  if (m < 0 || c <= 0 || (c % 4)) return gf_set_error(GF_ERR_ARG, "gf_gather_rows: bad shape");
  if (m == 0) return GF_OK;
  gather_rows_kernel<<<gf_cdiv(m * (c / 4), 256), 256, 0, STREAM>>>(feat, l, c, b_ids, tok_ids, m, out);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
