// Linear attention (model/loftr_src/loftr/loftr_module/linear_attention.py:33-49) and small
// element-wise helpers.  Q and K arrive already feature-mapped (elu+1 fused into the projection epilogue).
#include "common.cuh"

#include <cuda_fp16.h>

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

// 4 consecutive elements as float4 / one element from float: the kernels are templated on the storage type of the
// projected Q/K/V and of the message (fp32, or fp16 written by the OUT16 GEMM epilogue); all arithmetic stays fp32.
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }

constexpr int kChunk = 256;      // source tokens per partial reduction block (fp32 kernels)
constexpr int kChunkMma = 128;   // same for the mma.sync kernel (2 resident CTAs per SM: more, smaller blocks)

// ---------------------------------------------------------------------------------------------
// reduce, stage 1.  One CTA per (256-token chunk, sample); warp w == head w (blockDim = 32 * heads).
// Lane v owns column v of the head's dim x dim block: acc[d] += K[s,h,d] * V[s,h,v] with K broadcast
// from shared memory as float4 (dim/4 LDS.128 + 1 LDS per dim FMAs).  dim == 32.
// partial[(n, chunk, h)][dim*dim + dim]: KV block (row d, col v) then Ksum.
// ---------------------------------------------------------------------------------------------
template <int D, typename T>
__global__ void linattn_partial_kernel(const T* __restrict__ K, int ldk, const T* __restrict__ V, int ldv,
                                       int s, int heads, float inv_s, float* __restrict__ partial) {
  extern __shared__ float sh[];
  const int c = heads * D;
  float* Ks = sh;                 // [32 tokens][c]
  float* Vs = sh + 32 * c;        // [32 tokens][c]
  const int chunk = blockIdx.x, n = blockIdx.y;
  const int nchunks = gridDim.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = chunk * kChunk, s1 = min(s, s0 + kChunk);
  float acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  float ksum = 0.f;
  for (int t0 = s0; t0 < s1; t0 += 32) {
    const int cnt = min(32, s1 - t0);
    for (int e = threadIdx.x; e < 32 * (c / 4); e += blockDim.x) {
      const int r = e / (c / 4), q = e - r * (c / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (r < cnt) {
        const int64_t tok = (int64_t)n * s + t0 + r;
        kv = ld4<T>(K + tok * ldk + 4 * q);
        vv = ld4<T>(V + tok * ldv + 4 * q);
        vv.x *= inv_s; vv.y *= inv_s; vv.z *= inv_s; vv.w *= inv_s;      // values / v_length (linear_attention.py:46)
      }
      *reinterpret_cast<float4*>(Ks + r * c + 4 * q) = kv;
      *reinterpret_cast<float4*>(Vs + r * c + 4 * q) = vv;
    }
    __syncthreads();
    if (lane < D) {
      for (int r = 0; r < cnt; ++r) {
        const float vv = Vs[r * c + h * D + lane];
        ksum += Ks[r * c + h * D + lane];
        const float4* kr = reinterpret_cast<const float4*>(Ks + r * c + h * D);
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
          const float4 k4 = kr[q];
          acc[4 * q] = fmaf(k4.x, vv, acc[4 * q]); acc[4 * q + 1] = fmaf(k4.y, vv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(k4.z, vv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(k4.w, vv, acc[4 * q + 3]);
        }
      }
    }
    __syncthreads();
  }
  if (lane < D) {
    float* out = partial + (((int64_t)n * nchunks + chunk) * heads + h) * (D * D + D);
#pragma unroll
    for (int d = 0; d < D; ++d) out[d * D + lane] = acc[d];
    out[D * D + lane] = ksum;
  }
}

// reduce, stage 2: one thread per (sample, head, entry); the chunk partials of an entry are nchunks independent,
// coalesced loads (4 in flight per thread)
__global__ void linattn_finalize_kernel(const float* __restrict__ partial, int nchunks, int heads, int dim,
                                        float* __restrict__ KV, float* __restrict__ Ksum) {
  const int entries = dim * dim, per = entries + dim;
  const int nh = blockIdx.y;                      // sample * heads + head
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per) return;
  const int n = nh / heads, h = nh - n * heads;
  const float* p = partial + ((int64_t)n * nchunks * heads + h) * per + e;
  const int64_t stride = (int64_t)heads * per;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = 0;
  for (; c + 4 <= nchunks; c += 4) {
    a0 += p[(c + 0) * stride]; a1 += p[(c + 1) * stride]; a2 += p[(c + 2) * stride]; a3 += p[(c + 3) * stride];
  }
  for (; c < nchunks; ++c) a0 += p[c * stride];
  const float a = (a0 + a1) + (a2 + a3);
  if (e < entries) KV[(int64_t)nh * entries + e] = a;
  else Ksum[(int64_t)nh * dim + (e - entries)] = a;
}

// ---------------------------------------------------------------------------------------------
// apply.  out[n,l,h,v] = (sum_d Q[l,h,d] KV[h,d,v]) * (1 / (Q[l,h,:].Ksum[h,:] + eps)) * S
// blockDim = heads*dim (thread == output channel); the thread keeps its KV column and the head's Ksum in
// registers; 128 tokens per CTA, Q rows staged in shared memory and read as broadcast float4.
// ---------------------------------------------------------------------------------------------
template <int D, typename T>
__global__ void linattn_apply_kernel(const T* __restrict__ Q, int ldq, const float* __restrict__ KV,
                                     const float* __restrict__ Ksum, T* __restrict__ out, int l, int heads,
                                     float s_len) {
  extern __shared__ float sh[];
  const int c = heads * D;
  float* qs = sh;                       // [32][c]
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int h = t / D, v = t - h * D;
  float kvcol[D], ks[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    kvcol[d] = KV[(((int64_t)n * heads + h) * D + d) * D + v];
    ks[d] = Ksum[((int64_t)n * heads + h) * D + d];
  }
  const int l_begin = blockIdx.x * 128, l_end = min(l, l_begin + 128);
  for (int l0 = l_begin; l0 < l_end; l0 += 32) {
    const int cnt = min(32, l_end - l0);
    __syncthreads();
    for (int e = t; e < 32 * (c / 4); e += blockDim.x) {
      const int r = e / (c / 4), q = e - r * (c / 4);
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < cnt) qv = ld4<T>(Q + ((int64_t)n * l + l0 + r) * ldq + 4 * q);
      *reinterpret_cast<float4*>(qs + r * c + 4 * q) = qv;
    }
    __syncthreads();
    for (int r = 0; r < cnt; ++r) {
      const float4* qr = reinterpret_cast<const float4*>(qs + r * c + h * D);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        const float4 q4 = qr[q];
        num = fmaf(q4.x, kvcol[4 * q], num); den = fmaf(q4.x, ks[4 * q], den);
        num = fmaf(q4.y, kvcol[4 * q + 1], num); den = fmaf(q4.y, ks[4 * q + 1], den);
        num = fmaf(q4.z, kvcol[4 * q + 2], num); den = fmaf(q4.z, ks[4 * q + 2], den);
        num = fmaf(q4.w, kvcol[4 * q + 3], num); den = fmaf(q4.w, ks[4 * q + 3], den);
      }
      st1(out + ((int64_t)n * l + l0 + r) * c + t, num * (1.f / (den + 1e-6f)) * s_len);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fine level: one CTA per 25-token window; blockDim = heads*dim (=128), dim == 16.  Both phases fused;
// K / Q rows are read from shared memory as broadcast float4.
// ---------------------------------------------------------------------------------------------
template <int D, int TOK>
__global__ void linattn_window_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
                                      const float* __restrict__ V, int ldv, float* __restrict__ out, int tokens_rt,
                                      int heads) {
  const int tokens = TOK > 0 ? TOK : tokens_rt;     // compile-time token count -> fully unrolled, 3*TOK loads in flight
  extern __shared__ float sh[];
  const int c = heads * D;
  float* qs = sh;                  // [tokens][c]
  float* ks = qs + tokens * c;
  float* vs = ks + tokens * c;
  float* ksum = vs + tokens * c;   // [c]
  const int64_t w = blockIdx.x;
  const int t = threadIdx.x;
  const float ftok = (float)tokens;
  float a = 0.f;
  if constexpr (TOK > 0) {
    float qr[TOK], kr[TOK], vr[TOK];
#pragma unroll
    for (int r = 0; r < TOK; ++r) {
      const int64_t row = w * TOK + r;
      qr[r] = __ldg(Q + row * ldq + t); kr[r] = __ldg(K + row * ldk + t); vr[r] = __ldg(V + row * ldv + t);
    }
#pragma unroll
    for (int r = 0; r < TOK; ++r) {
      qs[r * c + t] = qr[r]; ks[r * c + t] = kr[r]; a += kr[r]; vs[r * c + t] = vr[r] / ftok;
    }
  } else {
    for (int r = 0; r < tokens; ++r) {
      const int64_t row = w * tokens + r;
      qs[r * c + t] = Q[row * ldq + t];
      const float kk = K[row * ldk + t];
      ks[r * c + t] = kk;
      a += kk;
      vs[r * c + t] = V[row * ldv + t] / ftok;
    }
  }
  ksum[t] = a;
  __syncthreads();
  const int h = t / D;
  float kvcol[D];
#pragma unroll
  for (int d = 0; d < D; ++d) kvcol[d] = 0.f;
  for (int r = 0; r < tokens; ++r) {
    const float vv = vs[r * c + t];
    const float4* kr = reinterpret_cast<const float4*>(ks + r * c + h * D);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 k4 = kr[q];
      kvcol[4 * q] = fmaf(k4.x, vv, kvcol[4 * q]); kvcol[4 * q + 1] = fmaf(k4.y, vv, kvcol[4 * q + 1]);
      kvcol[4 * q + 2] = fmaf(k4.z, vv, kvcol[4 * q + 2]); kvcol[4 * q + 3] = fmaf(k4.w, vv, kvcol[4 * q + 3]);
    }
  }
  float ksr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) ksr[d] = ksum[h * D + d];
  for (int r = 0; r < tokens; ++r) {
    const float4* qr = reinterpret_cast<const float4*>(qs + r * c + h * D);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 q4 = qr[q];
      num = fmaf(q4.x, kvcol[4 * q], num); den = fmaf(q4.x, ksr[4 * q], den);
      num = fmaf(q4.y, kvcol[4 * q + 1], num); den = fmaf(q4.y, ksr[4 * q + 1], den);
      num = fmaf(q4.z, kvcol[4 * q + 2], num); den = fmaf(q4.z, ksr[4 * q + 2], den);
      num = fmaf(q4.w, kvcol[4 * q + 3], num); den = fmaf(q4.w, ksr[4 * q + 3], den);
    }
    out[(w * tokens + r) * c + t] = num * (1.f / (den + 1e-6f)) * ftok;
  }
}

// ---------------------------------------------------------------------------------------------
// fp16-storage path on the warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate): 8 heads of dim 32.
// Both products of the linear attention have a tiny M/N (32 x 32 per head) and a long or streaming third dimension,
// so they run as register-fragment MMAs with the operands staged once in shared memory:
//   reduce: C' = V_h^T K_h over the tokens (tokens = MMA K); a constant "ones" A tile yields Ksum = 1^T K_h for free
//   apply : out = Q_h (KV_h | Ksum_h/S): the KV block is held as B fragments in registers for the whole CTA
// Rows are padded to 528 B in shared memory so that every ldmatrix phase (8 rows x 16 B) is conflict-free.
// ---------------------------------------------------------------------------------------------
constexpr int kPitch = 528;          // bytes per staged row: 256 fp16 channels + 16 B pad

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// reduce, stage 1 (same partial[] format as linattn_partial_kernel).  grid (chunks of 256 tokens, n), 256 threads.
__global__ void __launch_bounds__(256)
linattn_partial_mma_kernel(const __half* __restrict__ K, int ldk, const __half* __restrict__ V, int ldv, int s,
                           int chunk_tokens, float inv_s, float* __restrict__ partial) {
  constexpr int D = 32, HEADS = 8, SLAB = 32, BUF = 2 * SLAB * kPitch;       // one buffer: K slab | V slab
  extern __shared__ __align__(16) uint8_t dyn_smem[];                        // two buffers, filled by cp.async one slab ahead
  const int chunk = blockIdx.x, n = blockIdx.y, nchunks = gridDim.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3, lrow = lane & 7, lmat = lane >> 3;
  const int s0 = chunk * chunk_tokens, s1 = min(s, s0 + chunk_tokens);
  float cv[2][4][4], co[4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) cv[a][b][c] = 0.f;
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int c = 0; c < 4; ++c) co[b][c] = 0.f;
  const uint32_t sm_a = smem_u32(dyn_smem);
  const uint32_t one = (g == 0) ? 0x3C003C00u : 0u;
  const uint32_t ao[4] = {one, 0u, one, 0u};
  auto issue = [&](int t0, int buf) {           // rows past the end of the sequence are zero-filled (src-size 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + 256 * i, r = e >> 5, q = e & 31;
      const bool ok = t0 + r < s1;
      const int64_t tok = (int64_t)n * s + (ok ? t0 + r : s0);
      const uint32_t dst = sm_a + buf * BUF + r * kPitch + q * 16;
      const uint32_t nbytes = ok ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(K + tok * ldk + q * 8), "r"(nbytes) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + SLAB * kPitch), "l"(V + tok * ldv + q * 8), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(s0, 0);
  int buf = 0;
  for (int t0 = s0; t0 < s1; t0 += SLAB, buf ^= 1) {
    if (t0 + SLAB < s1) { issue(t0 + SLAB, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint32_t ks_a = sm_a + buf * BUF, vs_a = ks_a + SLAB * kPitch;
#pragma unroll
    for (int kk = 0; kk < SLAB / 16; ++kk) {
      uint32_t b[2][4];
#pragma unroll
      for (int ip = 0; ip < 2; ++ip) {      // n-tiles (d1) 2ip, 2ip+1: matrices (tok lo, c), (tok hi, c), (tok lo, c+1), (tok hi, c+1)
        const int row = kk * 16 + lrow + ((lmat & 1) << 3), c16 = 4 * h + 2 * ip + (lmat >> 1);
        ldsm4t(b[ip], ks_a + row * kPitch + c16 * 16);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {      // d2 = 16 mt ..: matrices (tok lo, ch lo), (tok lo, ch hi), (tok hi, ch lo), (tok hi, ch hi)
        uint32_t a[4];
        const int row = kk * 16 + lrow + ((lmat >> 1) << 3), c16 = 4 * h + 2 * mt + (lmat & 1);
        ldsm4t(a, vs_a + row * kPitch + c16 * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_f16(cv[mt][i], a, b[i >> 1][2 * (i & 1)], b[i >> 1][2 * (i & 1) + 1]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) mma_f16(co[i], ao, b[i >> 1][2 * (i & 1)], b[i >> 1][2 * (i & 1) + 1]);
    }
    __syncthreads();                             // this buffer is refilled by the cp.async issued in the next iteration
  }
  float* out = partial + (((int64_t)n * nchunks + chunk) * HEADS + h) * (D * D + D);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) {           // cv[mt][i]: rows d2 = 16mt + g (+8), cols d1 = 8i + 2t (+1); stored as KV[d1][d2]
      const int d2 = 16 * mt + g, d1 = 8 * i + 2 * t;
      out[d1 * D + d2] = cv[mt][i][0] * inv_s; out[(d1 + 1) * D + d2] = cv[mt][i][1] * inv_s;
      out[d1 * D + d2 + 8] = cv[mt][i][2] * inv_s; out[(d1 + 1) * D + d2 + 8] = cv[mt][i][3] * inv_s;
    }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[D * D + 8 * i + 2 * t] = co[i][0]; out[D * D + 8 * i + 2 * t + 1] = co[i][1]; }
  }
}

// apply.  grid (ceil(l / 64), n), 256 threads (warp == head); the message overwrites Q in shared memory and leaves
// as full 512-byte rows.
__global__ void __launch_bounds__(256)
linattn_apply_mma_kernel(const __half* __restrict__ Q, int ldq, const float* __restrict__ KV, const float* __restrict__ Ksum,
                         __half* __restrict__ out, int l, float s_len) {
  constexpr int D = 32, HEADS = 8, ROWS = 64, C = HEADS * D;
  __shared__ __align__(16) uint8_t qs[ROWS * kPitch];
  const int n = blockIdx.y, l0 = blockIdx.x * ROWS;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3, lrow = lane & 7, lmat = lane >> 3;
  const int rows = min(ROWS, l - l0);
  for (int e = threadIdx.x; e < ROWS * 32; e += 256) {
    const int r = e >> 5, q = e & 31;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows) v = __ldg(reinterpret_cast<const uint4*>(Q + ((int64_t)n * l + l0 + r) * ldq) + q);
    *reinterpret_cast<uint4*>(qs + r * kPitch + q * 16) = v;
  }
  // B fragments: k = d1 (two k-steps), n = d2 (four tiles) + one tile whose column 0 is Ksum / S
  const float* kv = KV + ((int64_t)n * HEADS + h) * D * D;
  const float* ksum = Ksum + ((int64_t)n * HEADS + h) * D;
  const float inv_s = 1.f / s_len;
  uint32_t b[2][5][2];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d1 = 16 * kk + 2 * t, d2 = 8 * j + g;
      b[kk][j][0] = pk2(kv[d1 * D + d2], kv[(d1 + 1) * D + d2]);
      b[kk][j][1] = pk2(kv[(d1 + 8) * D + d2], kv[(d1 + 9) * D + d2]);
    }
    const int d1 = 16 * kk + 2 * t;
    b[kk][4][0] = (g == 0) ? pk2(ksum[d1] * inv_s, ksum[d1 + 1] * inv_s) : 0u;
    b[kk][4][1] = (g == 0) ? pk2(ksum[d1 + 8] * inv_s, ksum[d1 + 9] * inv_s) : 0u;
  }
  __syncthreads();
  const uint32_t qs_a = smem_u32(qs);
  const float eps = 1e-6f * inv_s;
#pragma unroll 1
  for (int mt = 0; mt < ROWS / 16; ++mt) {
    float acc[5][4];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {      // matrices (tok lo, k lo), (tok hi, k lo), (tok lo, k hi), (tok hi, k hi)
      uint32_t a[4];
      const int row = 16 * mt + lrow + ((lmat & 1) << 3), c16 = 4 * h + 2 * kk + (lmat >> 1);
      ldsm4(a, qs_a + row * kPitch + c16 * 16);
#pragma unroll
      for (int j = 0; j < 5; ++j) mma_f16(acc[j], a, b[kk][j][0], b[kk][j][1]);
    }
    const float z_lo = 1.f / (__shfl_sync(0xffffffffu, acc[4][0], lane & ~3) + eps);
    const float z_hi = 1.f / (__shfl_sync(0xffffffffu, acc[4][2], lane & ~3) + eps);
    uint8_t* row_lo = qs + (16 * mt + g) * kPitch + (D * h + 2 * t) * 2;
    uint8_t* row_hi = row_lo + 8 * kPitch;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<uint32_t*>(row_lo + 16 * j) = pk2(acc[j][0] * z_lo, acc[j][1] * z_lo);
      *reinterpret_cast<uint32_t*>(row_hi + 16 * j) = pk2(acc[j][2] * z_hi, acc[j][3] * z_hi);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < rows * 32; e += 256) {
    const int r = e >> 5, q = e & 31;
    reinterpret_cast<uint4*>(out + ((int64_t)n * l + l0 + r) * C)[q] = *reinterpret_cast<const uint4*>(qs + r * kPitch + q * 16);
  }
}

// fp16 storage variant (Q/K/V written by the OUT16 projection, message read by the fp16-operand merge GEMM):
// blockDim == heads * D == 128; thread t loads channel pair 2*(t & 63) of the rows of parity t >> 6 as half2 (full
// 128-byte warp requests), the message rows are staged in shared memory and leave as one flat 16-byte-vector copy.
template <int D, int TOK>
__global__ void __launch_bounds__(128)
linattn_window16_kernel(const __half* __restrict__ Q, int ldq, const __half* __restrict__ K, int ldk,
                        const __half* __restrict__ V, int ldv, __half* __restrict__ out) {
  constexpr int c = 128, NR = (TOK + 1) / 2;
  extern __shared__ float sh[];
  float* qs = sh;                  // [TOK][c]
  float* ks = qs + TOK * c;
  float* vs = ks + TOK * c;
  float* ksum = vs + TOK * c;      // [c]
  const int64_t w = blockIdx.x;
  const int t = threadIdx.x;
  const int cp = (t & 63) * 2, par = t >> 6;
  constexpr float ftok = (float)TOK;
  {
    __half2 qr[NR], kr[NR], vr[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = 2 * i + par;
      if (r < TOK) {
        const int64_t row = w * TOK + r;
        qr[i] = __ldg(reinterpret_cast<const __half2*>(Q + row * ldq + cp));
        kr[i] = __ldg(reinterpret_cast<const __half2*>(K + row * ldk + cp));
        vr[i] = __ldg(reinterpret_cast<const __half2*>(V + row * ldv + cp));
      }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = 2 * i + par;
      if (r < TOK) {
        *reinterpret_cast<float2*>(qs + r * c + cp) = __half22float2(qr[i]);
        *reinterpret_cast<float2*>(ks + r * c + cp) = __half22float2(kr[i]);
        const float2 v2 = __half22float2(vr[i]);
        *reinterpret_cast<float2*>(vs + r * c + cp) = make_float2(v2.x / ftok, v2.y / ftok);
      }
    }
  }
  __syncthreads();
  {
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < TOK; ++r) a += ks[r * c + t];
    ksum[t] = a;
  }
  const int h = t / D;
  float kvcol[D];
#pragma unroll
  for (int d = 0; d < D; ++d) kvcol[d] = 0.f;
#pragma unroll 5
  for (int r = 0; r < TOK; ++r) {
    const float vv = vs[r * c + t];
    const float4* kr = reinterpret_cast<const float4*>(ks + r * c + h * D);
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 k4 = kr[q];
      kvcol[4 * q] = fmaf(k4.x, vv, kvcol[4 * q]); kvcol[4 * q + 1] = fmaf(k4.y, vv, kvcol[4 * q + 1]);
      kvcol[4 * q + 2] = fmaf(k4.z, vv, kvcol[4 * q + 2]); kvcol[4 * q + 3] = fmaf(k4.w, vv, kvcol[4 * q + 3]);
    }
  }
  __syncthreads();                 // ksum complete; every thread is done with vs (reused below as the output stage)
  float ksr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) ksr[d] = ksum[h * D + d];
  __half* os = reinterpret_cast<__half*>(vs);
#pragma unroll 5
  for (int r = 0; r < TOK; ++r) {
    const float4* qr = reinterpret_cast<const float4*>(qs + r * c + h * D);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 q4 = qr[q];
      num = fmaf(q4.x, kvcol[4 * q], num); den = fmaf(q4.x, ksr[4 * q], den);
      num = fmaf(q4.y, kvcol[4 * q + 1], num); den = fmaf(q4.y, ksr[4 * q + 1], den);
      num = fmaf(q4.z, kvcol[4 * q + 2], num); den = fmaf(q4.z, ksr[4 * q + 2], den);
      num = fmaf(q4.w, kvcol[4 * q + 3], num); den = fmaf(q4.w, ksr[4 * q + 3], den);
    }
    os[r * c + t] = __float2half_rn(num * (1.f / (den + 1e-6f)) * ftok);
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + w * (TOK * c));
  const uint4* src = reinterpret_cast<const uint4*>(os);
  for (int i = t; i < TOK * c / 8; i += 128) dst[i] = src[i];
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
  return (int64_t)n * heads * gf_cdiv(s, kChunkMma) * (dim * dim + dim);      // sized for the finer of the two chunkings
}

template <typename T>
static int linattn_reduce_impl(const T* K, int ldk, const T* V, int ldv, int n, int s, int heads, int dim,
                               float* partial, float* KV, float* Ksum, gf_stream_t stream) {
  if (n <= 0 || s <= 0 || heads <= 0 || heads > 32 || dim != 32 || (ldk % 4) || (ldv % 4))
    return gf_set_error(GF_ERR_ARG, "gf_linattn_reduce: dim must be 32, heads <= 32, 16-byte aligned rows");
  const int nchunks = gf_cdiv(s, kChunk);
  const size_t smem = (size_t)2 * 32 * heads * dim * sizeof(float);
  GF_SMEM_OPTIN((linattn_partial_kernel<32, T>), 96 * 1024);
  if (smem > 96 * 1024) return gf_set_error(GF_ERR_ARG, "gf_linattn_reduce: shared memory");
  linattn_partial_kernel<32, T><<<dim3(nchunks, n), 32 * heads, smem, STREAM>>>(K, ldk, V, ldv, s, heads, 1.f / (float)s, partial);
  linattn_finalize_kernel<<<dim3(gf_cdiv(dim * dim + dim, 128), n * heads), 128, 0, STREAM>>>(partial, nchunks, heads, dim, KV, Ksum);
  g_launches += 2;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

template <typename T>
static int linattn_apply_impl(const T* Q, int ldq, const float* KV, const float* Ksum, T* out, int n, int l,
                              int s, int heads, int dim, gf_stream_t stream) {
  const int c = heads * dim;
  if (n <= 0 || l <= 0 || c > 1024 || dim != 32 || (ldq % 4)) return gf_set_error(GF_ERR_ARG, "gf_linattn_apply: dim must be 32");
  const size_t smem = (size_t)32 * c * sizeof(float);
  linattn_apply_kernel<32, T><<<dim3(gf_cdiv(l, 128), n), c, smem, STREAM>>>(Q, ldq, KV, Ksum, out, l, heads, (float)s);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_linattn_reduce(const float* K, int ldk, const float* V, int ldv, int n, int s, int heads, int dim,
                                 float* partial, float* KV, float* Ksum, gf_stream_t stream) {
  return linattn_reduce_impl<float>(K, ldk, V, ldv, n, s, heads, dim, partial, KV, Ksum, stream);
}
extern "C" int gf_linattn_apply(const float* Q, int ldq, const float* KV, const float* Ksum, float* out, int n, int l,
                                int s, int heads, int dim, gf_stream_t stream) {
  return linattn_apply_impl<float>(Q, ldq, KV, Ksum, out, n, l, s, heads, dim, stream);
}
// fp16-storage variants (Q/K/V and the message in fp16; KV / Ksum and all arithmetic in fp32)
extern "C" int gf_linattn_reduce_f16(const void* K, int ldk, const void* V, int ldv, int n, int s, int heads, int dim,
                                     float* partial, float* KV, float* Ksum, gf_stream_t stream) {
  if (heads == 8 && dim == 32 && n > 0 && s > 0 && (ldk % 8) == 0 && (ldv % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(K) % 16) == 0 && (reinterpret_cast<uintptr_t>(V) % 16) == 0) {
    // ~600 blocks (two resident per SM) whatever the batch: fewer, longer chunks keep the partial-sum traffic small
    int chunk_tokens = (int)(((int64_t)s * n / 592 + 31) / 32 * 32);
    chunk_tokens = chunk_tokens < kChunkMma ? kChunkMma : (chunk_tokens > 512 ? 512 : chunk_tokens);
    const int nchunks = gf_cdiv(s, chunk_tokens);
    constexpr int kSmem = 2 * 2 * 32 * kPitch;
    GF_SMEM_OPTIN(linattn_partial_mma_kernel, kSmem);
    linattn_partial_mma_kernel<<<dim3(nchunks, n), 256, kSmem, STREAM>>>((const __half*)K, ldk, (const __half*)V, ldv, s,
                                                                        chunk_tokens, 1.f / (float)s, partial);
    linattn_finalize_kernel<<<dim3(gf_cdiv(dim * dim + dim, 128), n * heads), 128, 0, STREAM>>>(partial, nchunks, heads, dim, KV, Ksum);
    g_launches += 2;
    GF_CHECK_LAUNCH();
    return GF_OK;
  }
  return linattn_reduce_impl<__half>((const __half*)K, ldk, (const __half*)V, ldv, n, s, heads, dim, partial, KV, Ksum, stream);
}
extern "C" int gf_linattn_apply_f16(const void* Q, int ldq, const float* KV, const float* Ksum, void* out, int n, int l,
                                    int s, int heads, int dim, gf_stream_t stream) {
  if (heads == 8 && dim == 32 && n > 0 && l > 0 && (ldq % 8) == 0 && (reinterpret_cast<uintptr_t>(Q) % 16) == 0) {
    linattn_apply_mma_kernel<<<dim3(gf_cdiv(l, 64), n), 256, 0, STREAM>>>((const __half*)Q, ldq, KV, Ksum, (__half*)out, l, (float)s);
    g_launches++;
    GF_CHECK_LAUNCH();
    return GF_OK;
  }
  return linattn_apply_impl<__half>((const __half*)Q, ldq, KV, Ksum, (__half*)out, n, l, s, heads, dim, stream);
}
extern "C" int gf_linattn_window_f16(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* out,
                                     int64_t n_windows, int tokens, int heads, int dim, gf_stream_t stream) {
  if (n_windows < 0 || tokens != 25 || dim != 16 || heads != 8 || (ldq % 2) || (ldk % 2) || (ldv % 2))
    return gf_set_error(GF_ERR_ARG, "gf_linattn_window_f16: needs 25-token windows, 8 heads of dim 16");
  if (n_windows == 0) return GF_OK;
  const size_t smem = (size_t)(3 * 25 * 128 + 128) * sizeof(float);
  GF_SMEM_OPTIN((linattn_window16_kernel<16, 25>), 96 * 1024);
  linattn_window16_kernel<16, 25><<<(unsigned)n_windows, 128, smem, STREAM>>>((const __half*)Q, ldq, (const __half*)K, ldk,
                                                                            (const __half*)V, ldv, (__half*)out);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_linattn_window(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* out,
                                 int64_t n_windows, int tokens, int heads, int dim, gf_stream_t stream) {
  const int c = heads * dim;
  if (n_windows < 0 || tokens <= 0 || tokens > 64 || dim != 16 || c > 1024 || (c % 32))
    return gf_set_error(GF_ERR_ARG, "gf_linattn_window: dim must be 16");
  if (n_windows == 0) return GF_OK;
  const size_t smem = (size_t)(3 * tokens * c + c) * sizeof(float);
  GF_SMEM_OPTIN((linattn_window_kernel<16, 0>), 96 * 1024);
  GF_SMEM_OPTIN((linattn_window_kernel<16, 25>), 96 * 1024);
  if (smem > 96 * 1024) return gf_set_error(GF_ERR_ARG, "gf_linattn_window: shared memory");
  if (tokens == 25) linattn_window_kernel<16, 25><<<(unsigned)n_windows, c, smem, STREAM>>>(Q, ldq, K, ldk, V, ldv, out, tokens, heads);
  else              linattn_window_kernel<16, 0><<<(unsigned)n_windows, c, smem, STREAM>>>(Q, ldq, K, ldk, V, ldv, out, tokens, heads);
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
