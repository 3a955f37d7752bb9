// Fine-level kernels (loftr_module/fine_preprocess.py:41-56, model/fine_matching2.py:52-126).
#include "common.cuh"

#include <atomic>
#include <cuda_fp16.h>

namespace gf {
extern std::atomic<int64_t> g_launches;

// same gather from an fp16 NHWC fine map (the tcgen05 backbone's native output): 8-byte loads; windows out as fp32, or as
// fp16 (a copy) when they feed the fp16-operand merge GEMM
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a); u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename TO>
__global__ void fine_gather_f16_kernel(const __half* __restrict__ fine, int hf, int wf, int c,
                                        const int64_t* __restrict__ b_ids, const int64_t* __restrict__ tok_ids, int wc,
                                        int stride, int window, TO* __restrict__ out) {
  const int64_t m = blockIdx.x;
  const int b = (int)b_ids[m];
  const int tok = (int)tok_ids[m];
  const int cy = (tok / wc) * stride, cx = (tok % wc) * stride;
  const int half = window / 2;
  const int c4 = c >> 2;
  TO* o = out + m * window * window * c;
  for (int e = threadIdx.x; e < window * window * c4; e += blockDim.x) {
    const int slot = e / c4, q = e - slot * c4;
    const int py = cy + slot / window - half, px = cx + slot % window - half;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (py >= 0 && py < hf && px >= 0 && px < wf) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(fine + (((int64_t)b * hf + py) * wf + px) * c) + q);
      const __half2 lo = *reinterpret_cast<const __half2*>(&raw.x);
      const __half2 hi = *reinterpret_cast<const __half2*>(&raw.y);
      v = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
    }
    store4(o + 4 * e, v);
  }
}

// 5x5 stride-4 pad-2 window of the NHWC fine map around coarse token (row r, col c): pixels
// (stride*r + ky - half, stride*c + kx - half); slot = ky*window + kx; zero outside (F.unfold padding).
// One CTA per match, one thread per float4 of channels.
__global__ void fine_gather_kernel(const float* __restrict__ fine, int hf, int wf, int c, const int64_t* __restrict__ b_ids,
                                   const int64_t* __restrict__ tok_ids, int wc, int stride, int window,
                                   float* __restrict__ out) {
  const int64_t m = blockIdx.x;
  const int b = (int)b_ids[m];
  const int tok = (int)tok_ids[m];
  const int cy = (tok / wc) * stride, cx = (tok % wc) * stride;
  const int half = window / 2;
  const int c4 = c >> 2;
  float4* o = reinterpret_cast<float4*>(out + m * window * window * c);
  for (int e = threadIdx.x; e < window * window * c4; e += blockDim.x) {
    const int slot = e / c4, q = e - slot * c4;
    const int py = cy + slot / window - half, px = cx + slot % window - half;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (py >= 0 && py < hf && px >= 0 && px < wf)
      v = reinterpret_cast<const float4*>(fine + (((int64_t)b * hf + py) * wf + px) * c)[q];
    o[e] = v;
  }
}

// Per match: S = f0 f1^T / (C * T), conf = softmax(S,1)*softmax(S,2), global arg-max (first on ties), thr.
// blockDim = 128; WW <= 25; C <= 128.
__global__ void __launch_bounds__(128)
fine_match_kernel(const float* __restrict__ f0, const float* __restrict__ f1, int ww, int c, float temperature, float thr,
                  int* __restrict__ sel, int* __restrict__ fi, int* __restrict__ fj, float* __restrict__ fconf,
                  float* __restrict__ fine_matrix) {
  __shared__ __align__(16) float a[25][132], b[25][132];     // stride 132: 16-byte aligned rows, conflict-free LDS.128
  __shared__ float S[25][26];
  __shared__ float rmax[25], rsum[25], cmax[25], csum[25];
  __shared__ float best_v[4];
  __shared__ int best_i[4];
  const int64_t m = blockIdx.x;
  const int tid = threadIdx.x;
  const float norm = sqrtf((float)c);
  for (int e = tid; e < ww * c; e += 128) {
    const int r = e / c, k = e - r * c;
    a[r][k] = f0[(m * ww + r) * c + k] / norm;     // feat / C**.5 (fine_matching2.py:52)
    b[r][k] = f1[(m * ww + r) * c + k] / norm;
  }
  __syncthreads();
  for (int e = tid; e < ww * ww; e += 128) {
    const int i = e / ww, j = e - i * ww;
    float acc = 0.f;
    const float4* ar = reinterpret_cast<const float4*>(&a[i][0]);
    const float4* br = reinterpret_cast<const float4*>(&b[j][0]);
#pragma unroll 8
    for (int k = 0; k < c / 4; ++k) {
      const float4 x = ar[k], y = br[k];
      acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    for (int k = c & ~3; k < c; ++k) acc = fmaf(a[i][k], b[j][k], acc);
    S[i][j] = acc / temperature;
  }
  __syncthreads();
  if (tid < ww) {                       // softmax over dim=2 (row i, across j)
    float mx = -INFINITY;
    for (int j = 0; j < ww; ++j) mx = fmaxf(mx, S[tid][j]);
    float su = 0.f;
    for (int j = 0; j < ww; ++j) su += expf(S[tid][j] - mx);
    rmax[tid] = mx; rsum[tid] = su;
  } else if (tid >= 32 && tid < 32 + ww) {   // softmax over dim=1 (column j, across i)
    const int j = tid - 32;
    float mx = -INFINITY;
    for (int i = 0; i < ww; ++i) mx = fmaxf(mx, S[i][j]);
    float su = 0.f;
    for (int i = 0; i < ww; ++i) su += expf(S[i][j] - mx);
    cmax[j] = mx; csum[j] = su;
  }
  __syncthreads();
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int e = tid; e < ww * ww; e += 128) {
    const int i = e / ww, j = e - i * ww;
    const float s = S[i][j];
    const float cf = (expf(s - cmax[j]) / csum[j]) * (expf(s - rmax[i]) / rsum[i]);
    if (fine_matrix) fine_matrix[m * ww * ww + e] = cf;
    if (cf > bv) { bv = cf; bi = e; }   // e increases per thread -> keeps the first index on ties
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { best_v[tid >> 5] = bv; best_i[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 4; ++w)
      if (best_v[w] > bv || (best_v[w] == bv && best_i[w] < bi)) { bv = best_v[w]; bi = best_i[w]; }
    const bool keep = bv > thr;
    sel[m] = keep ? 1 : 0;
    fi[m] = bi / ww; fj[m] = bi % ww;
    fconf[m] = bv;
  }
}

// Warp-per-match variant for the shipped shape (25 cells, 128 channels): lane t < 25 owns the 5 x 5 block
// S[5(t/5) .. +4][5(t%5) .. +4] in registers, so a 4-channel step costs 10 LDS.128 for 100 FMAs (the CTA-per-match
// kernel above needs 2 LDS.128 per 4 FMAs and is shared-memory bound).  Channels are staged 32 at a time
// (2 x 25 rows x 128 B, coalesced), everything else is warp-local: no block-wide barrier, 8 matches per CTA.
__global__ void __launch_bounds__(256)
fine_match_warp_kernel(const float* __restrict__ f0, const float* __restrict__ f1, int64_t m_total, float temperature,
                       float thr, int* __restrict__ sel, int* __restrict__ fi, int* __restrict__ fj,
                       float* __restrict__ fconf, float* __restrict__ fine_matrix) {
  constexpr int WW = 25, C = 128, KC = 32, PITCH = KC + 4, SP = 27;
  extern __shared__ __align__(16) float fm_smem[];          // per warp: a chunk | b chunk | S
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + warp;
  if (m >= m_total) return;
  float* sa = fm_smem + warp * (2 * WW * PITCH + WW * SP + 1);
  float* sb = sa + WW * PITCH;
  float* S = sb + WW * PITCH;
  const float inv_norm = 1.f / sqrtf((float)C);           // feat / C**.5 (fine_matching2.py:52)
  const int ti = (lane < WW ? lane : 0) / 5, tj = (lane < WW ? lane : 0) % 5;
  float acc[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
  const float4* g0 = reinterpret_cast<const float4*>(f0 + m * WW * C);
  const float4* g1 = reinterpret_cast<const float4*>(f1 + m * WW * C);
#pragma unroll 1
  for (int kc = 0; kc < C / KC; ++kc) {
    __syncwarp();
    for (int e = lane; e < WW * (KC / 4); e += 32) {      // 25 rows x 8 float4 per operand
      const int r = e >> 3, q = e & 7;
      float4 x = __ldg(g0 + r * (C / 4) + kc * (KC / 4) + q), y = __ldg(g1 + r * (C / 4) + kc * (KC / 4) + q);
      x.x *= inv_norm; x.y *= inv_norm; x.z *= inv_norm; x.w *= inv_norm;
      y.x *= inv_norm; y.y *= inv_norm; y.z *= inv_norm; y.w *= inv_norm;
      *reinterpret_cast<float4*>(sa + r * PITCH + 4 * q) = x;
      *reinterpret_cast<float4*>(sb + r * PITCH + 4 * q) = y;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KC / 4; ++q) {
      float4 a[5], b[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        a[i] = *reinterpret_cast<const float4*>(sa + (5 * ti + i) * PITCH + 4 * q);
        b[i] = *reinterpret_cast<const float4*>(sb + (5 * tj + i) * PITCH + 4 * q);
      }
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]); acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]); acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
  }
  if (lane < WW) {
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) S[(5 * ti + i) * SP + 5 * tj + j] = acc[i][j] / temperature;
  }
  __syncwarp();
  // lane t < 25: statistics of row t (softmax over j) and of column t (softmax over i)
  const int t = lane < WW ? lane : 0;
  float rmax = -INFINITY, cmax = -INFINITY;
  for (int j = 0; j < WW; ++j) { rmax = fmaxf(rmax, S[t * SP + j]); cmax = fmaxf(cmax, S[j * SP + t]); }
  float rsum = 0.f, csum = 0.f;
  for (int j = 0; j < WW; ++j) { rsum += expf(S[t * SP + j] - rmax); csum += expf(S[j * SP + t] - cmax); }
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int j = 0; j < WW; ++j) {
    const float cm = __shfl_sync(0xffffffffu, cmax, j), cs = __shfl_sync(0xffffffffu, csum, j);
    const float sv = S[t * SP + j];
    const float cf = (expf(sv - cm) / cs) * (expf(sv - rmax) / rsum);
    if (lane < WW) {
      if (fine_matrix) fine_matrix[m * WW * WW + t * WW + j] = cf;
      if (cf > bv) { bv = cf; bi = t * WW + j; }           // ascending j: first index on ties inside the row
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    sel[m] = bv > thr ? 1 : 0;
    fi[m] = bi / WW; fj[m] = bi % WW;
    fconf[m] = bv;
  }
}

// Ordered compaction of kept fine matches (single CTA, 1024 threads, running base).
//   mkpts_f = ([cell % W - W/2, cell / W - W/2] + mkpts_c / coarse_scale * c2f_scale) * fine_scale
__global__ void __launch_bounds__(1024)
compact_fine_kernel(const int* __restrict__ sel, const int* __restrict__ fi, const int* __restrict__ fj,
                    const float* __restrict__ fconf, const float* __restrict__ k0c, const float* __restrict__ k1c,
                    const int64_t* __restrict__ b_ids, int64_t m, int window, float coarse_scale, float c2f, float fine_scale,
                    float* __restrict__ k0f, float* __restrict__ k1f, float* __restrict__ mconf,
                    int64_t* __restrict__ m_bids, int* __restrict__ total) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = window / 2;
  int base = 0;
  for (int64_t i0 = 0; i0 < m; i0 += 1024) {
    const int64_t i = i0 + threadIdx.x;
    const bool flag = i < m && sel[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const int v = warp_tot[w]; if (w < warp) off += v; tot += v; }
    __syncthreads();
    if (flag) {
      const int pos = base + off + __popc(bal & ((1u << lane) - 1u));
      const int ci = fi[i], cj = fj[i];
      const float c0x = k0c[2 * i] / coarse_scale * c2f, c0y = k0c[2 * i + 1] / coarse_scale * c2f;
      const float c1x = k1c[2 * i] / coarse_scale * c2f, c1y = k1c[2 * i + 1] / coarse_scale * c2f;
      k0f[2 * pos] = ((float)(ci % window - half) + c0x) * fine_scale;
      k0f[2 * pos + 1] = ((float)(ci / window - half) + c0y) * fine_scale;
      k1f[2 * pos] = ((float)(cj % window - half) + c1x) * fine_scale;
      k1f[2 * pos + 1] = ((float)(cj / window - half) + c1y) * fine_scale;
      mconf[pos] = fconf[i];
      m_bids[pos] = b_ids[i];
    }
    base += tot;
  }
  if (threadIdx.x == 0) *total = base;
}

}  // namespace gf

using namespace gf;
#define STREAM ((cudaStream_t)stream)

extern "C" int gf_fine_gather(const float* fine_nhwc, int hf, int wf, int c, const int64_t* b_ids,
                              const int64_t* tok_ids, int64_t m, int wc, int stride, int window, float* out,
                              gf_stream_t stream) {
  if (m < 0 || c <= 0 || (c % 4) || window <= 0) return gf_set_error(GF_ERR_ARG, "gf_fine_gather: bad shape");
  if (m == 0) return GF_OK;
  fine_gather_kernel<<<(unsigned)m, 128, 0, STREAM>>>(fine_nhwc, hf, wf, c, b_ids, tok_ids, wc, stride, window, out);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_fine_gather_f16(const void* fine_nhwc, int hf, int wf, int c, const int64_t* b_ids,
                                   const int64_t* tok_ids, int64_t m, int wc, int stride, int window, float* out,
                                   gf_stream_t stream) {
  if (m < 0 || c <= 0 || (c % 4) || window <= 0) return gf_set_error(GF_ERR_ARG, "gf_fine_gather_f16: bad shape");
  if (m == 0) return GF_OK;
  fine_gather_f16_kernel<float><<<(unsigned)m, 128, 0, STREAM>>>((const __half*)fine_nhwc, hf, wf, c, b_ids, tok_ids, wc,
                                                                 stride, window, out);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_fine_gather_f16_f16(const void* fine_nhwc, int hf, int wf, int c, const int64_t* b_ids,
                                       const int64_t* tok_ids, int64_t m, int wc, int stride, int window, void* out,
                                       gf_stream_t stream) {
  if (m < 0 || c <= 0 || (c % 4) || window <= 0) return gf_set_error(GF_ERR_ARG, "gf_fine_gather_f16_f16: bad shape");
  if (m == 0) return GF_OK;
  fine_gather_f16_kernel<__half><<<(unsigned)m, 128, 0, STREAM>>>((const __half*)fine_nhwc, hf, wf, c, b_ids, tok_ids, wc,
                                                                  stride, window, (__half*)out);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_fine_match(const float* f0, const float* f1, int64_t m, int ww, int c, float temperature, float thr,
                             int* sel, int* fi, int* fj, float* fconf, float* fine_matrix, gf_stream_t stream) {
  if (m < 0 || ww <= 0 || ww > 25 || c <= 0 || c > 128) return gf_set_error(GF_ERR_ARG, "gf_fine_match: ww <= 25, c <= 128");
  if (m == 0) return GF_OK;
  if (ww == 25 && c == 128) {
    constexpr int kSmem = 8 * (2 * 25 * 36 + 25 * 27 + 1) * 4;
    GF_SMEM_OPTIN(fine_match_warp_kernel, kSmem);
    fine_match_warp_kernel<<<(unsigned)((m + 7) / 8), 256, kSmem, STREAM>>>(f0, f1, m, temperature, thr, sel, fi, fj, fconf, fine_matrix);
    g_launches++;
    GF_CHECK_LAUNCH();
    return GF_OK;
  }
  fine_match_kernel<<<(unsigned)m, 128, 0, STREAM>>>(f0, f1, ww, c, temperature, thr, sel, fi, fj, fconf, fine_matrix);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_compact_fine(const int* sel, const int* fi, const int* fj, const float* fconf, const float* mkpts0_c,
                               const float* mkpts1_c, const int64_t* b_ids, int64_t m, int window, float coarse_scale,
                               float c2f_scale, float fine_scale, float* mkpts0_f, float* mkpts1_f, float* mconf,
                               int64_t* m_bids, int* total, gf_stream_t stream) {
  if (m < 0 || window <= 0) return gf_set_error(GF_ERR_ARG, "gf_compact_fine: bad shape");
  compact_fine_kernel<<<1, 1024, 0, STREAM>>>(sel, fi, fj, fconf, mkpts0_c, mkpts1_c, b_ids, m, window, coarse_scale,
                                              c2f_scale, fine_scale, mkpts0_f, mkpts1_f, mconf, m_bids, total);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
