// Coarse matching kernels: split-fp16 operand packing, dual-softmax statistics / confidence,
// mutual-nearest-neighbour selection and ordered compaction.
// Reference: model/loftr_src/loftr/utils/coarse_matching.py:110-125 (conf), 161-212 (matches).
// All of these are HBM-bound sweeps over the [N, L, S] matrix (92 MB per 640x480 pair) or over vectors.
#include "common.cuh"

#include <atomic>
#include <cuda_fp16.h>

namespace gf {
extern std::atomic<int64_t> g_launches;

// ---------------------------------------------------------------------------------------------
// x*scale = hi + lo (fp16, fp16).  role 0: [hi | hi | lo]   role 1: [hi | lo | hi]
// ---------------------------------------------------------------------------------------------
__global__ void pack_split_f16_kernel(const float* __restrict__ x, __half* __restrict__ out, int64_t rows, int c,
                                      float scale, int role) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c) return;
  const int64_t r = idx / c;
  const int col = (int)(idx - r * c);
  const float v = x[idx] * scale;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  __half* o = out + r * 3 * c;
  o[col] = hi;
  o[c + col] = role == 0 ? hi : lo;
  o[2 * c + col] = role == 0 ? lo : hi;
}

// ---------------------------------------------------------------------------------------------
// reference similarity (plain FFMA): sim[n,l,s] = sum_c (f0*in_scale)(f1*in_scale) * out_scale
// ---------------------------------------------------------------------------------------------
__global__ void similarity_ref_kernel(const float* __restrict__ f0, const float* __restrict__ f1, float* __restrict__ sim,
                                      int l, int s, int c, float in_scale, float out_scale) {
  __shared__ float a[32][33], b[32][33];
  const int n = blockIdx.z;
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc = 0.f;
  for (int k0 = 0; k0 < c; k0 += 32) {
    const int ra = row0 + ty, rb = col0 + ty;
    a[ty][tx] = (ra < l && k0 + tx < c) ? f0[((int64_t)n * l + ra) * c + k0 + tx] * in_scale : 0.f;
    b[ty][tx] = (rb < s && k0 + tx < c) ? f1[((int64_t)n * s + rb) * c + k0 + tx] * in_scale : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(a[ty][k], b[tx][k], acc);
    __syncthreads();
  }
  if (row0 + ty < l && col0 + tx < s) sim[((int64_t)n * l + row0 + ty) * s + col0 + tx] = acc * out_scale;
}

// ---------------------------------------------------------------------------------------------
// soft-max statistics.  rows: one warp per row.  columns: 32-column strips, 32 row-lanes per strip.
// ---------------------------------------------------------------------------------------------
__global__ void row_stats_kernel(const float* __restrict__ sim, int64_t rows, int s, float* __restrict__ row_max,
                                 float* __restrict__ row_sum) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = sim + row * s;
  // online (max, sum exp) per lane: a single read of the row; lanes are merged afterwards
  float m = -INFINITY, sum = 0.f;
  for (int j = lane; j < s; j += 32) {
    const float v = p[j];
    if (v > m) { sum = sum * expf(m - v) + 1.f; m = v; } else { sum += expf(v - m); }
  }
  const float mm = warp_max(m);
  sum = (m > -INFINITY) ? sum * expf(m - mm) : 0.f;
  sum = warp_sum(sum);
  if (lane == 0) { row_max[row] = mm; row_sum[row] = sum; }
}

// MODE 0: (max, sum exp(x - max)) per column;  MODE 1: max only
template <int MODE>
__global__ void col_stats_kernel(const float* __restrict__ mat, int l, int s, float* __restrict__ col_max,
                                 float* __restrict__ col_sum) {
  __shared__ float sm[32][33], ss[32][33];
  const int n = blockIdx.y;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const float* p = mat + (int64_t)n * l * s;
  float m = -INFINITY, sum = 0.f;
  if (col < s) {
    for (int r = threadIdx.y; r < l; r += 32) {
      const float v = p[(int64_t)r * s + col];
      if (MODE == 0) {
        if (v > m) { sum = sum * expf(m - v) + 1.f; m = v; } else { sum += expf(v - m); }
      } else {
        m = fmaxf(m, v);
      }
    }
  }
  sm[threadIdx.y][threadIdx.x] = m;
  ss[threadIdx.y][threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.y == 0 && col < s) {
    float mm = -INFINITY;
    for (int i = 0; i < 32; ++i) mm = fmaxf(mm, sm[i][threadIdx.x]);
    col_max[(int64_t)n * s + col] = mm;
    if (MODE == 0) {
      float tot = 0.f;
      for (int i = 0; i < 32; ++i) {
        const float mi = sm[i][threadIdx.x];
        if (mi > -INFINITY) tot += ss[i][threadIdx.x] * expf(mi - mm);
      }
      col_sum[(int64_t)n * s + col] = tot;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused statistics: ONE read of sim produces row and column (max, sum exp) partials.
// CTA = 128 rows x 256 columns; thread == column (coalesced 1 KB row segments).  Column statistics are an
// online (max, sum) in the thread's registers; row statistics are reduced across the CTA's 256 columns with
// warp shuffles + shared memory.  Partials: rowp[n][l][cblocks] and colp[n][rblocks][s] as float2 (max, sum).
// ---------------------------------------------------------------------------------------------
constexpr int kStatRows = 128;
constexpr int kStatSub = 32;       // rows staged per shared-memory sub-tile
__global__ void __launch_bounds__(256)
fused_stats_kernel(const float* __restrict__ sim, int l, int s, float2* __restrict__ rowp, float2* __restrict__ colp) {
  __shared__ float tile[kStatSub][256 + 4];
  const int n = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
  const int col = cb * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = rb * kStatRows;
  const int rows = min(kStatRows, l - r0);
  const bool ok = col < s;
  const float* p = sim + ((int64_t)n * l + r0) * s + col;
  float cm = -INFINITY, cs = 0.f;
  for (int rs = 0; rs < rows; rs += kStatSub) {
    const int sub = min(kStatSub, rows - rs);
    // column pass: thread == column, online (max, sum); the values are parked in shared memory for the row pass
    float v[kStatSub];
#pragma unroll
    for (int r = 0; r < kStatSub; ++r) v[r] = (ok && r < sub) ? p[(int64_t)(rs + r) * s] : -INFINITY;   // 32 loads in flight
#pragma unroll
    for (int r = 0; r < kStatSub; ++r) {
      tile[r][threadIdx.x] = v[r];
      if (v[r] > cm) { cs = cs * __expf(cm - v[r]) + 1.f; cm = v[r]; } else if (v[r] > -INFINITY) { cs += __expf(v[r] - cm); }
    }
    __syncthreads();
    // row pass: each warp owns 4 of the 32 rows; lane reads 8 values, then one shuffle tree per row
#pragma unroll
    for (int q = 0; q < kStatSub / 8; ++q) {
      const int r = warp * (kStatSub / 8) + q;
      float x[8];
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) { x[k] = tile[r][lane + 32 * k]; m = fmaxf(m, x[k]); }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += (x[k] > -INFINITY) ? __expf(x[k] - m) : 0.f;
      sum = warp_sum(sum);
      if (lane == 0 && r < sub) rowp[((int64_t)n * l + r0 + rs + r) * gridDim.x + cb] = make_float2(m, sum);
    }
    __syncthreads();
  }
  if (ok) colp[((int64_t)n * gridDim.y + rb) * s + col] = make_float2(cm, cs);
}

// merge `parts` (max, sum) partials per output element; partial k of element e is at in[e*estride + k*kstride]
__global__ void merge_stats_kernel(const float2* __restrict__ in, int64_t elems, int parts, int64_t estride,
                                   int64_t kstride, int64_t group, int64_t gstride, float* __restrict__ out_max,
                                   float* __restrict__ out_sum) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  // element e = (g, i) with i < group: base = g*gstride + i*estride
  const int64_t g = e / group, i = e - g * group;
  const float2* p = in + g * gstride + i * estride;
  float m = -INFINITY;
  for (int k = 0; k < parts; ++k) m = fmaxf(m, p[k * kstride].x);
  float sum = 0.f;
  for (int k = 0; k < parts; ++k) { const float2 v = p[k * kstride]; sum += v.y * __expf(v.x - m); }
  out_max[e] = m;
  out_sum[e] = sum;
}

// conf = softmax(sim, dim=1) * softmax(sim, dim=2) in place, plus max of conf per row and per column, in ONE
// sweep.  CTA = 256 columns x 32 rows; thread == column (coalesced), loops the 32 rows.  Row maxima are merged
// with warp shuffles + shared atomics, column maxima with one global atomicMax per (column, 32-row strip)
// (conf >= 0, so the float bit pattern orders like an unsigned integer).
__global__ void __launch_bounds__(256, 4)
conf_kernel(float* __restrict__ sim, int l, int s, const float* __restrict__ row_max, const float* __restrict__ row_sum,
            const float* __restrict__ col_max, const float* __restrict__ col_sum, unsigned* __restrict__ conf_row_max,
            unsigned* __restrict__ conf_col_max) {
  __shared__ float tile[32][256 + 4];
  __shared__ float rm[32], rs[32];
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * 32;
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = min(32, l - r0);
  if (threadIdx.x < 32) {
    const int r = r0 + threadIdx.x;
    rm[threadIdx.x] = r < l ? row_max[(int64_t)n * l + r] : 0.f;
    rs[threadIdx.x] = r < l ? 1.f / row_sum[(int64_t)n * l + r] : 1.f;      // reciprocal row sum
  }
  __syncthreads();
  const bool ok = col < s;
  const float cm = ok ? col_max[(int64_t)n * s + col] : 0.f;
  const float ics = ok ? 1.f / col_sum[(int64_t)n * s + col] : 1.f;
  float* p = sim + ((int64_t)n * l + r0) * s + col;
  float cbest = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = (ok && half * 16 + r < rows) ? p[(int64_t)(half * 16 + r) * s] : -INFINITY;   // 16 loads in flight
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int rr = half * 16 + r;
      float c = 0.f;
      if (ok && rr < rows) {
        c = (__expf(v[r] - cm) * ics) * (__expf(v[r] - rm[rr]) * rs[rr]);
        p[(int64_t)rr * s] = c;
        cbest = fmaxf(cbest, c);
      }
      tile[rr][threadIdx.x] = c;
    }
  }
  if (ok) atomicMax(&conf_col_max[(int64_t)n * s + col], __float_as_uint(cbest));
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {                       // warp w reduces rows 4w .. 4w+3 over the CTA's 256 columns
    const int r = warp * 4 + q;
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(m, tile[r][lane + 32 * k]);
    m = warp_max(m);
    if (lane == 0 && r < rows) atomicMax(&conf_row_max[(int64_t)n * l + r0 + r], __float_as_uint(m));
  }
}

__global__ void row_max_kernel(const float* __restrict__ mat, int64_t rows, int s, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = mat + row * s;
  float m = -INFINITY;
  for (int j = lane; j < s; j += 32) m = fmaxf(m, p[j]);
  m = warp_max(m);
  if (lane == 0) out[row] = m;
}

// ---------------------------------------------------------------------------------------------
// match_j[row] = first j with conf>thr && conf==rowmax && conf==colmax && border ok, else -1
// ---------------------------------------------------------------------------------------------
__global__ void mnn_select_kernel(const float* __restrict__ conf, int64_t rows, int l, int s, float thr, int border,
                                  int h0c, int w0c, int h1c, int w1c, const float* __restrict__ conf_row_max,
                                  const float* __restrict__ conf_col_max, int* __restrict__ match_j,
                                  float* __restrict__ match_conf) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t n = row / l;
  const int i = (int)(row - n * l);
  const float rm = conf_row_max[row];
  bool ok = rm > thr;
  if (border > 0) {
    const int r0 = i / w0c, c0 = i % w0c;
    ok = ok && r0 >= border && r0 < h0c - border && c0 >= border && c0 < w0c - border;
  }
  int found = -1;
  if (ok) {
    const float* p = conf + row * s;
    const float* cm = conf_col_max + n * s;
    for (int j0 = 0; j0 < s && found < 0; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < s) {
        const float v = p[j];
        hit = (v == rm) && (v == cm[j]);
        if (hit && border > 0) {
          const int r1 = j / w1c, c1 = j % w1c;
          hit = r1 >= border && r1 < h1c - border && c1 >= border && c1 < w1c - border;
        }
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (b) found = j0 + __ffs(b) - 1;
    }
  }
  if (lane == 0) {
    match_j[row] = found;
    match_conf[row] = found >= 0 ? rm : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// ordered compaction.  Pass 1: per-sample counts.  Pass 2: scatter in (b, i) order.
// ---------------------------------------------------------------------------------------------
__global__ void count_matches_kernel(const int* __restrict__ match_j, int l, int* __restrict__ counts) {
  __shared__ int acc;
  if (threadIdx.x == 0) acc = 0;
  __syncthreads();
  const int n = blockIdx.x;
  int c = 0;
  for (int i = threadIdx.x; i < l; i += blockDim.x) c += match_j[(int64_t)n * l + i] >= 0;
  c = (int)warp_sum((float)c);   // counts < 2^24: exact in fp32
  if ((threadIdx.x & 31) == 0) atomicAdd(&acc, c);
  __syncthreads();
  if (threadIdx.x == 0) counts[n] = acc;
}

// block-wide exclusive position of a flag among 1024 threads, in thread order
__device__ __forceinline__ int block_rank_1024(bool flag, int* warp_tot, int& block_total) {
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_tot[warp] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < 32; ++w) {
    const int v = warp_tot[w];
    if (w < warp) base += v;
    tot += v;
  }
  __syncthreads();
  block_total = tot;
  return base + __popc(b & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(1024)
compact_coarse_kernel(const int* __restrict__ match_j, const float* __restrict__ match_conf, int n_samples, int l,
                      int w0c, int w1c, float scale, int64_t* __restrict__ b_ids, int64_t* __restrict__ i_ids,
                      int64_t* __restrict__ j_ids, float* __restrict__ mconf, float* __restrict__ k0,
                      float* __restrict__ k1, const int* __restrict__ counts, int* __restrict__ total,
                      int64_t capacity) {
  __shared__ int warp_tot[32];
  const int n = blockIdx.x;
  int base = 0;
  for (int b = 0; b < n; ++b) base += counts[b];
  if (n == 0 && threadIdx.x == 0) {
    int t = 0;
    for (int b = 0; b < n_samples; ++b) t += counts[b];
    *total = t;
  }
  for (int i0 = 0; i0 < l; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    int j = -1;
    if (i < l) j = match_j[(int64_t)n * l + i];
    int tot;
    const int pos = base + block_rank_1024(j >= 0, warp_tot, tot);
    if (j >= 0 && pos < capacity) {
      b_ids[pos] = n; i_ids[pos] = i; j_ids[pos] = j;
      mconf[pos] = match_conf[(int64_t)n * l + i];
      k0[2 * pos] = (float)(i % w0c) * scale; k0[2 * pos + 1] = (float)(i / w0c) * scale;
      k1[2 * pos] = (float)(j % w1c) * scale; k1[2 * pos + 1] = (float)(j / w1c) * scale;
    }
    base += tot;
  }
}

}  // namespace gf

using namespace gf;
#define STREAM ((cudaStream_t)stream)

extern "C" int gf_pack_split_f16(const float* x, void* out_f16, int64_t rows, int c, float scale, int role,
                                 gf_stream_t stream) {
  if (rows < 0 || c <= 0) return gf_set_error(GF_ERR_ARG, "gf_pack_split_f16: bad shape");
  if (rows == 0) return GF_OK;
  const int64_t total = rows * c;
  pack_split_f16_kernel<<<gf_cdiv(total, 256), 256, 0, STREAM>>>(x, (__half*)out_f16, rows, c, scale, role);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_similarity_ref(const float* f0, const float* f1, float* sim, int n, int l, int s, int c,
                                 float in_scale, float out_scale, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0 || c <= 0) return gf_set_error(GF_ERR_ARG, "gf_similarity_ref: bad shape");
  dim3 grid(gf_cdiv(s, 32), gf_cdiv(l, 32), n);
  similarity_ref_kernel<<<grid, dim3(32, 32), 0, STREAM>>>(f0, f1, sim, l, s, c, in_scale, out_scale);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int64_t gf_dual_softmax_workspace_floats(int n, int l, int s) {
  return 2 * ((int64_t)n * l * gf_cdiv(s, 256) + (int64_t)n * gf_cdiv(l, kStatRows) * s);
}

extern "C" int gf_dual_softmax_stats(const float* sim, int n, int l, int s, float* row_max, float* row_sum,
                                     float* col_max, float* col_sum, float* workspace, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0) return gf_set_error(GF_ERR_ARG, "gf_dual_softmax_stats: bad shape");
  const int64_t rows = (int64_t)n * l;
  if (workspace == nullptr) {       // two-sweep fallback (no workspace supplied)
    row_stats_kernel<<<gf_cdiv(rows, 8), 256, 0, STREAM>>>(sim, rows, s, row_max, row_sum);
    col_stats_kernel<0><<<dim3(gf_cdiv(s, 32), n), dim3(32, 32), 0, STREAM>>>(sim, l, s, col_max, col_sum);
    g_launches += 2;
  } else {
    const int cblocks = gf_cdiv(s, 256), rblocks = gf_cdiv(l, kStatRows);
    float2* rowp = reinterpret_cast<float2*>(workspace);
    float2* colp = rowp + (int64_t)n * l * cblocks;
    fused_stats_kernel<<<dim3(cblocks, rblocks, n), 256, 0, STREAM>>>(sim, l, s, rowp, colp);
    // rows: element (g = 0, i = row): partials contiguous (estride = cblocks, kstride = 1)
    merge_stats_kernel<<<gf_cdiv(rows, 256), 256, 0, STREAM>>>(rowp, rows, cblocks, cblocks, 1, rows, 0, row_max, row_sum);
    // columns: element (g = sample, i = col): partials strided by s (kstride = s), sample stride rblocks*s
    merge_stats_kernel<<<gf_cdiv((int64_t)n * s, 256), 256, 0, STREAM>>>(colp, (int64_t)n * s, rblocks, 1, s, s,
                                                                        (int64_t)rblocks * s, col_max, col_sum);
    g_launches += 3;
  }
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_dual_softmax_conf(float* sim_conf, int n, int l, int s, const float* row_max, const float* row_sum,
                                    const float* col_max, const float* col_sum, float* conf_row_max,
                                    float* conf_col_max, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0) return gf_set_error(GF_ERR_ARG, "gf_dual_softmax_conf: bad shape");
  cudaMemsetAsync(conf_row_max, 0, sizeof(float) * (size_t)n * l, STREAM);
  cudaMemsetAsync(conf_col_max, 0, sizeof(float) * (size_t)n * s, STREAM);
  conf_kernel<<<dim3(gf_cdiv(s, 256), gf_cdiv(l, 32), n), 256, 0, STREAM>>>(
      sim_conf, l, s, row_max, row_sum, col_max, col_sum, (unsigned*)conf_row_max, (unsigned*)conf_col_max);
  g_launches += 1;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_conf_row_col_max(const float* conf, int n, int l, int s, float* conf_row_max, float* conf_col_max,
                                   gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0) return gf_set_error(GF_ERR_ARG, "gf_conf_row_col_max: bad shape");
  const int64_t rows = (int64_t)n * l;
  row_max_kernel<<<gf_cdiv(rows, 8), 256, 0, STREAM>>>(conf, rows, s, conf_row_max);
  col_stats_kernel<1><<<dim3(gf_cdiv(s, 32), n), dim3(32, 32), 0, STREAM>>>(conf, l, s, conf_col_max, nullptr);
  g_launches += 2;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_mnn_select(const float* conf, int n, int l, int s, float thr, int border, int h0c, int w0c, int h1c,
                             int w1c, const float* conf_row_max, const float* conf_col_max, int* match_j,
                             float* match_conf, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0 || h0c * w0c != l || h1c * w1c != s)
    return gf_set_error(GF_ERR_ARG, "gf_mnn_select: bad shape");
  const int64_t rows = (int64_t)n * l;
  mnn_select_kernel<<<gf_cdiv(rows, 8), 256, 0, STREAM>>>(conf, rows, l, s, thr, border, h0c, w0c, h1c, w1c,
                                                         conf_row_max, conf_col_max, match_j, match_conf);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

extern "C" int gf_compact_coarse(const int* match_j, const float* match_conf, int n, int l, int w0c, int w1c,
                                 float scale, int64_t* b_ids, int64_t* i_ids, int64_t* j_ids, float* mconf,
                                 float* mkpts0_c, float* mkpts1_c, int* counts, int* total, int64_t capacity,
                                 gf_stream_t stream) {
  if (n <= 0 || l <= 0 || w0c <= 0 || w1c <= 0) return gf_set_error(GF_ERR_ARG, "gf_compact_coarse: bad shape");
  count_matches_kernel<<<n, 256, 0, STREAM>>>(match_j, l, counts);
  compact_coarse_kernel<<<n, 1024, 0, STREAM>>>(match_j, match_conf, n, l, w0c, w1c, scale, b_ids, i_ids, j_ids, mconf,
                                                mkpts0_c, mkpts1_c, counts, total, capacity);
  g_launches += 2;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
