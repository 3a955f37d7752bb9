// GPU RANSAC homography: optional replacement of the host cv2.findHomography call inside
// the forward (model/geo_module.py:45-52).  NOT bit-identical to OpenCV (different sampling, DLT refit instead of
// LM), so it is gated behind GeoFormer.ransac = "gpu"; the default stays cv2 (parity by construction).
//
//   kernel 1: one warp per hypothesis — 4 distinct random matches -> exact 4-point homography (square->quad
//             composition, fp64) -> inlier count over all matches of the sample (reprojection error < thr)
//   kernel 2: one CTA per sample — best hypothesis (first maximum), inlier set, normalised DLT least-squares refit
//             over the inliers (9x9 normal matrix, cyclic Jacobi eigen-solver, fp64), inliers re-evaluated (2 rounds);
//             emits H, H^-1 (fp32 row-major), has_h and per-match inlier flags
//   kernel 3/4: anchor (inlier-token) lists of both images in ascending token order, as the reference's boolean maps.
// Semantics kept from the reference: no homography when a sample has <= 8 matches (or the fit degenerates); then the
// anchors are ALL first-pass matches and the cross layers are skipped for that sample (geo_module.py:76-88).
#include "common.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// projective map of the unit square (0,0),(1,0),(1,1),(0,1) onto the quad p[0..3] (Heckbert); returns false if degenerate
__device__ bool square_to_quad(const double* px, const double* py, double* m) {
  const double dx1 = px[1] - px[2], dx2 = px[3] - px[2], sx = px[0] - px[1] + px[2] - px[3];
  const double dy1 = py[1] - py[2], dy2 = py[3] - py[2], sy = py[0] - py[1] + py[2] - py[3];
  const double den = dx1 * dy2 - dx2 * dy1;
  if (fabs(den) < 1e-9) return false;
  const double g = (sx * dy2 - dx2 * sy) / den, h = (dx1 * sy - sx * dy1) / den;
  m[0] = px[1] - px[0] + g * px[1]; m[1] = px[3] - px[0] + h * px[3]; m[2] = px[0];
  m[3] = py[1] - py[0] + g * py[1]; m[4] = py[3] - py[0] + h * py[3]; m[5] = py[0];
  m[6] = g; m[7] = h; m[8] = 1.0;
  return true;
}

__device__ bool invert3(const double* a, double* o) {
  const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
  if (fabs(det) < 1e-18) return false;
  const double id = 1.0 / det;
  o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return true;
}

__device__ __forceinline__ void mul3(const double* a, const double* b, double* o) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}

__device__ __forceinline__ bool is_inlier(const double* H, float x0, float y0, float x1, float y1, float thr2) {
  const double w = H[6] * x0 + H[7] * y0 + H[8];
  if (fabs(w) < 1e-12) return false;
  const double ex = (H[0] * x0 + H[1] * y0 + H[2]) / w - x1, ey = (H[3] * x0 + H[4] * y0 + H[5]) / w - y1;
  return (float)(ex * ex + ey * ey) <= thr2;   // OpenCV counts err <= thr^2
}

// ---------------------------------------------------------------------------------------------
// kernel 1: hypotheses.  grid (hyps / 8, n), 256 threads (8 warps = 8 hypotheses)
// ---------------------------------------------------------------------------------------------
__global__ void ransac_hypotheses_kernel(const float* __restrict__ k0, const float* __restrict__ k1,
                                         const int* __restrict__ counts, int n, int hyps, float thr2, uint32_t seed,
                                         int* __restrict__ score, double* __restrict__ hmat) {
  const int b = blockIdx.y;
  const int hyp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  int off = 0;
  for (int i = 0; i < b; ++i) off += counts[i];
  const int cnt = counts[b];
  if (hyp >= hyps) return;
  int* sc = score + (int64_t)b * hyps + hyp;
  if (cnt <= 8) { if (lane == 0) *sc = -1; return; }
  // 4 distinct matches (deterministic hash RNG; duplicates are re-drawn a few times)
  int idx[4];
  uint32_t st = hash32(seed ^ (uint32_t)(b * 9781 + 1)) ^ hash32((uint32_t)hyp * 2654435761U + 12345U);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    for (int tries = 0; tries < 8; ++tries) {
      st = hash32(st + 0x9e3779b9U);
      idx[i] = (int)(st % (uint32_t)cnt);
      bool dup = false;
      for (int j = 0; j < i; ++j) dup |= idx[j] == idx[i];
      if (!dup) break;
    }
  }
  double H[9];
  bool ok = true;
  {
    double ax[4], ay[4], bx[4], by[4], A[9], B[9], Ai[9];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ax[i] = k0[2 * (off + idx[i])]; ay[i] = k0[2 * (off + idx[i]) + 1];
      bx[i] = k1[2 * (off + idx[i])]; by[i] = k1[2 * (off + idx[i]) + 1];
    }
    ok = square_to_quad(ax, ay, A) && square_to_quad(bx, by, B) && invert3(A, Ai);
    if (ok) {
      mul3(B, Ai, H);
      ok = fabs(H[8]) > 1e-12;
      if (ok) { const double s = 1.0 / H[8]; for (int i = 0; i < 9; ++i) H[i] *= s; }
    }
  }
  int c = 0;
  if (ok) {
    for (int i = lane; i < cnt; i += 32)
      c += is_inlier(H, k0[2 * (off + i)], k0[2 * (off + i) + 1], k1[2 * (off + i)], k1[2 * (off + i) + 1], thr2) ? 1 : 0;
  }
  c = (int)warp_sum((float)c);
  if (lane == 0) {
    *sc = ok ? c : -1;
    double* o = hmat + ((int64_t)b * hyps + hyp) * 9;
    for (int i = 0; i < 9; ++i) o[i] = ok ? H[i] : 0.0;
  }
}

// smallest eigenvector of a symmetric 9x9 matrix (cyclic Jacobi), single thread
__device__ void smallest_eigvec9(double* A, double* vec) {
  double V[81];
  for (int i = 0; i < 81; ++i) V[i] = (i % 10 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double offn = 0.0, diag = 0.0;
    for (int p = 0; p < 9; ++p) {
      diag += A[p * 9 + p] * A[p * 9 + p];
      for (int q = p + 1; q < 9; ++q) offn += A[p * 9 + q] * A[p * 9 + q];
    }
    if (offn <= 1e-28 * diag) break;
    for (int p = 0; p < 9; ++p) {
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p * 9 + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 9 + q] - A[p * 9 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) {
          const double akp = A[k * 9 + p], akq = A[k * 9 + q];
          A[k * 9 + p] = c * akp - s * akq; A[k * 9 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = A[p * 9 + k], aqk = A[q * 9 + k];
          A[p * 9 + k] = c * apk - s * aqk; A[q * 9 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k * 9 + p], vkq = V[k * 9 + q];
          V[k * 9 + p] = c * vkp - s * vkq; V[k * 9 + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int best = 0;
  for (int i = 1; i < 9; ++i) if (A[i * 9 + i] < A[best * 9 + best]) best = i;
  for (int k = 0; k < 9; ++k) vec[k] = V[k * 9 + best];
}

// ---------------------------------------------------------------------------------------------
// kernel 2: select + refit.  one CTA (256 threads) per sample
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ransac_select_refit_kernel(const float* __restrict__ k0, const float* __restrict__ k1, const int* __restrict__ counts,
                           int hyps, float thr2, const int* __restrict__ score, const double* __restrict__ hmat,
                           float* __restrict__ h_out, float* __restrict__ hinv_out, int* __restrict__ has_h,
                           int* __restrict__ inlier) {
  __shared__ int s_best[256], s_idx[256];
  __shared__ double s_acc[256];
  __shared__ double Hs[9];
  __shared__ double norm[8];           // cx0, cy0, s0, cx1, cy1, s1
  __shared__ double ATA[81];
  __shared__ int s_ok, s_cnt;
  const int b = blockIdx.x, tid = threadIdx.x;
  int off = 0;
  for (int i = 0; i < b; ++i) off += counts[i];
  const int cnt = counts[b];
  // ---- arg-max over hypotheses (first maximum)
  int bs = -1, bi = 0x7fffffff;
  for (int h = tid; h < hyps; h += 256) {
    const int sc = score[(int64_t)b * hyps + h];
    if (sc > bs) { bs = sc; bi = h; }
  }
  s_best[tid] = bs; s_idx[tid] = bi;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) {
      const int ob = s_best[tid + st], oi = s_idx[tid + st];
      if (ob > s_best[tid] || (ob == s_best[tid] && oi < s_idx[tid])) { s_best[tid] = ob; s_idx[tid] = oi; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    s_ok = (cnt > 8 && s_best[0] >= 4) ? 1 : 0;
    if (s_ok) for (int i = 0; i < 9; ++i) Hs[i] = hmat[((int64_t)b * hyps + s_idx[0]) * 9 + i];
  }
  __syncthreads();
  // ---- two rounds of: inliers of Hs -> normalised DLT least squares -> Hs
  for (int round = 0; round < 2 && s_ok; ++round) {
    // centroids / scales of the inliers (Hartley normalisation)
    double part[6] = {0, 0, 0, 0, 0, 0};
    int c = 0;
    for (int i = tid; i < cnt; i += 256) {
      const float x0 = k0[2 * (off + i)], y0 = k0[2 * (off + i) + 1], x1 = k1[2 * (off + i)], y1 = k1[2 * (off + i) + 1];
      if (is_inlier(Hs, x0, y0, x1, y1, thr2)) { part[0] += x0; part[1] += y0; part[3] += x1; part[4] += y1; ++c; }
    }
    for (int q = 0; q < 6; ++q) {
      if (q == 2 || q == 5) continue;
      s_acc[tid] = part[q];
      __syncthreads();
      for (int st = 128; st > 0; st >>= 1) { if (tid < st) s_acc[tid] += s_acc[tid + st]; __syncthreads(); }
      if (tid == 0) norm[q] = s_acc[0];
      __syncthreads();
    }
    s_best[tid] = c;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) s_best[tid] += s_best[tid + st]; __syncthreads(); }
    if (tid == 0) {
      s_cnt = s_best[0];
      if (s_cnt < 4) s_ok = 0;
      else { norm[0] /= s_cnt; norm[1] /= s_cnt; norm[3] /= s_cnt; norm[4] /= s_cnt; }
    }
    __syncthreads();
    if (!s_ok) break;
    double d0 = 0, d1 = 0;
    for (int i = tid; i < cnt; i += 256) {
      const float x0 = k0[2 * (off + i)], y0 = k0[2 * (off + i) + 1], x1 = k1[2 * (off + i)], y1 = k1[2 * (off + i) + 1];
      if (is_inlier(Hs, x0, y0, x1, y1, thr2)) {
        d0 += sqrt((x0 - norm[0]) * (x0 - norm[0]) + (y0 - norm[1]) * (y0 - norm[1]));
        d1 += sqrt((x1 - norm[3]) * (x1 - norm[3]) + (y1 - norm[4]) * (y1 - norm[4]));
      }
    }
    s_acc[tid] = d0; __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) s_acc[tid] += s_acc[tid + st]; __syncthreads(); }
    if (tid == 0) norm[2] = s_acc[0] > 0 ? 1.4142135623730951 * s_cnt / s_acc[0] : 1.0;
    __syncthreads();
    s_acc[tid] = d1; __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if (tid < st) s_acc[tid] += s_acc[tid + st]; __syncthreads(); }
    if (tid == 0) norm[5] = s_acc[0] > 0 ? 1.4142135623730951 * s_cnt / s_acc[0] : 1.0;
    __syncthreads();
    // normal matrix A^T A of the 2 DLT rows per inlier: rows (-x,-y,-1,0,0,0,ux,uy,u) and (0,0,0,-x,-y,-1,vx,vy,v)
    double acc[45];
    for (int q = 0; q < 45; ++q) acc[q] = 0.0;
    for (int i = tid; i < cnt; i += 256) {
      const float fx0 = k0[2 * (off + i)], fy0 = k0[2 * (off + i) + 1], fx1 = k1[2 * (off + i)], fy1 = k1[2 * (off + i) + 1];
      if (!is_inlier(Hs, fx0, fy0, fx1, fy1, thr2)) continue;
      const double x = (fx0 - norm[0]) * norm[2], y = (fy0 - norm[1]) * norm[2];
      const double u = (fx1 - norm[3]) * norm[5], v = (fy1 - norm[4]) * norm[5];
      const double r1[9] = {-x, -y, -1, 0, 0, 0, u * x, u * y, u}, r2[9] = {0, 0, 0, -x, -y, -1, v * x, v * y, v};
      int q = 0;
      for (int p = 0; p < 9; ++p) for (int r = p; r < 9; ++r) acc[q++] += r1[p] * r1[r] + r2[p] * r2[r];
    }
    for (int q = 0, p = 0, r = 0; q < 45; ++q) {
      s_acc[tid] = acc[q];
      __syncthreads();
      for (int st = 128; st > 0; st >>= 1) { if (tid < st) s_acc[tid] += s_acc[tid + st]; __syncthreads(); }
      if (tid == 0) { ATA[p * 9 + r] = s_acc[0]; ATA[r * 9 + p] = s_acc[0]; }
      __syncthreads();
      if (++r == 9) { ++p; r = p; }
    }
    if (tid == 0) {
      double hv[9], Hn[9], T1i[9], T0[9], tmp[9];
      smallest_eigvec9(ATA, hv);
      for (int i = 0; i < 9; ++i) Hn[i] = hv[i];
      // denormalise: H = T1^-1 Hn T0,  T = [[s,0,-s*cx],[0,s,-s*cy],[0,0,1]]
      T0[0] = norm[2]; T0[1] = 0; T0[2] = -norm[2] * norm[0]; T0[3] = 0; T0[4] = norm[2]; T0[5] = -norm[2] * norm[1]; T0[6] = 0; T0[7] = 0; T0[8] = 1;
      T1i[0] = 1.0 / norm[5]; T1i[1] = 0; T1i[2] = norm[3]; T1i[3] = 0; T1i[4] = 1.0 / norm[5]; T1i[5] = norm[4]; T1i[6] = 0; T1i[7] = 0; T1i[8] = 1;
      mul3(Hn, T0, tmp);
      mul3(T1i, tmp, Hn);
      if (fabs(Hn[8]) > 1e-12) { const double s = 1.0 / Hn[8]; for (int i = 0; i < 9; ++i) Hs[i] = Hn[i] * s; }
      else s_ok = 0;
    }
    __syncthreads();
  }
  // ---- outputs
  int c = 0;
  for (int i = tid; i < cnt; i += 256) {
    const bool in = s_ok && is_inlier(Hs, k0[2 * (off + i)], k0[2 * (off + i) + 1], k1[2 * (off + i)], k1[2 * (off + i) + 1], thr2);
    inlier[off + i] = in ? 1 : 0;
    c += in;
  }
  (void)c;
  if (tid == 0) {
    double Hi[9];
    const bool ok = s_ok && invert3(Hs, Hi);
    has_h[b] = ok ? 1 : 0;
    for (int i = 0; i < 9; ++i) { h_out[b * 9 + i] = ok ? (float)Hs[i] : 0.f; hinv_out[b * 9 + i] = ok ? (float)Hi[i] : 0.f; }
  }
}

// ---------------------------------------------------------------------------------------------
// kernels 3/4: anchor maps and ordered anchor lists (ascending token order) for both images
// ---------------------------------------------------------------------------------------------
__global__ void anchor_mark_kernel(const float* __restrict__ k0, const float* __restrict__ k1, const int64_t* __restrict__ b_ids,
                                   const int* __restrict__ inlier, const int* __restrict__ has_h, int64_t m, int scale,
                                   int l0, int w0c, int l1, int w1c, int* __restrict__ map0, int* __restrict__ map1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int b = (int)b_ids[i];
  if (has_h[b] && !inlier[i]) return;          // without a homography every first-pass match is an anchor
  const int t0 = ((int)k0[2 * i + 1] / scale) * w0c + (int)k0[2 * i] / scale;
  const int t1 = ((int)k1[2 * i + 1] / scale) * w1c + (int)k1[2 * i] / scale;
  map0[(int64_t)b * l0 + t0] = 1;
  map1[(int64_t)b * l1 + t1] = 1;
}

__global__ void __launch_bounds__(1024)
anchor_list_kernel(const int* __restrict__ map, int l, int cap, int* __restrict__ idx, int* __restrict__ cnt) {
  __shared__ int warp_tot[32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int i0 = 0; i0 < l; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool flag = i < l && map[(int64_t)b * l + i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const int v = warp_tot[w]; if (w < warp) off += v; tot += v; }
    __syncthreads();
    if (flag) {
      const int pos = base + off + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) idx[(int64_t)b * cap + pos] = i;
    }
    base += tot;
  }
  if (threadIdx.x == 0) cnt[b] = min(base, cap);
}

}  // namespace gf

using namespace gf;

extern "C" int64_t gf_ransac_workspace_bytes(int n, int hyps) { return (int64_t)n * hyps * (4 + 72) + 256; }

// k0/k1 [M,2] fp32 first-pass coarse matches grouped by sample (counts[n], device), b_ids [M].
// Outputs (device): hmat[n,9], hinv[n,9] fp32 row-major, has_h[n], inlier[M],
// anchor_idx{0,1}[n, cap] + anchor_cnt{0,1}[n] (cap >= max(l0, l1) is always sufficient).
extern "C" int gf_ransac_homography(const float* k0, const float* k1, const int64_t* b_ids, const int* counts, int64_t m,
                                    int n, int hyps, float thr, unsigned seed, int scale, int l0, int w0c, int l1, int w1c,
                                    void* workspace, float* hmat, float* hinv, int* has_h, int* inlier, int* map0,
                                    int* map1, int* anchor_idx0, int* anchor_cnt0, int* anchor_idx1, int* anchor_cnt1,
                                    int cap, gf_stream_t stream) {
  if (n <= 0 || hyps <= 0 || (hyps % 8) || m < 0 || workspace == nullptr) return gf_set_error(GF_ERR_ARG, "gf_ransac_homography: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  double* hm = reinterpret_cast<double*>(workspace);
  int* score = reinterpret_cast<int*>(hm + (int64_t)n * hyps * 9);
  ransac_hypotheses_kernel<<<dim3(hyps / 8, n), 256, 0, st>>>(k0, k1, counts, n, hyps, thr * thr, seed, score, hm);
  ransac_select_refit_kernel<<<n, 256, 0, st>>>(k0, k1, counts, hyps, thr * thr, score, hm, hmat, hinv, has_h, inlier);
  cudaMemsetAsync(map0, 0, sizeof(int) * (size_t)n * l0, st);
  cudaMemsetAsync(map1, 0, sizeof(int) * (size_t)n * l1, st);
  if (m > 0) anchor_mark_kernel<<<gf_cdiv(m, 256), 256, 0, st>>>(k0, k1, b_ids, inlier, has_h, m, scale, l0, w0c, l1, w1c, map0, map1);
  anchor_list_kernel<<<n, 1024, 0, st>>>(map0, l0, cap, anchor_idx0, anchor_cnt0);
  anchor_list_kernel<<<n, 1024, 0, st>>>(map1, l1, cap, anchor_idx1, anchor_cnt1);
  g_launches += 5;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
