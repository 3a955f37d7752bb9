// Fused dual-softmax coarse matching (utils/coarse_matching.py:110-125, 161-188) without materialising the
// L x S similarity / confidence matrix.  Two tcgen05 passes over the same split-fp16 operands (K = 3C):
//
//   pass 0 (stats): S tile in TMEM -> per-row online (max, sum 2^x) kept in the epilogue thread (thread == row) and
//                   per-column (max, sum) of each warp's 32 x 32 sub-block via a swizzled smem transpose;
//                   partials rowp[n][l][tiles_n], colp[n][tiles_m*4][s]  (log2 domain)
//   pass 1 (conf):  S tile recomputed (bit-identical) -> conf = 2^(2x - rmax_i - cmax_j) * rinv_i * cinv_j (ONE exp per
//                   element) -> per-row best (conf, first j) merged with a 64-bit atomicMax, per-column best conf with
//                   a 32-bit atomicMax.
//   then mnn_from_best_kernel applies threshold / border / mutual check on vectors.
// CTA pairs (r02): the two CTAs of a cluster take neighbouring 128-row blocks of f0 against the same 256 columns of f1 and
// run ONE tcgen05.mma.cta_group::2 (M = 256) per K step; each CTA stages only half of the f1 tile, so per 64-wide k-block a
// CTA's shared memory sees 32 KB written + 32 KB read (512 cycles at 128 B/clk) for 512 cycles of MMA, instead of
// 48 + 48 KB (768 cycles): the single-CTA kernel sat at 67-68 % tensor pipe for exactly this reason.
// Nothing of size L x S touches HBM: per call the kernels read the packed operands (2 x n*L*3C fp16) and write
// O(n*(L+S)*tiles) partials, so the similarity contraction is tensor-bound instead of bound by a 1.5 GB fp32 store.
//
// Exact tie semantics.  The reference's match of row i is the FIRST j with conf_ij == rowmax_i AND conf_ij == colmax_j AND
// the border test (coarse_matching.py:176-188: mask.max(dim=2) returns the first True).  Pass 1 keeps (rowmax_i, smallest
// j attaining it); that j is the reference's answer unless the row maximum is attained more than once and the smallest
// such j fails the column / border test.  Pass 1 therefore also raises a per-row tie flag (an equal value seen inside the
// thread's chunk scan, or an equal confidence returned by the 64-bit atomicMax when column tiles merge).  Rows that are
// flagged AND rejected are re-scanned exactly by pass 2 (MODE 2: the same tile recomputed bit-identically, every j tested
// against rowmax_i / colmax_j / border, smallest j kept with atomicMin).  Pass 2 exits at once when no row needs it
// (the normal case: a few microseconds) and only visits the 128-row blocks that contain such rows.
#include "common.cuh"
#include "ptx.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

namespace sf {
constexpr int kBM = 128, kBN = 256, kStages = 5;
constexpr int kStageA = kBM * 128, kStageB = (kBN / 2) * 128, kStage = kStageA + kStageB;     // B: this CTA's half of the f1 rows
constexpr int kEpiWarps = 8;            // two epilogue warps per TMEM lane quadrant (each takes 128 of the 256 columns):
                                        // with one warp per scheduler every dependency stall of the softmax math was exposed
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStaging = kEpiWarps * 4096;
constexpr int kSmem = kStages * kStage + kStaging + 1024 + 256;
constexpr int kTmemCols = 512;
}  // namespace sf

struct SimFusedParams {
  int n, l, s, kblocks, tiles_m, tiles_n;
  float scale2;                 // out_scale * log2(e): logits in the log2 domain
  float2* rowp; float2* colp;   // pass 0 outputs
  const float* row_m2; const float* row_inv; const float* col_m2; const float* col_inv;   // pass 1 inputs
  unsigned long long* row_best; unsigned* col_best;                                        // pass 1 outputs
  int* row_tie;                 // pass 1 output: row maximum attained more than once (may be a stale over-approximation)
  // pass 2 (exact re-scan of tied + rejected rows)
  const int* rescan_cnt;        // [n * tiles_m + 1]: flagged rows per 128-row block; last entry = total
  int* rescan_j;                // [n * l]: INT_MAX for rows to re-scan (atomicMin target), -2 otherwise
  int border, h0c, w0c, h1c, w1c;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE>
__global__ void __launch_bounds__(sf::kThreads, 1)
sim_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SimFusedParams p) {
  using namespace sf;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStaging);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mp = (p.tiles_m + 1) >> 1;                 // pairs of 128-row blocks
  const int total_tiles = p.n * tiles_mp * p.tiles_n;        // work items of a CTA pair
  if constexpr (MODE == 2) {
    if (p.rescan_cnt[p.n * p.tiles_m] == 0) return;        // nothing to re-scan (uniform over the grid, before any cluster barrier)
  }
  const uint32_t rank = ptx::cluster_ctarank();
  const int t0 = blockIdx.x >> 1, tstep = gridDim.x >> 1;
  // work item t -> (batch, row-block pair, column block); this CTA's row block is 2 * pair + rank (may lie beyond L:
  // TMA zero-fills it and every row fails row_ok)
  auto decode = [&](int t, int& batch, int& m_blk, int& n_blk) {
    n_blk = t % p.tiles_n;
    const int rest = t / p.tiles_n;
    m_blk = (rest % tiles_mp) * 2 + (int)rank;
    batch = rest / tiles_mp;
  };
  // MODE 2 visits only the pairs of 128-row blocks that hold a row to re-scan (all roles of both CTAs skip the same items)
  auto skip_tile = [&](int t) -> bool {
    if constexpr (MODE == 2) {
      const int rest = t / p.tiles_n, mp = rest % tiles_mp, batch = rest / tiles_mp;
      const int* c = p.rescan_cnt + batch * p.tiles_m + 2 * mp;
      return c[0] == 0 && (2 * mp + 1 >= p.tiles_m || c[1] == 0);
    } else return false;
  };

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
      // tmem_empty is used in the leader CTA only: the epilogue warps of both CTAs arrive there
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 2 * kEpiWarps); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc2(tmem_slot, kTmemCols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues (uniform-register operands, see conv_tc.cu).  Both CTAs'
    // boxes (own 128 rows of f0, own half of the 256 rows of f1) count on the LEADER's full barrier.
    {
      int stage = 0; uint32_t phase = 0;
      for (int t = t0; t < total_tiles; t += tstep) {
        if (skip_tile(t)) continue;
        int batch, m_blk, n_blk;
        decode(t, batch, m_blk, n_blk);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * kStage);
            ptx::tma_load_3d_2sm(sa, &tmA, &full_bar[stage], kb * 64, m_blk * kBM, batch);
            ptx::tma_load_3d_2sm(sa + kStageA, &tmB, &full_bar[stage], kb * 64, n_blk * kBN + (int)rank * (kBN / 2), batch);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (leader CTA only): whole warp converged, one elected lane issues; M = 256 over the CTA pair
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0 /*f16*/, 2 * kBM, kBN);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, ptx::smem_addr(smem), 0);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t eb0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(empty_bar) - smem);
      const uint32_t tf0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(tmem_full) - smem);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = t0; t < total_tiles; t += tstep) {
        if (skip_tile(t)) continue;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // both CTAs' epilogue warps have drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tb + acc * kBN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem0 + stage * kStage;
          const uint32_t a_lo = ptx::umma_desc_lo(sa), b_lo = ptx::umma_desc_lo(sa + kStageA);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma2_f16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, idesc, (kb | k) ? 1u : 0u);
            ptx::umma2_commit_mcast_addr(eb0 + stage * 8, (uint16_t)0x3);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma2_commit_mcast_addr(tf0 + acc * 8, (uint16_t)0x3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;                           // TMEM lane quadrant (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                    // which 128-column half of the tile this warp handles
    uint8_t* box = staging + (warp - 2) * 4096;          // this warp's 32 x 32 fp32 transpose box (128B-swizzled rows)
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = t0; t < total_tiles; t += tstep) {
      if (skip_tile(t)) continue;
      int batch, m_blk, n_blk;
      decode(t, batch, m_blk, n_blk);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_blk * kBM + quad * 32 + lane;
      const bool row_ok = row < p.l;
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * kBN + half * (kBN / 2);
      const int col0 = n_blk * kBN + half * (kBN / 2);
      // (a row block beyond L - the odd block's partner in the last pair - does no epilogue work and writes nothing)
      const int nch = m_blk >= p.tiles_m ? 0 : max(0, min(kBN / 64, (p.s - col0 + 31) / 32));
      float rm = -INFINITY, rsum = 0.f;                  // MODE 0: running row (max, sum)
      float r_m2 = 0.f, r_inv = 0.f;                     // MODE 1: final row stats
      float best = 0.f; int best_j = 0;
      bool tied = false;                                 // MODE 1: the running row maximum was seen more than once
      unsigned want_bits = 0u; int min_j = 0x7fffffff;   // MODE 2: row maximum to look for (0 = row not re-scanned), first hit
      bool row_border_ok = true;
      if constexpr (MODE >= 1) {
        if (row_ok) { r_m2 = p.row_m2[(int64_t)batch * p.l + row]; r_inv = p.row_inv[(int64_t)batch * p.l + row]; }
      }
      if constexpr (MODE == 2) {
        if (row_ok && p.rescan_j[(int64_t)batch * p.l + row] != -2) want_bits = (unsigned)(p.row_best[(int64_t)batch * p.l + row] >> 32);
        if (p.border > 0) {
          const int r0 = row / p.w0c, c0 = row % p.w0c;
          row_border_ok = r0 >= p.border && r0 < p.h0c - p.border && c0 >= p.border && c0 < p.w0c - p.border;
        }
      }
      float v[32];
      if (nch > 0) ptx::tmem_ld_32x32(t_row, v);
#pragma unroll 1
      for (int ci = 0; ci < nch; ++ci) {
        ptx::tmem_ld_wait();
        const int gc = col0 + ci * 32;
        const bool full = gc + 32 <= p.s;
        float c_m2 = 0.f, c_inv = 0.f;                   // MODE 1 / 2: stats of column gc + lane
        unsigned c_best = 0u;                            // MODE 2: column maximum of column gc + lane (0 if its border test fails)
        if constexpr (MODE >= 1) {
          if (gc + lane < p.s) { c_m2 = p.col_m2[(int64_t)batch * p.s + gc + lane]; c_inv = p.col_inv[(int64_t)batch * p.s + gc + lane]; }
        }
        if constexpr (MODE == 2) {
          if (gc + lane < p.s) {
            c_best = p.col_best[(int64_t)batch * p.s + gc + lane];
            if (p.border > 0) {
              const int r1 = (gc + lane) / p.w1c, c1 = (gc + lane) % p.w1c;
              if (!(r1 >= p.border && r1 < p.h1c - p.border && c1 >= p.border && c1 < p.w1c - p.border)) c_best = 0u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.scale2;
        if (!full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (gc + j >= p.s) v[j] = -INFINITY;
        }
        if constexpr (MODE == 0) {
          // ---- row (max, sum): thread-local online update over this chunk
          float cm = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
          const float nm = fmaxf(rm, cm);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += ex2f(v[j] - nm); a1 += ex2f(v[j + 1] - nm); a2 += ex2f(v[j + 2] - nm); a3 += ex2f(v[j + 3] - nm);
          }
          rsum = rsum * ex2f(rm - nm) + (a0 + a1) + (a2 + a3);
          rm = nm;
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = -INFINITY;      // rows beyond L must not enter the column statistics
          }
        } else if constexpr (MODE == 1) {
          // ---- conf = softmax_row * softmax_col with one exponential: 2^(2x - rmax - cmax_j) * rinv * cinv_j
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float cmj = __shfl_sync(0xffffffffu, c_m2, j), cij = __shfl_sync(0xffffffffu, c_inv, j);
            const float cf = row_ok ? ex2f(fmaf(2.f, v[j], -r_m2 - cmj)) * (r_inv * cij) : 0.f;   // -inf logits -> 0
            v[j] = cf;
            tied = (cf == best) ? true : ((cf > best) ? false : tied);
            if (cf > best) { best = cf; best_j = gc + j; }     // strict >: the first j wins inside this row
          }
        } else {
          // ---- exact re-scan: the same expression as pass 1 (bit-identical), every j tested against both maxima
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float cmj = __shfl_sync(0xffffffffu, c_m2, j), cij = __shfl_sync(0xffffffffu, c_inv, j);
            const unsigned cbj = __shfl_sync(0xffffffffu, c_best, j);
            const float cf = row_ok ? ex2f(fmaf(2.f, v[j], -r_m2 - cmj)) * (r_inv * cij) : 0.f;
            const unsigned bits = __float_as_uint(cf);
            if (want_bits != 0u && bits == want_bits && bits == cbj && row_border_ok && gc + j < min_j) min_j = gc + j;
          }
          if (ci + 1 < nch) ptx::tmem_ld_32x32(t_row + (ci + 1) * 32, v);
          continue;
        }
        // ---- transpose through the swizzled box: lane <- column (gc + lane) over the warp's 32 rows
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        float x[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) x[r] = *reinterpret_cast<const float*>(box + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
        __syncwarp();
        if (ci + 1 < nch) ptx::tmem_ld_32x32(t_row + (ci + 1) * 32, v);     // v is free: prefetch the next chunk
        float m0 = x[0], m1 = x[1], m2 = x[2], m3 = x[3];
#pragma unroll
        for (int r = 4; r < 32; r += 4) { m0 = fmaxf(m0, x[r]); m1 = fmaxf(m1, x[r + 1]); m2 = fmaxf(m2, x[r + 2]); m3 = fmaxf(m3, x[r + 3]); }
        const float cmx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        if constexpr (MODE == 0) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          if (cmx > -INFINITY) {
#pragma unroll
            for (int r = 0; r < 32; r += 4) {
              a0 += ex2f(x[r] - cmx); a1 += ex2f(x[r + 1] - cmx); a2 += ex2f(x[r + 2] - cmx); a3 += ex2f(x[r + 3] - cmx);
            }
          }
          if (gc + lane < p.s)
            p.colp[((int64_t)batch * p.tiles_m * 4 + m_blk * 4 + quad) * p.s + gc + lane] = make_float2(cmx, (a0 + a1) + (a2 + a3));   // columns are disjoint between the two halves
        } else {
          if (gc + lane < p.s && cmx > 0.f) atomicMax(&p.col_best[(int64_t)batch * p.s + gc + lane], __float_as_uint(cmx));
        }
      }
      if (row_ok) {
        if constexpr (MODE == 0) {
          p.rowp[((int64_t)batch * p.l + row) * (p.tiles_n * 2) + n_blk * 2 + half] = make_float2(rm, rsum);
        } else if constexpr (MODE == 1) {
          if (best > 0.f) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)best_j);
            const unsigned long long old = atomicMax(&p.row_best[(int64_t)batch * p.l + row], key);
            // an equal confidence already stored by another column tile == a tie across tiles
            if (tied || (unsigned)(old >> 32) == __float_as_uint(best)) p.row_tie[(int64_t)batch * p.l + row] = 1;
          }
        } else if (min_j != 0x7fffffff) {
          atomicMin(&p.rescan_j[(int64_t)batch * p.l + row], min_j);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);     // the MMA issuer lives in the leader CTA
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                 // no CTA leaves (or frees tensor memory) while its peer may still use it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, sf::kTmemCols);
  }
}

// merge `parts` log2-domain (max, sum) partials -> final max and reciprocal sum (layout arguments as merge_stats_kernel)
__global__ void merge_stats2_kernel(const float2* __restrict__ in, int64_t elems, int parts, int64_t estride, int64_t kstride,
                                    int64_t group, int64_t gstride, float* __restrict__ out_m2, float* __restrict__ out_inv) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  const int64_t g = e / group, i = e - g * group;
  const float2* q = in + g * gstride + i * estride;
  float m = -INFINITY;
  for (int k = 0; k < parts; ++k) m = fmaxf(m, q[k * kstride].x);
  float sum = 0.f;
  for (int k = 0; k < parts; ++k) { const float2 v = q[k * kstride]; if (v.x > -INFINITY) sum += v.y * exp2f(v.x - m); }
  out_m2[e] = m;
  out_inv[e] = 1.f / sum;
}

// threshold + border + mutual check on the per-row / per-column bests (coarse_matching.py:161-188).  Rows whose maximum
// is tied and whose smallest-j candidate is rejected are queued for the exact re-scan (see the header comment).
__global__ void mnn_from_best_kernel(const unsigned long long* __restrict__ row_best, const unsigned* __restrict__ col_best,
                                     const int* __restrict__ row_tie, int n, int l, int s, float thr, int border, int h0c,
                                     int w0c, int h1c, int w1c, int tiles_m, int* __restrict__ match_j,
                                     float* __restrict__ match_conf, int* __restrict__ rescan_j, int* __restrict__ rescan_cnt) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * l) return;
  const int b = (int)(idx / l), i = (int)(idx - (int64_t)b * l);
  const unsigned long long key = row_best[idx];
  const unsigned cbits = (unsigned)(key >> 32);
  const int j = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
  const float conf = __uint_as_float(cbits);
  const bool cand = key != 0ull && conf > thr && j >= 0 && j < s;
  bool ok = cand && col_best[(int64_t)b * s + j] == cbits;
  if (ok && border > 0) {
    const int r0 = i / w0c, c0 = i % w0c, r1 = j / w1c, c1 = j % w1c;
    ok = r0 >= border && r0 < h0c - border && c0 >= border && c0 < w0c - border &&
         r1 >= border && r1 < h1c - border && c1 >= border && c1 < w1c - border;
  }
  match_j[idx] = ok ? j : -1;
  match_conf[idx] = ok ? conf : 0.f;
  const bool rescan = cand && !ok && row_tie[idx] != 0;
  rescan_j[idx] = rescan ? 0x7fffffff : -2;
  if (rescan) { atomicAdd(&rescan_cnt[b * tiles_m + i / sf::kBM], 1); atomicAdd(&rescan_cnt[n * tiles_m], 1); }
}

// after pass 2: re-scanned rows take the first j that passed every test (or stay unmatched)
__global__ void mnn_apply_rescan_kernel(const unsigned long long* __restrict__ row_best, const int* __restrict__ rescan_j,
                                        const int* __restrict__ rescan_cnt, int64_t rows, int total_idx,
                                        int* __restrict__ match_j, float* __restrict__ match_conf) {
  if (rescan_cnt[total_idx] == 0) return;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows) return;
  const int j = rescan_j[idx];
  if (j == -2 || j == 0x7fffffff) return;
  match_j[idx] = j;
  match_conf[idx] = __uint_as_float((unsigned)(row_best[idx] >> 32));
}

}  // namespace gf

using namespace gf;

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int64_t gf_coarse_match_fused_workspace_bytes(int n, int l, int s) {
  const int64_t tiles_m = gf_cdiv(l, sf::kBM), tiles_n = gf_cdiv(s, sf::kBN);
  int64_t b = 0;
  b += align_up((int64_t)n * l * tiles_n * 2 * 8, 256);        // rowp (two column halves per tile)
  b += align_up((int64_t)n * tiles_m * 4 * s * 8, 256);        // colp
  b += 2 * align_up((int64_t)n * l * 4, 256) + 2 * align_up((int64_t)n * s * 4, 256);   // row/col m2 + inv
  b += align_up((int64_t)n * l * 8, 256) + align_up((int64_t)n * s * 4, 256);           // row_best, col_best
  b += 2 * align_up((int64_t)n * l * 4, 256) + align_up(((int64_t)n * tiles_m + 1) * 4, 256);   // row_tie, rescan_j, rescan_cnt
  return b;
}

// a3 [n,l,c3] / b3 [n,s,c3]: split-fp16 packed operands (gf_pack_split_f16).  Outputs match_j / match_conf [n*l]
// (feed gf_compact_coarse).  out_scale = 1 / temperature.
static int coarse_match_fused_impl(const void* a3, const void* b3, int n, int l, int s, int c3, float out_scale,
                                   float thr, int border, int h0c, int w0c, int h1c, int w1c, void* workspace,
                                   int* match_j, float* match_conf, gf_stream_t stream, int only_pass);

extern "C" int gf_coarse_match_fused(const void* a3, const void* b3, int n, int l, int s, int c3, float out_scale,
                                     float thr, int border, int h0c, int w0c, int h1c, int w1c, void* workspace,
                                     int* match_j, float* match_conf, gf_stream_t stream) {
  return coarse_match_fused_impl(a3, b3, n, l, s, c3, out_scale, thr, border, h0c, w0c, h1c, w1c, workspace, match_j,
                                 match_conf, stream, -1);
}

// profiling hook: launch only the tensor pass `pass` (0 = statistics, 1 = confidences) on an already used workspace
extern "C" int gf_coarse_match_fused_pass(const void* a3, const void* b3, int n, int l, int s, int c3, float out_scale,
                                          void* workspace, int pass, gf_stream_t stream) {
  if (pass != 0 && pass != 1) return gf_set_error(GF_ERR_ARG, "gf_coarse_match_fused_pass: pass must be 0 or 1");
  return coarse_match_fused_impl(a3, b3, n, l, s, c3, out_scale, 0.f, 0, 1, l, 1, s, workspace, nullptr, nullptr, stream, pass);
}

static int coarse_match_fused_impl(const void* a3, const void* b3, int n, int l, int s, int c3, float out_scale,
                                   float thr, int border, int h0c, int w0c, int h1c, int w1c, void* workspace,
                                   int* match_j, float* match_conf, gf_stream_t stream, int only_pass) {
  if (n <= 0 || l <= 0 || s <= 0 || c3 <= 0 || (c3 % 64) || h0c * w0c != l || h1c * w1c != s || workspace == nullptr)
    return gf_set_error(GF_ERR_ARG, "gf_coarse_match_fused: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_m = gf_cdiv(l, sf::kBM), tiles_n = gf_cdiv(s, sf::kBN);
  uint8_t* w = (uint8_t*)workspace;
  float2* rowp = (float2*)w; w += align_up((int64_t)n * l * tiles_n * 2 * 8, 256);
  float2* colp = (float2*)w; w += align_up((int64_t)n * tiles_m * 4 * s * 8, 256);
  float* row_m2 = (float*)w; w += align_up((int64_t)n * l * 4, 256);
  float* row_inv = (float*)w; w += align_up((int64_t)n * l * 4, 256);
  float* col_m2 = (float*)w; w += align_up((int64_t)n * s * 4, 256);
  float* col_inv = (float*)w; w += align_up((int64_t)n * s * 4, 256);
  // [row_best | col_best | row_tie | rescan_cnt] are contiguous: one memset clears them
  uint8_t* zero0 = w;
  unsigned long long* row_best = (unsigned long long*)w; w += align_up((int64_t)n * l * 8, 256);
  unsigned* col_best = (unsigned*)w; w += align_up((int64_t)n * s * 4, 256);
  int* row_tie = (int*)w; w += align_up((int64_t)n * l * 4, 256);
  int* rescan_cnt = (int*)w; w += align_up(((int64_t)n * tiles_m + 1) * 4, 256);
  const size_t zero_bytes = (size_t)(w - zero0);
  int* rescan_j = (int*)w;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap(&ta, a3, 2, c3, l, n, c3, (int64_t)l * c3, sf::kBM))) return rc;
  if ((rc = make_tmap(&tb, b3, 2, c3, s, n, c3, (int64_t)s * c3, sf::kBN / 2))) return rc;     // half a column block per CTA of the pair
  GF_SMEM_OPTIN(sim_fused_kernel<0>, sf::kSmem);
  GF_SMEM_OPTIN(sim_fused_kernel<1>, sf::kSmem);
  GF_SMEM_OPTIN(sim_fused_kernel<2>, sf::kSmem);
  SimFusedParams p{};
  p.n = n; p.l = l; p.s = s; p.kblocks = c3 / 64; p.tiles_m = tiles_m; p.tiles_n = tiles_n;
  p.scale2 = out_scale * 1.4426950408889634f;
  p.rowp = rowp; p.colp = colp; p.row_m2 = row_m2; p.row_inv = row_inv; p.col_m2 = col_m2; p.col_inv = col_inv;
  p.row_best = row_best; p.col_best = col_best; p.row_tie = row_tie; p.rescan_cnt = rescan_cnt; p.rescan_j = rescan_j;
  p.border = border; p.h0c = h0c; p.w0c = w0c; p.h1c = h1c; p.w1c = w1c;
  const int64_t items = (int64_t)n * ((tiles_m + 1) / 2) * tiles_n;           // work items of a CTA pair
  const int grid = 2 * (int)(items < num_sms() / 2 ? items : num_sms() / 2);
  auto launch = [&](auto kern) -> int {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(sf::kThreads); cfg.dynamicSmemBytes = sf::kSmem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p) != cudaSuccess) return gf_set_error(GF_ERR_LAUNCH, cudaGetErrorString(cudaGetLastError()));
    return GF_OK;
  };
  if (only_pass == 0) { if ((rc = launch(sim_fused_kernel<0>))) return rc; g_launches++; GF_CHECK_LAUNCH(); return GF_OK; }
  if (only_pass == 1) { if ((rc = launch(sim_fused_kernel<1>))) return rc; g_launches++; GF_CHECK_LAUNCH(); return GF_OK; }
  if ((rc = launch(sim_fused_kernel<0>))) return rc;
  merge_stats2_kernel<<<gf_cdiv((int64_t)n * l, 256), 256, 0, st>>>(rowp, (int64_t)n * l, tiles_n * 2, tiles_n * 2, 1, (int64_t)n * l, 0,
                                                                   row_m2, row_inv);
  merge_stats2_kernel<<<gf_cdiv((int64_t)n * s, 256), 256, 0, st>>>(colp, (int64_t)n * s, tiles_m * 4, 1, s, s,
                                                                   (int64_t)tiles_m * 4 * s, col_m2, col_inv);
  cudaMemsetAsync(zero0, 0, zero_bytes, st);
  if ((rc = launch(sim_fused_kernel<1>))) return rc;
  mnn_from_best_kernel<<<gf_cdiv((int64_t)n * l, 256), 256, 0, st>>>(row_best, col_best, row_tie, n, l, s, thr, border, h0c, w0c,
                                                                    h1c, w1c, tiles_m, match_j, match_conf, rescan_j, rescan_cnt);
  // exact tie handling: both kernels return at once unless a tied row maximum was rejected (see the header comment)
  if ((rc = launch(sim_fused_kernel<2>))) return rc;
  mnn_apply_rescan_kernel<<<gf_cdiv((int64_t)n * l, 256), 256, 0, st>>>(row_best, rescan_j, rescan_cnt, (int64_t)n * l, n * tiles_m,
                                                                       match_j, match_conf);
  g_launches += 7;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
