// tcgen05 GEMM core of libgeoformer_sm100.so.
//
//   Y[b][M,N] = epilogue( [A | A2][b][M, K] * B[b][N, K]^T )          (both operands K-major)
//
// One persistent CTA per SM, warp-specialised (guide "Canonical Blackwell GEMM" anatomy):
//   warp 0      TMA producer   : cp.async.bulk.tensor (128B-swizzled boxes) into a 4..6 stage smem ring
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (M=128, N=BN, K=32 bytes per instruction),
//                                accumulating in TMEM; double-buffered accumulator (2 x BN columns)
//   warps 2..5  epilogue       : tcgen05.ld (thread == output row), fused bias / activation / LayerNorm /
//                                residual; 32x32 boxes staged in swizzled smem and written with TMA stores
// kind::tf32 consumes fp32 activations straight from HBM (no conversion pass); kind::f16 is used by the
// split-fp16 similarity.  Tile: 128 x BN x (128 bytes of K).
#include "common.cuh"
#include "ptx.cuh"

#include <cuda_fp16.h>

#include <atomic>
#include <cstdlib>

namespace gf {

struct GemmParams {
  float* Y;
  int64_t ldy;
  int64_t y_batch_stride;
  int M, N, batches;
  int kblocks1, kblocks2;
  int epi, act_cols;
  const float* bias;
  const float* rowbias;
  int rowbias_group;
  const float* gamma;
  const float* beta;
  const float* residual;
  int64_t ldres;
  float out_scale;
  const int* m_dev;
  int tiles_n;
  int cluster;     // 1 or 2 CTAs per cluster (2: B operand multicast between vertically adjacent M tiles)
  int tma_store;   // 1: epilogue stages 32x32 boxes in swizzled smem and TMA-stores them (needs ldy % 4 == 0)
};

constexpr int kBM = 128;
constexpr int kStageABytes = kBM * 128;

template <int BN, bool FULL = false> struct GemmCfg {
  static constexpr int kStageBBytes = BN * 128;
  static constexpr int kStageBytes = kStageABytes + kStageBBytes;
  static constexpr int kStages = (BN == 256) ? 3 : (FULL ? 4 : 5);     // the LayerNorm partials of the full epilogue need 4 KB
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStagingBytes = 4 * 4 * 4096;   // per epilogue warp: ring of 4 x (32 rows x 128 B) output boxes
  static constexpr int kLnBytes = FULL ? 2 * 2 * 128 * 8 : 0;   // LayerNorm partials: 2 tile parities x 2 column halves x 128 rows x (sum, sum of squares)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kLnBytes;
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float ex2_ftz(float x) {      // one MUFU.EX2; elu(x)+1 = e^x for x <= 0 needs no denormal care
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tanh(x) = sign(x) * (1 - 2 / (exp(2|x|) + 1)); abs error ~1e-7 (well below the tf32 operand rounding)
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.f * fabsf(x));
  const float t = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(t, x);
}

// FULL = false: lean epilogue (scale, bias, row-group bias, activation) with the TMEM loads software-pipelined one chunk
//                ahead; FULL = true : + LayerNorm, residual (staged through shared memory).
// CL = 2 (opt-in, env GF_CLUSTER2): two CTAs of a cluster work on vertically adjacent M tiles of the same N tile and
// share the B operand: each CTA TMA-loads half of every B box and multicasts it into both shared memories, cutting
// the per-SM L2 -> smem operand traffic from (A + B) to (A + B/2) per k-block.  Measured on B200: no gain (the
// short-K GEMMs of this model are bound by their HBM output stream, not by the operand feed), so CL = 1 is default.
// OUT16 (lean epilogue only): the output is stored as fp16 (32 x 32 boxes of 64-byte rows, 64B swizzle).  Used for
// intermediates whose only consumers round to a 10-bit mantissa anyway (tf32 / fp16 MMA operands) or are the linear
// attention kernels: halves the HBM stream these short-K GEMMs are bound by.
template <int KIND, int BN, bool FULL, int CL, bool OUT16>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY, const GemmParams p) {
  using Cfg = GemmCfg<BN, FULL>;
  constexpr int BKE = (KIND == 0) ? 32 : 64;   // elements per 128-byte K block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int m_live = p.M;
  if (p.m_dev != nullptr) m_live = min(p.M, max(0, *p.m_dev));
  const int cta_rank = (CL == 2) ? (int)ptx::cluster_ctarank() : 0;
  const int tiles_m = ((m_live + kBM - 1) / kBM + CL - 1) / CL;          // M-tile groups (pairs when CL == 2)
  const int total_tiles = p.batches * tiles_m * p.tiles_n;
  const int t_first = blockIdx.x / CL, t_step = gridDim.x / CL;
  const int kblocks = p.kblocks1 + p.kblocks2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmA2);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], CL); }
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 8); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  ptx::tc_fence_before();
  if constexpr (CL == 2) ptx::cluster_sync(); else __syncthreads();   // peers must see initialised barriers
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // whole warp converged, one elected lane issues (uniform-register operands, see conv_tc.cu)
    {
      int stage = 0; uint32_t phase = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int n_blk = t % p.tiles_n;
        const int rest = t / p.tiles_n;
        const int m_blk = (rest % tiles_m) * CL + cta_rank;
        const int batch = rest / tiles_m;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kStageABytes;
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (kb < p.kblocks1) ptx::tma_load_3d(sa, &tmA, &full_bar[stage], kb * BKE, m_blk * kBM, batch);
            else                 ptx::tma_load_3d(sa, &tmA2, &full_bar[stage], (kb - p.kblocks1) * BKE, m_blk * kBM, batch);
            if constexpr (CL == 2) {
              ptx::tma_load_3d_mcast(sb + cta_rank * (BN / 2) * 128, &tmB, &full_bar[stage], kb * BKE,
                                     n_blk * BN + cta_rank * (BN / 2), batch, (uint16_t)0x3);
            } else {
              ptx::tma_load_3d(sb, &tmB, &full_bar[stage], kb * BKE, n_blk * BN, batch);
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // whole warp converged, one elected lane issues: operands stay in uniform registers (see conv_tc.cu)
    {
      constexpr uint32_t idesc = ptx::umma_idesc(KIND == 0 ? 2 : 0, kBM, BN);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, ptx::smem_addr(smem), 0);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t fb0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(full_bar) - smem);
      const uint32_t eb0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(empty_bar) - smem);
      const uint32_t tf0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(tmem_full) - smem);
      (void)fb0;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tb + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem0 + stage * Cfg::kStageBytes;
          const uint32_t a_lo = ptx::umma_desc_lo(sa), b_lo = ptx::umma_desc_lo(sa + kStageABytes);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)      // advance 32 bytes of K inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              ptx::umma_lo<KIND>(d_tmem, a_lo + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, idesc, (kb | k) ? 1u : 0u);
            if constexpr (CL == 2) ptx::umma_commit_mcast_addr(eb0 + stage * 8, (uint16_t)0x3);   // frees the slot in both CTAs
            else ptx::umma_commit_addr(eb0 + stage * 8);      // frees the smem slot when these MMAs retire
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit_addr(tf0 + acc * 8);     // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    // 8 warps, two per TMEM lane quadrant, each owning half of the tile's columns (with one epilogue warp per scheduler
    // every dependency stall was exposed).  r02: also for the LayerNorm epilogue - with 4 warps and three sweeps over the
    // accumulator it took 17.6 us per 128 x 256 tile (ncu: tensor pipe 14.6 %, DRAM 37 %: bound by neither) where the HBM
    // streams of the tile need 8 us; now each warp makes ONE statistics sweep over its 128 columns (sum and sum of
    // squares), the two warps of a quadrant exchange partials through shared memory, then one emit sweep.
    constexpr int EW = 8;
    constexpr int WCOLS = BN / (EW / 4);                 // columns per epilogue warp
    constexpr int G = 1;                                 // boxes per TMA-store group (ring of 2 groups per warp)
    float2* lnx = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(tmem_slot) + 64);      // [2 parities][2 halves][128 rows]
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t grp = 0;          // TMA-store group counter: 2 boxes per fence / commit, ring of 2 groups per warp
    int pend = 0;
    int pend_col[2] = {0, 0};
    for (int t = t_first; t < total_tiles; t += t_step) {
      const int n_blk = t % p.tiles_n;
      const int rest = t / p.tiles_n;
      const int m_blk = (rest % tiles_m) * CL + cta_rank;
      const int batch = rest / tiles_m;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m_blk * kBM + quad * 32 + lane;
      const bool row_ok = row < m_live;
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + half * WCOLS;
      float* yrow = p.Y + (int64_t)batch * p.y_batch_stride + (int64_t)row * p.ldy;
      const float* rrow = p.residual ? p.residual + (int64_t)row * p.ldres : nullptr;
      const float* rb = p.rowbias ? p.rowbias + (int64_t)(row / p.rowbias_group) * p.N : nullptr;
      const int col0 = n_blk * BN + half * WCOLS;
      uint8_t* wstage = staging + (warp - 2) * (2 * G * 4096);   // this warp's ring of 2 groups x G boxes of 4 KB
      // residual tile slice of this lane for one 32x32 chunk: 8 coalesced float4 (4 rows x 128 B per instruction);
      // software-pipelined one chunk ahead so that the global-load latency hides behind the previous chunk
      float4 rnext[8];
      auto fetch_residual = [&](int gcol, float4* dst) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3), ch = lane & 7;
          const int grow = m_blk * kBM + quad * 32 + rr;
          dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < m_live && gcol + ch * 4 < p.N)
            dst[it] = __ldg(reinterpret_cast<const float4*>(p.residual + (int64_t)grow * p.ldres + gcol + ch * 4));
        }
      };
      const bool stage_res = FULL && p.tma_store && p.residual != nullptr;
      if constexpr (FULL) { if (stage_res) fetch_residual(col0, rnext); }
      float mean = 0.f, rstd = 1.f;
      if constexpr (FULL) {
        if (p.epi & GF_EPI_LN) {
          // LayerNorm over the full row (BN == N): one sweep over this warp's half of the columns for (sum, sum of
          // squares), partials of the two halves combined through shared memory (double-buffered by tile parity)
          float s1 = 0.f, s2 = 0.f;
          {
            float v[32];
            ptx::tmem_ld_32x32(t_row, v);
#pragma unroll 1
            for (int c = 0; c < WCOLS; c += 32) {
              ptx::tmem_ld_wait();
              float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float x0 = v[j] * p.out_scale, x1 = v[j + 1] * p.out_scale;
                a0 += x0; a1 += x1; q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
              }
              if (c + 32 < WCOLS) ptx::tmem_ld_32x32(t_row + c + 32, v);
              s1 += a0 + a1; s2 += q0 + q1;
            }
          }
          float2* mine = lnx + (acc * 2 + half) * 128 + quad * 32 + lane;
          float2* other = lnx + (acc * 2 + (half ^ 1)) * 128 + quad * 32 + lane;
          *mine = make_float2(s1, s2);
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");     // the two warps of this lane quadrant
          const float2 o = *other;
          mean = (s1 + o.x) * (1.f / BN);
          rstd = rsqrtf(fmaxf((s2 + o.y) * (1.f / BN) - mean * mean, 0.f) + 1e-5f);
        }
      }
      // one 32-column chunk: math on v[] (thread == row), swizzled smem box, grouped TMA store
      auto emit_chunk = [&](float (&v)[32], int c, bool prefetch_next) {
        const int gc = col0 + c;
        uint8_t* box = wstage + ((grp & 1) * G + pend) * 4096;
        if (p.tma_store) {
          if (pend == 0) {
            if (lane == 0) ptx::bulk_wait_read<1>();       // the group issued 2 groups ago no longer reads these boxes
            __syncwarp();
          }
          if constexpr (FULL) {
            if (stage_res) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3), ch = lane & 7;
                *reinterpret_cast<float4*>(box + rr * 128 + ((ch ^ (rr & 7)) << 4)) = rnext[it];
              }
              if (c + 32 < WCOLS && gc + 32 < p.N) fetch_residual(gc + 32, rnext);
              __syncwarp();
            }
          }
        }
        // Epilogue math as warp-uniform branches around whole 32-element loops: only the taken
        // variant issues instructions (a predicated all-in-one body cost ~280 issue slots per element).
        const bool full = gc + 32 <= p.N;
        if (p.out_scale != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
        }
        if (p.bias) {
          if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gc) + j);
              v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          } else {
            for (int j = 0; j < 32; ++j) if (gc + j < p.N) v[j] += __ldg(p.bias + gc + j);
          }
        }
        if (rb && row_ok) {        // per-row-group bias (fine merge_feat: one coarse-context row per 25-token window)
          if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb + gc) + j);
              v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
          } else {
            for (int j = 0; j < 32; ++j) if (gc + j < p.N) v[j] += __ldg(rb + gc + j);
          }
        }
        if (p.epi & GF_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.epi & GF_EPI_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
        }
        if ((p.epi & GF_EPI_ELU1) && gc < p.act_cols) {
          if (gc + 32 <= p.act_cols) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] + 1.f : ex2_ftz(v[j] * 1.4426950408889634f);
          } else {
            for (int j = 0; j < 32; ++j) if (gc + j < p.act_cols) v[j] = v[j] > 0.f ? v[j] + 1.f : ex2_ftz(v[j] * 1.4426950408889634f);
          }
        }
        if constexpr (FULL) {
          if (p.epi & GF_EPI_LN) {                            // BN == N: chunks are always full
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + gc) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + gc) + j);
              v[4 * j] = (v[4 * j] - mean) * rstd * g4.x + b4.x;
              v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
              v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
              v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
            }
          }
          if (rrow && row_ok && !p.tma_store) {
            for (int j = 0; j < 32; ++j) if (gc + j < p.N) v[j] += __ldg(rrow + gc + j);
          }
        }
        if constexpr (OUT16) {                               // host guarantees tma_store for fp16 outputs
          uint8_t* myrow = box + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_half2(v[8 * j], v[8 * j + 1]); u.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
            u.z = pack_half2(v[8 * j + 4], v[8 * j + 5]); u.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(myrow + ((j ^ ((lane >> 1) & 3)) << 4)) = u;      // 64B swizzle: addr[5:4] ^= addr[8:7]
          }
          if (prefetch_next) ptx::tmem_ld_32x32(t_row + c + 32, v);
          pend_col[pend++] = gc;
          const bool last = (c + 32 >= WCOLS) || (gc + 32 >= p.N);
          if (pend == G || last) {
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              uint8_t* gbase = wstage + (grp & 1) * (G * 4096);
              for (int i = 0; i < pend; ++i)
                ptx::tma_store_3d(&tmY, gbase + i * 4096, pend_col[i], m_blk * kBM + quad * 32, batch);
              ptx::bulk_commit();
            }
            pend = 0;
            ++grp;
          }
        } else if (p.tma_store) {
          uint8_t* myrow = box + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* dst = reinterpret_cast<float4*>(myrow + ((j ^ (lane & 7)) << 4));
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if constexpr (FULL) {
              if (stage_res) { const float4 r4 = *dst; o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w; }
            }
            *dst = o;
          }
          // v[] is consumed: start the TMEM load of the next chunk so that its latency hides behind the
          // proxy fence / TMA store issue below
          if (prefetch_next) ptx::tmem_ld_32x32(t_row + c + 32, v);
          pend_col[pend++] = gc;
          const bool last = (c + 32 >= WCOLS) || (gc + 32 >= p.N);
          if (pend == G || last) {                         // one proxy fence + commit per group of boxes
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              uint8_t* gbase = wstage + (grp & 1) * (G * 4096);
              for (int i = 0; i < pend; ++i)
                ptx::tma_store_3d(&tmY, gbase + i * 4096, pend_col[i], m_blk * kBM + quad * 32, batch);
              ptx::bulk_commit();
            }
            pend = 0;
            ++grp;
          }
        } else {
          if (row_ok) {
            for (int j = 0; j < 32; ++j) if (gc + j < p.N) yrow[gc + j] = v[j];
          }
          if (prefetch_next) ptx::tmem_ld_32x32(t_row + c + 32, v);
        }
      };
      const int nch = max(0, min(WCOLS / 32, (p.N - col0 + 31) / 32));    // live 32-column chunks of this warp's share
      {
        float v[32];
        if (nch > 0) ptx::tmem_ld_32x32(t_row, v);
#pragma unroll 1
        for (int ci = 0; ci < nch; ++ci) {
          ptx::tmem_ld_wait();
          emit_chunk(v, ci * 32, ci + 1 < nch);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) ptx::bulk_wait<0>();                    // all TMA stores of this warp have completed
  }

  ptx::tc_fence_before();
  if constexpr (CL == 2) ptx::cluster_sync(); else __syncthreads();   // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
std::atomic<int64_t> g_launches{0};

int init_driver(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return gf_set_error(GF_ERR_LAUNCH, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return gf_set_error(GF_ERR_DEVICE, "libgeoformer_sm100 needs an sm_100 (B200) device");
  g_num_sms = prop.multiProcessorCount;
  // the caller's current device is left as it is: every launch goes to the device of the stream it is given, and the
  // per-device kernel attributes are set lazily at the launch sites (GF_SMEM_OPTIN)
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr)
    return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return GF_OK;
}
int num_sms() { return g_num_sms; }
EncodeTiledFn encode_fn() { return g_encode; }

// 3-D K-major operand map: dims {K, rows, batches}; box {128 bytes of K, box_rows, 1}; 128B swizzle; OOB -> 0
int make_tmap(CUtensorMap* m, const void* base, int esize, int64_t k, int64_t rows, int64_t batches,
                     int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows) {
  if (!g_encode) return gf_set_error(GF_ERR_DRIVER, "gf_init() was not called");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)(row_stride_elems * esize), (cuuint64_t)(batch_stride_elems * esize)};
  if (batches == 1 && batch_stride_elems == 0) strides[1] = strides[0] * (cuuint64_t)rows;
  cuuint32_t box[3] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = g_encode(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled failed");
  return GF_OK;
}

// output map: dims {N, M, batches} fp32, box {32 cols, 32 rows, 1}, 128B swizzle (matches the epilogue staging)
int make_out_tmap(CUtensorMap* m, float* base, int64_t n, int64_t rows, int64_t batches, int64_t ld,
                         int64_t batch_stride) {
  if (!g_encode) return gf_set_error(GF_ERR_DRIVER, "gf_init() was not called");
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * 4), (cuuint64_t)(batch_stride * 4)};
  if (batches == 1 && batch_stride == 0) strides[1] = strides[0] * (cuuint64_t)rows;
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled(out) failed");
  return GF_OK;
}

// fp16 output map: box {32 cols, 32 rows, 1} of 64-byte rows, 64B swizzle (matches the OUT16 epilogue staging)
int make_out_tmap16(CUtensorMap* m, void* base, int64_t n, int64_t rows, int64_t batches, int64_t ld, int64_t batch_stride) {
  if (!g_encode) return gf_set_error(GF_ERR_DRIVER, "gf_init() was not called");
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)(ld * 2), (cuuint64_t)(batch_stride * 2)};
  if (batches == 1 && batch_stride == 0) strides[1] = strides[0] * (cuuint64_t)rows;
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled(out16) failed");
  return GF_OK;
}

template <int KIND, int BN, bool FULL, int CL, bool OUT16 = false>
static int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& ty,
                         const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, FULL>;
  auto kern = gemm_tc_kernel<KIND, BN, FULL, CL, OUT16>;
  GF_SMEM_OPTIN(kern, Cfg::kSmemBytes);
  const int64_t groups = (int64_t)p.batches * gf_cdiv(gf_cdiv(p.M, kBM), CL) * p.tiles_n;
  if (groups == 0) return GF_OK;
  const int64_t max_groups = g_num_sms / CL;
  const int grid = (int)(groups < max_groups ? groups : max_groups) * CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, ta2, tb, ty, p) != cudaSuccess) {
    cudaError_t e = cudaGetLastError();
    return gf_set_error(GF_ERR_LAUNCH, cudaGetErrorString(e));
  }
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

template <int KIND, int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& ty,
                       const GemmParams& p, cudaStream_t stream, bool out16 = false) {
  const bool full = (p.epi & GF_EPI_LN) || p.residual != nullptr || !p.tma_store;
  if (out16) {
    if (full) return gf_set_error(GF_ERR_ARG, "fp16 output is available with the lean epilogue only (no LN / residual)");
    if constexpr (KIND == 0) return launch_gemm_t<KIND, BN, false, 1, true>(ta, ta2, tb, ty, p, stream);
    else return gf_set_error(GF_ERR_ARG, "fp16 output needs fp32 (tf32) operands");
  }
  if (p.cluster == 2) {
    if (full) return launch_gemm_t<KIND, BN, true, 2>(ta, ta2, tb, ty, p, stream);
    return launch_gemm_t<KIND, BN, false, 2>(ta, ta2, tb, ty, p, stream);
  }
  if (full) return launch_gemm_t<KIND, BN, true, 1>(ta, ta2, tb, ty, p, stream);
  return launch_gemm_t<KIND, BN, false, 1>(ta, ta2, tb, ty, p, stream);
}

}  // namespace gf

using namespace gf;

// Y = epilogue([A | A2] W^T).  in_f16: A, A2 and W are fp16 (kind::f16, K % 64 == 0), else fp32 (kind::tf32, K % 32 == 0);
// out_f16: Y is fp16 (lean epilogue only), else fp32.
static int linear_impl(const void* A, const void* A2, const void* W, void* Y, int in_f16, int out_f16, int64_t M, int N,
                       int K1, int K2, int epi, int act_cols, const float* bias, const float* rowbias, int rowbias_group,
                       const float* gamma, const float* beta, const float* residual, const int* m_dev, gf_stream_t stream) {
  const int bke = in_f16 ? 64 : 32, esz = in_f16 ? 2 : 4;
  if (M < 0 || N <= 0 || K1 <= 0 || K2 < 0 || (N % 128) || (K1 % bke) || (K2 % bke) || M > 0x7fffff00LL)
    return gf_set_error(GF_ERR_ARG, "gf_linear: need N % 128 == 0, K % 32 == 0 (fp32) / K % 64 == 0 (fp16)");
  if ((epi & GF_EPI_LN) && !(N == 128 || N == 256)) return gf_set_error(GF_ERR_ARG, "gf_linear: LN epilogue needs N in {128,256}");
  if ((epi & GF_EPI_LN) && (!gamma || !beta)) return gf_set_error(GF_ERR_ARG, "gf_linear: LN epilogue needs gamma/beta");
  if (K2 > 0 && !A2) return gf_set_error(GF_ERR_ARG, "gf_linear: A2 missing");
  if (rowbias && rowbias_group <= 0) return gf_set_error(GF_ERR_ARG, "gf_linear: rowbias_group");
  if (M == 0) return GF_OK;
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap ta, ta2, tb;
  int rc;
  if ((rc = make_tmap(&ta, A, esz, K1, M, 1, K1, 0, kBM))) return rc;
  if (K2 > 0) { if ((rc = make_tmap(&ta2, A2, esz, K2, M, 1, K2, 0, kBM))) return rc; } else ta2 = ta;
  const int cl = (!in_f16 && !out_f16 && M > kBM && getenv("GF_CLUSTER2") != nullptr) ? 2 : 1;
  if ((rc = make_tmap(&tb, W, esz, K1 + K2, N, 1, K1 + K2, 0, BN / cl))) return rc;
  GemmParams p{};
  p.Y = reinterpret_cast<float*>(Y); p.ldy = N; p.y_batch_stride = 0; p.M = (int)M; p.N = N; p.batches = 1;
  p.kblocks1 = K1 / bke; p.kblocks2 = K2 / bke; p.epi = epi; p.act_cols = act_cols;
  p.bias = bias; p.rowbias = rowbias; p.rowbias_group = rowbias_group > 0 ? rowbias_group : 1;
  p.gamma = gamma; p.beta = beta; p.residual = residual; p.ldres = N; p.out_scale = 1.f; p.m_dev = m_dev;
  p.tiles_n = N / BN;
  p.cluster = cl;
  p.tma_store = 1;
  CUtensorMap ty;
  if (out_f16) { if ((rc = make_out_tmap16(&ty, Y, N, M, 1, N, 0))) return rc; }
  else if ((rc = make_out_tmap(&ty, reinterpret_cast<float*>(Y), N, M, 1, N, 0))) return rc;
  if (in_f16) {
    if (BN == 256) return launch_gemm<1, 256>(ta, ta2, tb, ty, p, (cudaStream_t)stream, out_f16 != 0);
    return launch_gemm<1, 128>(ta, ta2, tb, ty, p, (cudaStream_t)stream, out_f16 != 0);
  }
  if (BN == 256) return launch_gemm<0, 256>(ta, ta2, tb, ty, p, (cudaStream_t)stream, out_f16 != 0);
  return launch_gemm<0, 128>(ta, ta2, tb, ty, p, (cudaStream_t)stream, out_f16 != 0);
}

extern "C" int gf_linear_tf32(const float* A, const float* A2, const float* W, float* Y, int64_t M, int N, int K1,
                              int K2, int epi, int act_cols, const float* bias, const float* rowbias,
                              int rowbias_group, const float* gamma, const float* beta, const float* residual,
                              const int* m_dev, gf_stream_t stream) {
  return linear_impl(A, A2, W, Y, 0, 0, M, N, K1, K2, epi, act_cols, bias, rowbias, rowbias_group, gamma, beta, residual,
                     m_dev, stream);
}

extern "C" int gf_linear_mixed(const void* A, const void* A2, const void* W, void* Y, int in_f16, int out_f16, int64_t M,
                               int N, int K1, int K2, int epi, int act_cols, const float* bias, const float* rowbias,
                               int rowbias_group, const float* gamma, const float* beta, const float* residual,
                               const int* m_dev, gf_stream_t stream) {
  return linear_impl(A, A2, W, Y, in_f16, out_f16, M, N, K1, K2, epi, act_cols, bias, rowbias, rowbias_group, gamma, beta,
                     residual, m_dev, stream);
}

extern "C" int gf_similarity_f16x3(const void* a3, const void* b3, float* sim, int n, int l, int s, int c3,
                                   float out_scale, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || s <= 0 || c3 <= 0 || (c3 % 64)) return gf_set_error(GF_ERR_ARG, "gf_similarity_f16x3: c3 % 64 != 0");
  CUtensorMap ta, tb;
  int rc;
  constexpr int BN = 256;
  if ((rc = make_tmap(&ta, a3, 2, c3, l, n, c3, (int64_t)l * c3, kBM))) return rc;
  const int cl = (l > kBM && getenv("GF_CLUSTER2") != nullptr) ? 2 : 1;
  if ((rc = make_tmap(&tb, b3, 2, c3, s, n, c3, (int64_t)s * c3, BN / cl))) return rc;
  GemmParams p{};
  p.Y = sim; p.ldy = s; p.y_batch_stride = (int64_t)l * s; p.M = l; p.N = s; p.batches = n;
  p.kblocks1 = c3 / 64; p.kblocks2 = 0; p.epi = 0; p.act_cols = 0; p.rowbias_group = 1; p.ldres = s;
  p.out_scale = out_scale; p.tiles_n = gf_cdiv(s, BN);
  p.cluster = cl;
  p.tma_store = (s % 4 == 0) ? 1 : 0;
  CUtensorMap ty = ta;
  if (p.tma_store && (rc = make_out_tmap(&ty, sim, s, l, n, s, (int64_t)l * s))) return rc;
  return launch_gemm<1, 256>(ta, ta, tb, ty, p, (cudaStream_t)stream);
}

// Batched K-major GEMM with explicit strides (floats): Y[b] = out_scale * A[b] (M x K) * B[b]^T (N x K).
// Building block of the geo self-attention (Q K^T and P V per head, geo_attention.py:72-97).
extern "C" int gf_gemm_tf32_batched(const float* A, int64_t lda, int64_t a_batch_stride, const float* B, int64_t ldb,
                                    int64_t b_batch_stride, float* Y, int64_t ldy, int64_t y_batch_stride, int M,
                                    int N, int K, int batches, float out_scale, gf_stream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || batches <= 0 || (K % 32) || (lda % 4) || (ldb % 4) || (ldy % 4) ||
      (a_batch_stride % 4) || (b_batch_stride % 4) || (y_batch_stride % 4))
    return gf_set_error(GF_ERR_ARG, "gf_gemm_tf32_batched: K % 32 == 0 and 16-byte aligned strides required");
  const int BN = (N > 128) ? 256 : 128;
  CUtensorMap ta, tb, ty;
  int rc;
  if ((rc = make_tmap(&ta, A, 4, K, M, batches, lda, batches == 1 ? lda * M : a_batch_stride, kBM))) return rc;
  const int cl = (M > kBM && getenv("GF_CLUSTER2") != nullptr) ? 2 : 1;
  if ((rc = make_tmap(&tb, B, 4, K, N, batches, ldb, batches == 1 ? ldb * N : b_batch_stride, BN / cl))) return rc;
  if ((rc = make_out_tmap(&ty, Y, N, M, batches, ldy, batches == 1 ? ldy * M : y_batch_stride))) return rc;
  GemmParams p{};
  p.Y = Y; p.ldy = ldy; p.y_batch_stride = y_batch_stride; p.M = M; p.N = N; p.batches = batches;
  p.kblocks1 = K / 32; p.kblocks2 = 0; p.epi = 0; p.act_cols = 0; p.rowbias_group = 1; p.ldres = ldy;
  p.out_scale = out_scale; p.tiles_n = gf_cdiv(N, BN); p.tma_store = 1; p.cluster = cl;
  if (BN == 256) return launch_gemm<0, 256>(ta, ta, tb, ty, p, (cudaStream_t)stream);
  return launch_gemm<0, 128>(ta, ta, tb, ty, p, (cudaStream_t)stream);
}
