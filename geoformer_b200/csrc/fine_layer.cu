// One whole LoFTR encoder layer of the fine level (loftr_module/transformer.py:28-60 with linear_attention.py:33-49)
// as ONE persistent tcgen05 kernel: x/src [windows*25, 128] fp32 -> y [windows*25, 128] fp32.
//
// The fine level works on independent 25-token windows (one per coarse match), so nothing of a layer has to leave
// the SM: a CTA takes 5 windows (125 rows of a 128-row MMA tile) and chains
//   GEMM0  q|k|v = x Wq^T | src Wk^T | src Wv^T        (kind::tf32, A = the fp32 tile as TMA delivers it)
//   attention   Q,K,V -> fp16 smem; per (window, head) one warp computes V^T K and ones^T K (= Ksum) with mma.sync
//               (tokens are the MMA K dimension), whose accumulators ARE the B fragments of  Q (KV | Ksum):
//               message = Q KV / (Q Ksum + eps), written over Q in place in UMMA operand layout
//   GEMM1  merge (kind::f16, A = that message) -> LayerNorm1
//   GEMM2  MLP up: x W1x^T (tf32) + m1 W1m^T (f16) into one accumulator -> ReLU
//   GEMM3  MLP down (f16) -> LayerNorm2 -> + x (still resident) -> TMA store (boxes of 125 rows)
// Weights stream from L2 through a 3-stage TMA ring in a fixed order of 30 [128 x 128 B] blocks per tile
// (q|k|v, MLP-up x half, merge, MLP-up m1 half, MLP-down).
// HBM traffic per token-layer: 512 B (x) [+ 512 B src] + 512 B (y) instead of ~6.9 KB for the unfused kernels.
//
// warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2..9: epilogue (two warps per TMEM lane quadrant).
// smem: R1 x tile (64 KB) | R2 src tile -> K,V fp16 -> m1 -> hidden -> output staging (64 KB) |
//       R3 Q fp16 -> message (32 KB) | weight ring (48 KB).   TMEM: D0 q|k|v [0,384) -> D2 [0,256), D3 [256,384); D1 [384,512).
#include "common.cuh"
#include "ptx.cuh"

#include <cuda_fp16.h>

#include <atomic>
#include <cstdlib>

namespace gf {
extern std::atomic<int64_t> g_launches;

namespace fl {
constexpr int C = 128, TOK = 25, WIN = 5, ROWS = TOK * WIN, HEADS = 8, D = 16;
constexpr int BLK = 16384;                     // one operand block: 128 rows x 128 B
constexpr int NBLK = 30;                       // weight blocks streamed per tile
constexpr int STAGES = 3;
constexpr int R1 = 0, R2 = 65536, R3 = 131072;
constexpr int KV_WSTRIDE = HEADS * (D * D + D) * 4 + 16;     // bytes per window (+16: windows of a warp hit different banks)
constexpr int R3_BYTES = 44032;
constexpr int RING = R3 + R3_BYTES;
constexpr int BARS = RING + STAGES * BLK;
constexpr int LNX = BARS + 256;                  // LayerNorm partial sums: 2 phases x 2 halves x 128 rows x float2
constexpr int SMEM_BYTES = LNX + 4096 + 1024;
constexpr int D0 = 0, D1 = 384, D2 = 0, D3 = 256;

struct Params {
  float* y;
  const float* gamma1; const float* beta1; const float* gamma2; const float* beta2;
  int64_t rows;          // windows * 25
  int tiles;
  int cross;             // src != x
  int debug;             // timing experiments only (env GF_FL_DEBUG): 1 = weights streamed for the first tile only,
                         // 2 = skip the attention math, 4 = skip E1..E3 math (results are garbage)
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float elu_plus1(float v) { return v > 0.f ? v + 1.f : ex2f(v * 1.4426950408889634f); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// warp-level tensor-core helpers for the tiny per-(window, head) products of the linear attention
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// LayerNorm statistics of a 128-wide row split over the two warps of a TMEM lane quadrant (64 columns each):
// partial (sum, sum of squares) exchanged through shared memory
__device__ __forceinline__ void row_stats64(const float (&v)[64], float2* lnx, int half, int r, float& mean, float& rstd) {
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
  lnx[half * 128 + r] = make_float2(s, ss);
  epi_sync();
  const float2 o = lnx[(half ^ 1) * 128 + r];
  mean = (s + o.x) * (1.f / 128.f);
  const float var = fmaxf((ss + o.y) * (1.f / 128.f) - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-5f);
}
__device__ __forceinline__ void normalize64(float (&v)[64], float mean, float rstd, const float* __restrict__ gamma,
                                            const float* __restrict__ beta) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + j);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + j);
    v[4 * j] = (v[4 * j] - mean) * rstd * g4.x + b4.x;
    v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
    v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
    v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
  }
}

__global__ void __launch_bounds__(320, 1)
fine_layer_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BARS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* x_full = empty_bar + STAGES;
  uint64_t* tile_free = x_full + 1;
  uint64_t* d_full = tile_free + 1;          // [4]: D0..D3 accumulators complete
  uint64_t* a_full = d_full + 4;             // [3]: message, m1, hidden written as A operands
  uint64_t* d0_free = a_full + 3;            // q|k|v copied out of TMEM: the MLP-up accumulator may overwrite them
  uint64_t* r1_free = d0_free + 1;           // residual read done: the next x tile may land
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r1_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmS); ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmY); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
      ptx::mbar_init(x_full, 1);
      ptx::mbar_init(tile_free, 1);
      for (int i = 0; i < 4; ++i) ptx::mbar_init(&d_full[i], 1);
      for (int i = 0; i < 3; ++i) ptx::mbar_init(&a_full[i], 256);
      ptx::mbar_init(d0_free, 1);
      ptx::mbar_init(r1_free, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        if (it > 0) ptx::mbar_wait(r1_free, (it - 1) & 1);
        const int row0 = t * ROWS;
        ptx::mbar_expect_tx(x_full, p.cross ? 2 * 4 * BLK : 4 * BLK);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(smem + R1 + kb * BLK, &tmX, x_full, kb * 32, row0, 0);
        if (p.cross) {
          if (it > 0) ptx::mbar_wait(tile_free, (it - 1) & 1);      // R2 (output staging) has been read by the TMA store
          for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(smem + R2 + kb * BLK, &tmS, x_full, kb * 32, row0, 0);
        }
        if ((p.debug & 1) && it > 0) continue;
        for (int b = 0; b < NBLK; ++b) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], BLK);
          ptx::tma_load_3d(smem + RING + stage * BLK, &tmW, &full_bar[stage], 0, b * 128, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t id_tf32 = ptx::umma_idesc(2, 128, 128);
      constexpr uint32_t id_f16 = ptx::umma_idesc(0, 128, 128);
      const uint32_t s_r1 = ptx::smem_addr(smem + R1), s_r2 = ptx::smem_addr(smem + R2), s_r3 = ptx::smem_addr(smem + R3);
      const uint32_t s_ring = ptx::smem_addr(smem + RING);
      int stage = 0; uint32_t phase = 0;
      // one weight block against one resident A block: 4 MMAs (32 bytes of K each)
      int it = 0;
      auto step = [&](int kind, uint32_t a_addr, uint32_t d_col, bool first) {
        const bool stream = !((p.debug & 1) && it > 0);
        if (stream) ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t adesc = ptx::umma_desc_sw128(a_addr);
        const uint64_t bdesc = ptx::umma_desc_sw128(s_ring + stage * BLK);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kind == 0) ptx::umma<0>(tmem_base + d_col, adesc + 2 * k, bdesc + 2 * k, id_tf32, (first && k == 0) ? 0u : 1u);
          else           ptx::umma<1>(tmem_base + d_col, adesc + 2 * k, bdesc + 2 * k, id_f16, (first && k == 0) ? 0u : 1u);
        }
        if (stream) ptx::umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        ptx::mbar_wait(x_full, par);
        ptx::tc_fence_after();
        // GEMM0: q from x, k|v from src
        for (int nc = 0; nc < 3; ++nc) {
          const uint32_t a0 = (nc == 0 || !p.cross) ? s_r1 : s_r2;
          for (int kb = 0; kb < 4; ++kb) step(0, a0 + kb * BLK, D0 + nc * 128, kb == 0);
        }
        ptx::umma_commit(&d_full[0]);
        // GEMM2, x half (tf32): runs under the attention math as soon as q|k|v have left TMEM
        ptx::mbar_wait(d0_free, par);
        ptx::tc_fence_after();
        for (int nc = 0; nc < 2; ++nc)
          for (int kb = 0; kb < 4; ++kb) step(0, s_r1 + kb * BLK, D2 + nc * 128, kb == 0);
        // GEMM1: merge(message)
        ptx::mbar_wait(&a_full[0], par);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) step(1, s_r3 + kb * BLK, D1, kb == 0);
        ptx::umma_commit(&d_full[1]);
        // GEMM2, m1 half (f16) into the same accumulator
        ptx::mbar_wait(&a_full[1], par);
        ptx::tc_fence_after();
        for (int nc = 0; nc < 2; ++nc)
          for (int kb = 0; kb < 2; ++kb) step(1, s_r2 + 2 * BLK + kb * BLK, D2 + nc * 128, false);
        ptx::umma_commit(&d_full[2]);
        // GEMM3: MLP down
        ptx::mbar_wait(&a_full[2], par);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) step(1, s_r2 + kb * BLK, D3, kb == 0);
        ptx::umma_commit(&d_full[3]);
      }
    }
  } else {
    // ------------------------------ epilogue: 8 warps, two per TMEM lane quadrant ------------------------------
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                  // which half of the columns / heads this warp takes
    const int r = quad * 32 + lane;                    // row inside the tile
    const int et = threadIdx.x - 64;                   // 0..255
    const int ew = warp - 2;                           // 0..7
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    uint8_t* r1 = smem + R1;
    uint8_t* r2 = smem + R2;
    uint8_t* r3 = smem + R3;
    float2* lnx = reinterpret_cast<float2*>(smem + LNX);
    const int sw7 = r & 7;
    const int g = lane >> 2, tq = lane & 3;            // mma.sync fragment coordinates
    const int lrow = lane & 7, lmat = lane >> 3;       // ldmatrix: row inside the 8x8 matrix, matrix id
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int64_t row0 = (int64_t)t * ROWS;
      // ---------------- E0a: Q (elu+1) -> fp16 UMMA layout in R3; K (elu+1), V -> fp16 [row][256 B] in R2 ----------------
      ptx::mbar_wait(&d_full[0], par);
      ptx::tc_fence_after();
      epi_sync();                                      // thread 0 has seen the previous tile's TMA store finish reading R2
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        const int c = 2 * i + half;                    // 32-column chunk of q|k|v: 0..3 q, 4..7 k, 8..11 v
        float v[32];
        ptx::tmem_ld_32x32(t_lane + D0 + 32 * c, v);
        ptx::tmem_ld_wait();
        if (c < 8) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = elu_plus1(v[j]);
        }
        uint4 u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[j].x = pack2(v[8 * j], v[8 * j + 1]); u[j].y = pack2(v[8 * j + 2], v[8 * j + 3]);
          u[j].z = pack2(v[8 * j + 4], v[8 * j + 5]); u[j].w = pack2(v[8 * j + 6], v[8 * j + 7]);
        }
        if (c < 4) {
          uint8_t* dst = r3 + (c >> 1) * BLK + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + j) ^ sw7) << 4)) = u[j];
        } else {
          uint8_t* dst = r2 + ((c - 4) >> 2) * 32768 + r * 256;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + (((((c - 4) & 3) * 4 + j) ^ (r & 15)) << 4)) = u[j];
        }
      }
      ptx::tc_fence_before();
      epi_sync();
      if (et == 0) ptx::mbar_arrive(d0_free);
      // ---------------- E0b: linear attention per (window, head) on mma.sync; message overwrites Q in place ----------------
      if (!(p.debug & 2)) {
        const uint32_t k16 = ptx::smem_addr(r2), v16 = k16 + 32768, q16 = ptx::smem_addr(r3);
#pragma unroll 1
        for (int pi = 0; pi < 5; ++pi) {
          const int pair = ew * 5 + pi;
          const int w = pair >> 3, h = pair & 7;
          // C' = V_h^T K_h  ([d2][d1], two n-tiles over d1) and C'' = ones^T K_h (row 0 = Ksum); tokens = MMA K
          float cv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, co[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int tok0 = TOK * w + 16 * ks;
            uint32_t a[4], b[4];
            const int arow = tok0 + lrow + ((lmat >> 1) << 3), achunk = 2 * h + (lmat & 1);
            ldsm_x4_trans(a, v16 + arow * 256 + ((achunk ^ (arow & 15)) << 4));
            const int brow = tok0 + lrow + ((lmat & 1) << 3), bchunk = 2 * h + (lmat >> 1);
            ldsm_x4_trans(b, k16 + brow * 256 + ((bchunk ^ (brow & 15)) << 4));
            uint32_t o0 = (g == 0) ? 0x3C003C00u : 0u, o2 = o0;
            if (ks == 1) {                               // tokens 25.. of the 32-token span belong to the next window
              if (tq == 0) { a[2] &= 0xFFFFu; a[3] &= 0xFFFFu; o2 &= 0xFFFFu; } else { a[2] = 0u; a[3] = 0u; o2 = 0u; }
            }
            const uint32_t ao[4] = {o0, 0u, o2, 0u};
            mma16816(cv[0], a, b[0], b[1]); mma16816(cv[1], a, b[2], b[3]);
            mma16816(co[0], ao, b[0], b[1]); mma16816(co[1], ao, b[2], b[3]);
          }
          // the accumulators are exactly the B fragments of  out = Q_h KV  (k = d1, n = d2) and of  den = Q_h Ksum
          const uint32_t bq[3][2] = {{pack2(cv[0][0], cv[0][1]), pack2(cv[1][0], cv[1][1])},
                                     {pack2(cv[0][2], cv[0][3]), pack2(cv[1][2], cv[1][3])},
                                     {pack2(co[0][0], co[0][1]), pack2(co[1][0], co[1][1])}};
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int tok0 = TOK * w + 16 * mt;
            uint32_t aq[4];
            const int qrow = tok0 + lrow + ((lmat & 1) << 3), qchunk = (h & 3) * 2 + (lmat >> 1);
            ldsm_x4(aq, q16 + (h >> 2) * BLK + qrow * 128 + ((qchunk ^ (qrow & 7)) << 4));
            float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, dn[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(o0, aq, bq[0][0], bq[0][1]);
            mma16816(o1, aq, bq[1][0], bq[1][1]);
            mma16816(dn, aq, bq[2][0], bq[2][1]);
            const float z_lo = 1.f / (__shfl_sync(0xffffffffu, dn[0], lane & ~3) + 1e-6f);
            const float z_hi = 1.f / (__shfl_sync(0xffffffffu, dn[2], lane & ~3) + 1e-6f);
            const int row_lo = tok0 + g, row_hi = row_lo + 8;
            uint8_t* blk = r3 + (h >> 2) * BLK;
            const int c0 = (h & 3) * 2;
            *reinterpret_cast<uint32_t*>(blk + row_lo * 128 + ((c0 ^ (row_lo & 7)) << 4) + 4 * tq) = pack2(o0[0] * z_lo, o0[1] * z_lo);
            *reinterpret_cast<uint32_t*>(blk + row_lo * 128 + (((c0 + 1) ^ (row_lo & 7)) << 4) + 4 * tq) = pack2(o1[0] * z_lo, o1[1] * z_lo);
            if (mt == 0 || g == 0) {
              *reinterpret_cast<uint32_t*>(blk + row_hi * 128 + ((c0 ^ (row_hi & 7)) << 4) + 4 * tq) = pack2(o0[2] * z_hi, o0[3] * z_hi);
              *reinterpret_cast<uint32_t*>(blk + row_hi * 128 + (((c0 + 1) ^ (row_hi & 7)) << 4) + 4 * tq) = pack2(o1[2] * z_hi, o1[3] * z_hi);
            }
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&a_full[0]);
      // ---------------- E1: LayerNorm1(merge) -> m1 (fp16 A operand, R2 + 32 KB) ----------------
      ptx::mbar_wait(&d_full[1], par);
      ptx::tc_fence_after();
      {
        float v[64];
        ptx::tmem_ld_32x32(t_lane + D1 + 64 * half, v);
        ptx::tmem_ld_32x32(t_lane + D1 + 64 * half + 32, v + 32);
        ptx::tmem_ld_wait();
        float mean, rstd;
        row_stats64(v, lnx, half, r, mean, rstd);          // contains one epi_sync
        if (!(p.debug & 4)) {
          normalize64(v, mean, rstd, p.gamma1 + 64 * half, p.beta1 + 64 * half);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 u;
            u.x = pack2(v[8 * c], v[8 * c + 1]); u.y = pack2(v[8 * c + 2], v[8 * c + 3]);
            u.z = pack2(v[8 * c + 4], v[8 * c + 5]); u.w = pack2(v[8 * c + 6], v[8 * c + 7]);
            *reinterpret_cast<uint4*>(r2 + 2 * BLK + half * BLK + r * 128 + ((c ^ sw7) << 4)) = u;
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&a_full[1]);
      // ---------------- E2: ReLU(MLP up) -> hidden (fp16 A operand, all of R2) ----------------
      ptx::mbar_wait(&d_full[2], par);
      ptx::tc_fence_after();
      if (!(p.debug & 4)) {
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * half + i;
          float v[32];
          ptx::tmem_ld_32x32(t_lane + D2 + 32 * c, v);
          ptx::tmem_ld_wait();
          uint8_t* dst = r2 + (c >> 1) * BLK + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack2(fmaxf(v[8 * j], 0.f), fmaxf(v[8 * j + 1], 0.f)); u.y = pack2(fmaxf(v[8 * j + 2], 0.f), fmaxf(v[8 * j + 3], 0.f));
            u.z = pack2(fmaxf(v[8 * j + 4], 0.f), fmaxf(v[8 * j + 5], 0.f)); u.w = pack2(fmaxf(v[8 * j + 6], 0.f), fmaxf(v[8 * j + 7], 0.f));
            *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + j) ^ sw7) << 4)) = u;
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&a_full[2]);
      // ---------------- E3: LayerNorm2(MLP down) + x -> staging (R2) -> coalesced store ----------------
      ptx::mbar_wait(&d_full[3], par);
      ptx::tc_fence_after();
      {
        float v[64];
        ptx::tmem_ld_32x32(t_lane + D3 + 64 * half, v);
        ptx::tmem_ld_32x32(t_lane + D3 + 64 * half + 32, v + 32);
        ptx::tmem_ld_wait();
        float mean, rstd;
        row_stats64(v, lnx + 256, half, r, mean, rstd);    // contains one epi_sync (all GEMM3 reads of R2 are long done)
        if (!(p.debug & 4)) {
          normalize64(v, mean, rstd, p.gamma2 + 64 * half, p.beta2 + 64 * half);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * half + jj;                  // float4 index inside the 128-wide row
            const int off = (j >> 3) * BLK + r * 128 + (((j & 7) ^ sw7) << 4);     // same swizzled block layout for x and y
            const float4 x4 = *reinterpret_cast<const float4*>(r1 + off);
            *reinterpret_cast<float4*>(r2 + off) =
                make_float4(v[4 * jj] + x4.x, v[4 * jj + 1] + x4.y, v[4 * jj + 2] + x4.z, v[4 * jj + 3] + x4.w);
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      epi_sync();
      if (et == 0) {
        ptx::mbar_arrive(r1_free);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_store_3d(&tmY, r2 + kb * BLK, kb * 32, (int)row0, 0);   // box = 125 rows
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();
        ptx::mbar_arrive(tile_free);
      }
    }
    if (et == 0) ptx::bulk_wait<0>();                  // output stores have landed
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fl
}  // namespace gf

using namespace gf;

// x, src: [windows*25, 128] fp32 (src == x for a self layer); wpack: the layer's 30 weight blocks [30][128][128 B] in
// streaming order (see engine.pack_fine_layer); y: [windows*25, 128] fp32 (must not alias x or src).
extern "C" int gf_fine_layer(const float* x, const float* src, const void* wpack, const float* gamma1, const float* beta1,
                             const float* gamma2, const float* beta2, float* y, int64_t windows, gf_stream_t stream) {
  if (windows < 0 || !x || !src || !wpack || !y || y == x || y == src) return gf_set_error(GF_ERR_ARG, "gf_fine_layer: bad arguments");
  if (windows == 0) return GF_OK;
  const int64_t rows = windows * fl::TOK;
  if (rows > 0x7fffff00LL) return gf_set_error(GF_ERR_ARG, "gf_fine_layer: too many rows");
  CUtensorMap tx, ts, tw, ty;
  int rc;
  if ((rc = make_tmap(&tx, x, 4, fl::C, rows, 1, fl::C, 0, 128))) return rc;
  if ((rc = make_tmap(&ts, src, 4, fl::C, rows, 1, fl::C, 0, 128))) return rc;
  if ((rc = make_tmap(&tw, wpack, 4, 32, (int64_t)fl::NBLK * 128, 1, 32, 0, 128))) return rc;
  if ((rc = make_tmap(&ty, y, 4, fl::C, rows, 1, fl::C, 0, fl::ROWS))) return rc;      // store boxes of 125 rows: tiles do not overlap
  fl::Params p{};
  p.y = y; p.gamma1 = gamma1; p.beta1 = beta1; p.gamma2 = gamma2; p.beta2 = beta2;
  p.rows = rows; p.tiles = (int)((windows + fl::WIN - 1) / fl::WIN); p.cross = (src != x) ? 1 : 0;
  { const char* dbg = getenv("GF_FL_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  GF_SMEM_OPTIN(fl::fine_layer_kernel, fl::SMEM_BYTES);
  const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
  fl::fine_layer_kernel<<<grid, 320, fl::SMEM_BYTES, (cudaStream_t)stream>>>(tx, ts, tw, ty, p);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
