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
// (MLP-up x half, merge, MLP-up m1 half, MLP-down, then q|k|v of the NEXT tile).
// HBM traffic per token-layer: 512 B (x) [+ 512 B src] + 512 B (y) instead of ~6.9 KB for the unfused kernels (the residual
// re-read of x in the last epilogue is an L2 hit: the tile was fetched by TMA a few microseconds earlier).
//
// Software pipeline across tiles (r02): the x tile is only an MMA operand (the residual add reads x from global / L2 and
// y is stored straight from registers), so R1 is free as soon as the MLP-up x half has been issued and the NEXT tile's x
// lands there while the current tile is still in its merge / MLP phases; the MLP-down accumulator lives in the columns
// the merge accumulator vacated, so GEMM0 of the next tile is issued right behind the MLP-down GEMM and runs under the
// last epilogue (LayerNorm2 + residual + store).  Per tile the epilogue warps then wait only for the three short f16
// GEMMs in the middle of the chain.
//
// warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2..17: epilogue, FOUR warps per TMEM lane quadrant
// (thread == row, each warp a quarter of the columns): the epilogue phases are latency-bound chains (TMEM load -> math ->
// pack -> st.shared), so 16 warps in flight hide what 8 could not (r01: 8 warps, 17 us per tile; ncu 28.8 % tensor pipe).
// smem: R1 x tile (64 KB) | R2 src tile -> K,V fp16 -> m1 -> hidden (64 KB) | R3 Q fp16 -> message (32 KB) |
//       weight ring (48 KB).   TMEM: D0 q|k|v [0,384) -> D2 [0,256); D1 [384,512) -> D3 [384,512).
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
constexpr int R3_BYTES = 32768;
constexpr int RING = R3 + R3_BYTES;
constexpr int BARS = RING + STAGES * BLK;
constexpr int LNX = BARS + 256;                  // LayerNorm partial sums: 2 phases x 4 column quarters x 128 rows x float2
constexpr int SMEM_BYTES = LNX + 8192 + 1024;
constexpr int EPI_WARPS = 16, THREADS = 64 + 32 * EPI_WARPS;
constexpr int D0 = 0, D1 = 384, D2 = 0, D3 = 384;

struct Params {
  const float* x;        // residual stream (also the TMA source of the x tile)
  float* y;
  const float* gamma1; const float* beta1; const float* gamma2; const float* beta2;
  int64_t rows;          // windows * 25
  int tiles;
  int cross;             // src != x
  int debug;             // timing experiments only (env GF_FL_DEBUG): 2 = skip the attention math, 4 = skip E1..E3 math
                         // (results are garbage)
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float elu_plus1(float v) { return v > 0.f ? v + 1.f : ex2f(v * 1.4426950408889634f); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

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

// LayerNorm statistics of a 128-wide row split over the four warps of a TMEM lane quadrant (32 columns each):
// partial (sum, sum of squares) exchanged through shared memory
__device__ __forceinline__ void row_stats32(const float (&v)[32], float2* lnx, int part, int r, float& mean, float& rstd) {
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
  lnx[part * 128 + r] = make_float2(s, ss);
  epi_sync();
  const float2 p0 = lnx[r], p1 = lnx[128 + r], p2 = lnx[256 + r], p3 = lnx[384 + r];
  const float ts = (p0.x + p1.x) + (p2.x + p3.x), tss = (p0.y + p1.y) + (p2.y + p3.y);
  mean = ts * (1.f / 128.f);
  const float var = fmaxf(tss * (1.f / 128.f) - mean * mean, 0.f);
  rstd = rsqrtf(var + 1e-5f);
}
__device__ __forceinline__ void normalize32(float (&v)[32], float mean, float rstd, const float* __restrict__ gamma,
                                            const float* __restrict__ beta) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + j);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + j);
    v[4 * j] = (v[4 * j] - mean) * rstd * g4.x + b4.x;
    v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
    v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
    v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
  }
}
// a writer warp hands its shared-memory operand rows to the tensor core: writes -> async proxy, then ONE arrival per warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar);
}

__global__ void __launch_bounds__(THREADS, 1)
fine_layer_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS,
                  const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BARS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* x_full = empty_bar + STAGES;     // x tile of the current / next iteration has landed in R1
  uint64_t* s_full = x_full + 1;             // src tile (cross layers) has landed in R2
  uint64_t* d_full = s_full + 1;             // [4]: D0..D3 accumulators complete
  uint64_t* a_full = d_full + 4;             // [3]: message, m1, hidden written as A operands
  uint64_t* d0_free = a_full + 3;            // q|k|v copied out of TMEM: the MLP-up accumulator may overwrite them
  uint64_t* r1_free = d0_free + 1;           // tcgen05.commit after the MLP-up x half: R1 may take the next x tile
  uint64_t* r2_free = r1_free + 1;           // tcgen05.commit after the MLP-down GEMM: R2 may take the next src tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r2_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmS); ptx::prefetch_tmap(&tmW); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
      ptx::mbar_init(x_full, 1);
      ptx::mbar_init(s_full, 1);
      for (int i = 0; i < 4; ++i) ptx::mbar_init(&d_full[i], 1);
      for (int i = 0; i < 3; ++i) ptx::mbar_init(&a_full[i], EPI_WARPS);
      ptx::mbar_init(d0_free, 1);
      ptx::mbar_init(r1_free, 1);
      ptx::mbar_init(r2_free, 1);
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
      auto weights = [&](int b0, int b1) {               // blocks [b0, b1) of the layer's 30 through the ring
        for (int b = b0; b < b1; ++b) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], BLK);
          ptx::tma_load_3d(smem + RING + stage * BLK, &tmW, &full_bar[stage], 0, b * 128, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      auto load_x = [&](int t) {
        ptx::mbar_expect_tx(x_full, 4 * BLK);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(smem + R1 + kb * BLK, &tmX, x_full, kb * 32, t * ROWS, 0);
      };
      auto load_src = [&](int t) {
        ptx::mbar_expect_tx(s_full, 4 * BLK);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(smem + R2 + kb * BLK, &tmS, s_full, kb * 32, t * ROWS, 0);
      };
      // prologue: operands of GEMM0 of the first tile
      load_x(blockIdx.x);
      if (p.cross) load_src(blockIdx.x);
      weights(0, 12);
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int tn = t + gridDim.x;
        const bool next = tn < p.tiles;
        weights(12, 20);                                 // MLP-up, x half
        if (next) { ptx::mbar_wait(r1_free, it & 1); load_x(tn); }
        weights(20, 30);                                 // merge, MLP-up m1 half, MLP-down
        if (next) {
          weights(0, 4);                                 // q of the next tile (A = its x tile)
          if (p.cross) { ptx::mbar_wait(r2_free, it & 1); load_src(tn); }
          weights(4, 12);                                // k | v of the next tile
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // The whole warp runs this loop converged and one elected lane issues (operands stay in uniform registers; from
    // inside `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY loop of ~35 instructions,
    // which put ~100 cycles of issue latency per MMA on this kernel's serial GEMM chain).
    {
      constexpr uint32_t id_tf32 = ptx::umma_idesc(2, 128, 128);
      constexpr uint32_t id_f16 = ptx::umma_idesc(0, 128, 128);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, ptx::smem_addr(smem), 0);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t s_r1 = smem0 + R1, s_r2 = smem0 + R2, s_r3 = smem0 + R3, s_ring = smem0 + RING;
      const uint32_t bar0 = smem0 + BARS;                // full_bar[0]; the other barriers follow in declaration order
      auto bar_addr = [&](const uint64_t* b) -> uint32_t { return bar0 + (uint32_t)((const uint8_t*)b - (const uint8_t*)full_bar); };
      int stage = 0; uint32_t phase = 0;
      // one weight block against one resident A block: 4 MMAs (32 bytes of K each)
      auto step = [&](int kind, uint32_t a_addr, uint32_t d_col, bool first) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_lo = ptx::umma_desc_lo(a_addr), b_lo = ptx::umma_desc_lo(s_ring + stage * BLK);
        if (ptx::elect_one()) {
          if (kind == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_lo<0>(tb + d_col, a_lo + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, id_tf32, (first && k == 0) ? 0u : 1u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_lo<1>(tb + d_col, a_lo + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, id_f16, (first && k == 0) ? 0u : 1u);
          }
          ptx::umma_commit_addr(bar_addr(&empty_bar[stage]));
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      auto commit = [&](const uint64_t* b) { if (ptx::elect_one()) ptx::umma_commit_addr(bar_addr(b)); };
      // GEMM0 of iteration `i`: q from x, k|v from src (== x for a self layer)
      auto gemm0 = [&](int i) {
        ptx::mbar_wait(x_full, i & 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) step(0, s_r1 + kb * BLK, D0, kb == 0);
        if (p.cross) { ptx::mbar_wait(s_full, i & 1); ptx::tc_fence_after(); }
        const uint32_t a0 = p.cross ? s_r2 : s_r1;
        for (int nc = 1; nc < 3; ++nc)
          for (int kb = 0; kb < 4; ++kb) step(0, a0 + kb * BLK, D0 + nc * 128, kb == 0);
        commit(&d_full[0]);
      };
      gemm0(0);
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        // GEMM2, x half (tf32): runs under the attention math as soon as q|k|v have left TMEM
        ptx::mbar_wait(d0_free, par);
        ptx::tc_fence_after();
        for (int nc = 0; nc < 2; ++nc)
          for (int kb = 0; kb < 4; ++kb) step(0, s_r1 + kb * BLK, D2 + nc * 128, kb == 0);
        commit(r1_free);                                 // last read of the x tile: R1 may take the next one
        // GEMM1: merge(message)
        ptx::mbar_wait(&a_full[0], par);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) step(1, s_r3 + kb * BLK, D1, kb == 0);
        commit(&d_full[1]);
        // GEMM2, m1 half (f16) into the same accumulator
        ptx::mbar_wait(&a_full[1], par);
        ptx::tc_fence_after();
        for (int nc = 0; nc < 2; ++nc)
          for (int kb = 0; kb < 2; ++kb) step(1, s_r2 + 2 * BLK + kb * BLK, D2 + nc * 128, false);
        commit(&d_full[2]);
        // GEMM3: MLP down, into the columns of the (consumed) merge accumulator
        ptx::mbar_wait(&a_full[2], par);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) step(1, s_r2 + kb * BLK, D3, kb == 0);
        commit(&d_full[3]);
        commit(r2_free);                                 // last read of R2 (hidden): it may take the next src tile
        // GEMM0 of the next tile: D2 [0,256) was drained before a_full[2], [256,384) before d0_free -> runs under E3
        if (t + (int)gridDim.x < p.tiles) gemm0(it + 1);
      }
    }
  } else {
    // ------------------------------ epilogue: 16 warps, four per TMEM lane quadrant ------------------------------
    const int quad = warp & 3;                         // TMEM lane quadrant (hardware: warp id % 4)
    const int part = (warp - 2) >> 2;                  // which quarter of the columns this warp takes
    const int r = quad * 32 + lane;                    // row inside the tile
    const int et = threadIdx.x - 64;                   // 0..511
    const int ew = warp - 2;                           // 0..15
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
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
      // warp `part` takes the 32-column chunks part (q), 4 + part (k), 8 + part (v); the next chunk's TMEM load is in
      // flight while the current one is converted
      ptx::mbar_wait(&d_full[0], par);
      ptx::tc_fence_after();
      {
        float va[32], vb[32];
        ptx::tmem_ld_32x32(t_lane + D0 + 32 * part, va);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_lane + D0 + 32 * (4 + part), vb);
        uint4 u[4];
        // q chunk
#pragma unroll
        for (int j = 0; j < 32; ++j) va[j] = elu_plus1(va[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[j].x = pack2(va[8 * j], va[8 * j + 1]); u[j].y = pack2(va[8 * j + 2], va[8 * j + 3]);
          u[j].z = pack2(va[8 * j + 4], va[8 * j + 5]); u[j].w = pack2(va[8 * j + 6], va[8 * j + 7]);
        }
        {
          uint8_t* dst = r3 + (part >> 1) * BLK + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + ((((part & 1) * 4 + j) ^ sw7) << 4)) = u[j];
        }
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(t_lane + D0 + 32 * (8 + part), va);
        // k chunk
#pragma unroll
        for (int j = 0; j < 32; ++j) vb[j] = elu_plus1(vb[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[j].x = pack2(vb[8 * j], vb[8 * j + 1]); u[j].y = pack2(vb[8 * j + 2], vb[8 * j + 3]);
          u[j].z = pack2(vb[8 * j + 4], vb[8 * j + 5]); u[j].w = pack2(vb[8 * j + 6], vb[8 * j + 7]);
        }
        {
          uint8_t* dst = r2 + r * 256;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + (((part * 4 + j) ^ (r & 15)) << 4)) = u[j];
        }
        ptx::tmem_ld_wait();
        // v chunk (no feature map)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[j].x = pack2(va[8 * j], va[8 * j + 1]); u[j].y = pack2(va[8 * j + 2], va[8 * j + 3]);
          u[j].z = pack2(va[8 * j + 4], va[8 * j + 5]); u[j].w = pack2(va[8 * j + 6], va[8 * j + 7]);
        }
        {
          uint8_t* dst = r2 + 32768 + r * 256;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + (((part * 4 + j) ^ (r & 15)) << 4)) = u[j];
        }
      }
      ptx::tc_fence_before();
      epi_sync();
      if (et == 0) ptx::mbar_arrive(d0_free);
      // ---------------- E0b: linear attention per (window, head) on mma.sync; message overwrites Q in place ----------------
      if (!(p.debug & 2)) {
        const uint32_t k16 = ptx::smem_addr(r2), v16 = k16 + 32768, q16 = ptx::smem_addr(r3);
#pragma unroll 1
        for (int pair = ew; pair < WIN * HEADS; pair += EPI_WARPS) {
          const int w = pair >> 3, h = pair & 7;
          // C' = V_h^T K_h  ([d2][d1], two n-tiles over d1) and C'' = ones^T K_h (row 0 = Ksum); tokens = MMA K
          float cv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, co[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int tok0 = TOK * w + 16 * ks;
            uint32_t a[4], b[4];
            // rows 25.. of the 32-token span belong to the next window (or lie beyond the tile): their values are masked
            // below, so the addresses are clamped to this window's last token - no warp ever reads rows that another warp
            // owns (keeps compute-sanitizer racecheck clean)
            const int wlast = TOK * w + TOK - 1;
            const int arow = min(tok0 + lrow + ((lmat >> 1) << 3), wlast), achunk = 2 * h + (lmat & 1);
            ldsm_x4_trans(a, v16 + arow * 256 + ((achunk ^ (arow & 15)) << 4));
            const int brow = min(tok0 + lrow + ((lmat & 1) << 3), wlast), bchunk = 2 * h + (lmat >> 1);
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
            const int qrow = min(tok0 + lrow + ((lmat & 1) << 3), TOK * w + TOK - 1), qchunk = (h & 3) * 2 + (lmat >> 1);   // clamped: see above
            ldsm_x4(aq, q16 + (h >> 2) * BLK + qrow * 128 + ((qchunk ^ (qrow & 7)) << 4));
            float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, dn[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(o0, aq, bq[0][0], bq[0][1]);
            mma16816(o1, aq, bq[1][0], bq[1][1]);
            mma16816(dn, aq, bq[2][0], bq[2][1]);
            const float z_lo = 1.f / (__shfl_sync(0xffffffffu, dn[0], lane & ~3) + 1e-6f);
            const float z_hi = 1.f / (__shfl_sync(0xffffffffu, dn[2], lane & ~3) + 1e-6f);
            const int row_lo = tok0 + g, row_hi = row_lo + 8;
            __syncwarp();                                // (the message overwrites the Q rows this warp's ldmatrix just read)
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
      warp_arrive(&a_full[0], lane);
      // ---------------- E1: LayerNorm1(merge) -> m1 (fp16 A operand, R2 + 32 KB) ----------------
      ptx::mbar_wait(&d_full[1], par);
      ptx::tc_fence_after();
      {
        float v[32];
        ptx::tmem_ld_32x32(t_lane + D1 + 32 * part, v);
        ptx::tmem_ld_wait();
        float mean, rstd;
        row_stats32(v, lnx, part, r, mean, rstd);          // contains one epi_sync
        if (!(p.debug & 4)) {
          normalize32(v, mean, rstd, p.gamma1 + 32 * part, p.beta1 + 32 * part);
          uint8_t* dst = r2 + 2 * BLK + (part >> 1) * BLK + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack2(v[8 * j], v[8 * j + 1]); u.y = pack2(v[8 * j + 2], v[8 * j + 3]);
            u.z = pack2(v[8 * j + 4], v[8 * j + 5]); u.w = pack2(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(dst + ((((part & 1) * 4 + j) ^ sw7) << 4)) = u;
          }
        }
      }
      warp_arrive(&a_full[1], lane);
      // ---------------- E2: ReLU(MLP up) -> hidden (fp16 A operand, all of R2): chunks part and 4 + part ----------------
      ptx::mbar_wait(&d_full[2], par);
      ptx::tc_fence_after();
      {
        float va[32], vb[32];
        ptx::tmem_ld_32x32(t_lane + D2 + 32 * part, va);
        ptx::tmem_ld_32x32(t_lane + D2 + 32 * (4 + part), vb);
        ptx::tmem_ld_wait();
        if (!(p.debug & 4)) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = part + 4 * q;
            const float* v = q ? vb : va;
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
      }
      warp_arrive(&a_full[2], lane);
      // ---------------- E3: LayerNorm2(MLP down) + x -> y ----------------
      // Thread == row in TMEM, but global memory wants contiguous bytes per group of lanes: each warp transposes its
      // 32 rows x 16 columns through a private 2 KB box in R3 (the message there was consumed by GEMM1), twice, and then
      // reads the residual x from global (an L2 hit: TMA fetched this tile moments ago) and writes y with
      // lane = (row, 16-byte chunk): 64 contiguous bytes per 4 lanes.  The first round's residual loads are issued before
      // the wait for GEMM3 so that their latency hides behind it.
      uint8_t* box = r3 + ew * 2048;
      const int64_t grow0 = row0 + quad * 32;              // global row of the warp's row 0
      const int trow = lane >> 2, tch = lane & 3;           // store mapping: row 8 i + trow, chunk tch
      float4 xr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + trow;
        xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (quad * 32 + rr < ROWS && grow0 + rr < p.rows)
          xr[i] = __ldg(reinterpret_cast<const float4*>(p.x + (grow0 + rr) * C + 32 * part) + tch);
      }
      ptx::mbar_wait(&d_full[3], par);
      ptx::tc_fence_after();
      {
        float v[32];
        ptx::tmem_ld_32x32(t_lane + D3 + 32 * part, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();                            // D3 has been read: the next tile's merge GEMM may overwrite it
        float mean, rstd;
        row_stats32(v, lnx + 512, part, r, mean, rstd);    // contains one epi_sync
        if (!(p.debug & 4)) {
          normalize32(v, mean, rstd, p.gamma2 + 32 * part, p.beta2 + 32 * part);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(box + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                  make_float4(v[16 * c + 4 * j], v[16 * c + 4 * j + 1], v[16 * c + 4 * j + 2], v[16 * c + 4 * j + 3]);
            __syncwarp();
            float4 xn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = 8 * i + trow;
              const bool live = quad * 32 + rr < ROWS && grow0 + rr < p.rows;
              if (c == 0) {                                // residual of the second 16-column round
                xn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) xn[i] = __ldg(reinterpret_cast<const float4*>(p.x + (grow0 + rr) * C + 32 * part + 16) + tch);
              }
              const float4 y4 = *reinterpret_cast<const float4*>(box + rr * 64 + ((tch ^ ((rr >> 1) & 3)) << 4));
              if (live)
                *(reinterpret_cast<float4*>(p.y + (grow0 + rr) * C + 32 * part + 16 * c) + tch) =
                    make_float4(y4.x + xr[i].x, y4.y + xr[i].y, y4.z + xr[i].z, y4.w + xr[i].w);
            }
            if (c == 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i) xr[i] = xn[i];
            }
            __syncwarp();
          }
        }
      }
      epi_sync();      // other warps stage in rows of R3 that this warp's next Q write will overwrite
    }
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
  CUtensorMap tx, ts, tw;
  int rc;
  if ((rc = make_tmap(&tx, x, 4, fl::C, rows, 1, fl::C, 0, 128))) return rc;
  if ((rc = make_tmap(&ts, src, 4, fl::C, rows, 1, fl::C, 0, 128))) return rc;
  if ((rc = make_tmap(&tw, wpack, 4, 32, (int64_t)fl::NBLK * 128, 1, 32, 0, 128))) return rc;
  fl::Params p{};
  p.x = x; p.y = y; p.gamma1 = gamma1; p.beta1 = beta1; p.gamma2 = gamma2; p.beta2 = beta2;
  p.rows = rows; p.tiles = (int)((windows + fl::WIN - 1) / fl::WIN); p.cross = (src != x) ? 1 : 0;
  { const char* dbg = getenv("GF_FL_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  GF_SMEM_OPTIN(fl::fine_layer_kernel, fl::SMEM_BYTES);
  const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
  fl::fine_layer_kernel<<<grid, fl::THREADS, fl::SMEM_BYTES, (cudaStream_t)stream>>>(tx, ts, tw, p);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
