// 3x3 (pad 1) and 1x1 convolutions, stride 1 or 2, as a tcgen05 implicit GEMM (backbone: resnet_fpn.py:32-40, 70-82, 100-118;
// SURVEY.md §8f rank 1).  NHWC fp16 activations (fp32 accumulation in TMEM), BN folded into weights + bias.  fp16, not
// bf16: same tensor-core rate, 8x smaller storage rounding (2^-11 vs 2^-8) through ~20 chained layers - measured on the
// 480x640 reference goldens the bf16 backbone alone cost 10-40 % of the final match-set identity, fp16 costs ~1-3 %
// (profiles/r02_parity_precision_probe.txt).  Outputs saturate at +-65504 instead of overflowing to inf.
//
//   Y[b,y,x,co] = act( bias[co] + sum_{dy,dx,ci} X[b,y+dy-1,x+dx-1,ci] * W[co,dy,dx,ci]  (+ R[b,y,x,co]) )
//
// GEMM view: M = 128 output pixels (an 8 x 16 patch), N = all output channels (128 / 208 / 256), K = 9 taps x Cin.
// The im2col never exists: for tap (dy,dx) and 64-channel block cb the A tile is ONE 4-D TMA box
// {64 ch, 16 px, 8 rows, 1} at (cb*64, x0+dx-1, y0+dy-1, b); TMA zero-fills the padding halo and the channel tail,
// and lands the box as 128 rows x 128 B with the 128B swizzle == a K-major UMMA operand.
// CTA pairs (r02).  ncu on the single-CTA kernel (profiles/r02_conv_ncu_before_pair.txt): tensor pipe 49-52 % (74 % for the
// K = N = 256 layer) while the MMA warp never waited for operands and the epilogue warps waited for the MMAs - neither the
// L2 feed (2-CTA weight multicast: no change), nor the epilogue (8 instead of 4 warps: no change), nor the issue rate (lean
// uniform-register issue loop: no change) was the limit.  What fits all layers is the shared-memory port: with M = 128
// per MMA every weight byte is written once (TMA) and read once per 128 output pixels, e.g. N = 256: 48 KB written + 48 KB
// read per 64-channel k-block = 768 cycles at 128 B/clk for 512 cycles of MMA.  Hence tcgen05.mma.cta_group::2: the two
// CTAs of a cluster take neighbouring pixel tiles (M = 256 across the pair); each holds its own 128 pixels of A and only
// HALF of the weight rows, so weight traffic through each SM's shared memory is halved (N = 256: 64 KB per k-block =
// 512 cycles).  The leader CTA (rank 0) issues every MMA; both CTAs' TMA loads count on the leader's full barrier; MMA
// completion is multicast to both CTAs' empty / tmem_full barriers; the peer's epilogue releases its accumulator with a
// remote arrive on the leader's tmem_empty barrier.
// Same warp-specialised persistent structure as gemm_tc.cu (TMA producer / single-thread MMA issuer / 8 epilogue
// warps, double-buffered TMEM accumulator); epilogue fuses bias, residual add, ReLU / LeakyReLU, fp16 pack and
// writes 4-D TMA boxes {64 ch, 16 px, 2 rows}.
#include "common.cuh"
#include "ptx.cuh"

#include <atomic>
#include <cstdlib>
#include <cuda_fp16.h>

namespace gf {
extern std::atomic<int64_t> g_launches;

struct ConvParams {
  const float* bias;                 // [cout_p] fp32
  const __half* residual;            // NHWC [b,h,w,cout_p] or null
  int batch, h, w, cout_p;
  int kbc;                           // 64-channel k-blocks per tap
  int last_k;                        // MMAs (16 channels each) that carry real channels in the last k-block of a tap
  int tail;                          // last_k == 1: the last k-block of a tap is loaded as a 32-byte-wide box (16 channels)
                                     // through the *_t tensor maps instead of a mostly zero-filled 128-byte one
  int taps, stride, pad;             // 9 taps / pad 1 (3x3) or 1 tap / pad 0 (1x1); stride 1 or 2; h, w are OUTPUT sizes
  int tiles_x, tiles_y;
  int act;                           // 0 none, 1 relu, 2 leaky relu (0.01)
};

// MT = pixel tiles (8 x 16 each, stacked vertically) per k-block: with MT == 2 one weight box feeds two accumulators,
// i.e. 48 KB instead of 64 KB of operands per 2 x 128 pixels (the N = 128 layers are bound by the L2 -> smem feed).
template <int BN, int MT = 1> struct ConvCfg {
  static constexpr int kStageA = MT * 128 * 128;
  static constexpr int kStageB = (BN / 2) * 128;                  // this CTA's half of the weight rows
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN <= 128) ? (MT == 2 ? 4 : 6) : 5;
  static constexpr int kAccStride = (BN <= 128) ? MT * 128 : 256;      // TMEM column offset of accumulator buffer 1
  static constexpr int kTmemCols = (BN <= 128 && MT == 1) ? 256 : 512;
  static constexpr int kStaging = 8 * 4096;                       // per epilogue warp: one box of 32 px x 128 B
  static constexpr int kSmem = kStages * kStage + kStaging + 1024 + 256;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_addr(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  // saturating: a value beyond fp16's range becomes +-65504, never inf (which would turn into NaN downstream)
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BN, int MT = 1>
__global__ void __launch_bounds__(320, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmXt,
                  const __grid_constant__ CUtensorMap tmWt, const ConvParams p) {
  using Cfg = ConvCfg<BN, MT>;
  static_assert(MT == 1 || BN <= 128, "two pixel tiles per step need 2 x 2 x BN TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStaging);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.batch * tiles_per_img;
  const int kblocks = p.taps * p.kbc;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmY);
    if (p.tail) { ptx::prefetch_tmap(&tmXt); ptx::prefetch_tmap(&tmWt); }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < Cfg::kStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
      // tmem_empty is only used in the leader CTA: 8 epilogue warps of each CTA of the pair arrive there
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 16); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc2(tmem_slot, Cfg::kTmemCols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                     // the peer's barriers are initialised before anything is multicast into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t rank = ptx::cluster_ctarank();
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int n_pairs = (total_tiles + 1) >> 1;
  // tile of this CTA in iteration `pair`: 2 * pair + rank; the odd tail tile's partner recomputes the last tile (no store)
  auto tile_of = [&](int pair, bool& ghost) -> int {
    const int t = 2 * pair + (int)rank;
    ghost = t >= total_tiles;
    return ghost ? total_tiles - 1 : t;
  };

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues (uniform-register operands)
    {
      int stage = 0; uint32_t phase = 0;
      for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        bool ghost;
        const int t = tile_of(pair, ghost);
        const int b = t / tiles_per_img, r = t - b * tiles_per_img;
        const int y0 = (r / p.tiles_x) * (8 * MT), x0 = (r % p.tiles_x) * 16;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int tap = kb / p.kbc, cb = kb - tap * p.kbc;
          const int dy = tap / 3, dx = tap - dy * 3;       // single-tap (1x1) convolutions: tap == 0, pad == 0
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);    // the pair's MMAs have consumed this stage (commit is multicast)
          uint8_t* sa = smem + stage * Cfg::kStage;
          const bool tail_blk = p.tail && cb == p.kbc - 1;
          if (ptx::elect_one()) {
          // the LEADER's full barrier counts the bytes of both CTAs (each: own A box(es) + own half of the weight box)
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (tail_blk ? MT * 4096 + (BN / 2) * 32 : Cfg::kStage));
          if (tail_blk) {
            // channel tail (e.g. channels 192..199 of 200): 32-byte-wide boxes, a quarter of the bytes of a full k-block
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_4d_2sm(sa + mt * 4096, &tmXt, &full_bar[stage], cb * 64, x0 * p.stride + dx - p.pad,
                                   (y0 + 8 * mt) * p.stride + dy - p.pad, b);
            ptx::tma_load_3d_2sm(sa + Cfg::kStageA, &tmWt, &full_bar[stage], kb * 64, (int)rank * (BN / 2), 0);
          } else {
            // the input map traverses W and H with element stride == conv stride: the box is 16 x 8 OUTPUT pixels
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ptx::tma_load_4d_2sm(sa + mt * 16384, &tmX, &full_bar[stage], cb * 64, x0 * p.stride + dx - p.pad,
                                   (y0 + 8 * mt) * p.stride + dy - p.pad, b);
            ptx::tma_load_3d_2sm(sa + Cfg::kStageA, &tmW, &full_bar[stage], kb * 64, (int)rank * (BN / 2), 0);
          }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: only in the leader CTA of the pair.  The WHOLE warp runs this loop converged and one elected lane
    // issues: every operand (descriptors, TMEM address, barrier address) is then provably warp-uniform and lives in
    // uniform registers (issuing from inside `if (lane == 0)` made ptxas wrap each tcgen05.mma / commit in an
    // ELECT + R2UR + BRA.U.ANY loop of ~35 SASS instructions).
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0 /*f16*/, 256, BN);      // M = 256 over the CTA pair
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = __shfl_sync(0xffffffffu, ptx::smem_addr(smem), 0);
      const uint32_t fb0 = smem0 + (uint32_t)(reinterpret_cast<uint8_t*>(full_bar) - smem);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // both CTAs' epilogues have drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tb + acc * Cfg::kAccStride;
        int cb = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);              // both CTAs' boxes of this stage have landed
          ptx::tc_fence_after();
          const uint32_t sa = smem0 + stage * Cfg::kStage;
          // channel tail (e.g. 200 = 3 x 64 + 8): the zero-filled part of the last k-block is not multiplied
          const bool last = cb == p.kbc - 1;
          const int nk = last ? p.last_k : 4;
          if (++cb == p.kbc) cb = 0;
          const uint32_t b_lo = ptx::umma_desc_lo(sa + Cfg::kStageA);
          const uint32_t a_lo0 = ptx::umma_desc_lo(sa);
          if (last && p.tail) {                            // one K = 16 step from the 32-byte-wide boxes
            if (ptx::elect_one()) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                ptx::umma2_f16_lo(d_tmem + mt * 128, a_lo0 + mt * 256, b_lo, ptx::kDescHiSw32, idesc, kb ? 1u : 0u);
            }
          } else if (nk == 4) {                            // hot path: no per-MMA predicate, everything uniform
            if (ptx::elect_one()) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma2_f16_lo(d_tmem + mt * 128, a_lo0 + mt * 1024 + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, idesc, (kb | k) ? 1u : 0u);
            }
          } else {                                         // channel tail of 32 or 48 channels (not in this network)
            for (int mt = 0; mt < MT; ++mt)
              for (int k = 0; k < nk; ++k)
                if (ptx::elect_one())
                  ptx::umma2_f16_lo(d_tmem + mt * 128, a_lo0 + mt * 1024 + 2 * k, b_lo + 2 * k, ptx::kDescHiSw128, idesc, (kb | k) ? 1u : 0u);
          }
          if (ptx::elect_one()) ptx::umma2_commit_mcast_addr(fb0 + (Cfg::kStages + stage) * 8, (uint16_t)0x3);   // empty_bar[stage], both CTAs
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma2_commit_mcast_addr(fb0 + (2 * Cfg::kStages + acc) * 8, (uint16_t)0x3);    // tmem_full[acc], both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // 8 epilogue warps, two per TMEM lane quadrant, each taking every other 64-channel chunk: with 4 warps the epilogue
    // (~3.6 us per chunk: residual staging, tcgen05.ld, bias / act / pack, TMA store) took longer than the MMAs of every
    // layer with K < 256 (r02: N = 208 layers ran at 72 % of the sustained tensor peak, N = 128 at 74 %)
    const int quad = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    uint8_t* box = staging + (warp - 2) * 4096;
    int acc = 0; uint32_t acc_phase = 0;
    for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
      bool ghost;
      const int t = tile_of(pair, ghost);
      const int b = t / tiles_per_img, r = t - b * tiles_per_img;
      const int ty0 = (r / p.tiles_x) * (8 * MT), x0 = (r % p.tiles_x) * 16;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int y0 = ty0 + 8 * mt;
      if (y0 >= p.h || ghost) break;
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * Cfg::kAccStride + mt * 128;
      for (int c0 = ehalf * 64; c0 < BN; c0 += 128) {
        if (c0 >= p.cout_p) break;
        if (lane == 0) ptx::bulk_wait_read<0>();          // the previous store of this warp has read the box
        __syncwarp();
        if (p.residual) {
          // coalesced: 4 pixels x 128 B per instruction (pixel = it*4 + lane/8, 16-byte chunk = lane%8)
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int px = it * 4 + (lane >> 3), ch = lane & 7;
            const int y = y0 + quad * 2 + (px >> 4), x = x0 + (px & 15);
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (y < p.h && x < p.w && c0 + ch * 8 < p.cout_p)
              val = __ldg(reinterpret_cast<const uint4*>(p.residual + (((int64_t)b * p.h + y) * p.w + x) * p.cout_p + c0 + ch * 8));
            *reinterpret_cast<uint4*>(box + px * 128 + ((ch ^ (px & 7)) << 4)) = val;
          }
          __syncwarp();
        }
        float v[64];
        ptx::tmem_ld_32x32(t_row + c0, v);
        ptx::tmem_ld_32x32(t_row + c0 + 32, v + 32);
        ptx::tmem_ld_wait();
        if (c0 + 64 <= p.cout_p) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        } else {
          for (int j = 0; j < 64; ++j) if (c0 + j < p.cout_p) v[j] += __ldg(p.bias + c0 + j);
        }
        uint8_t* myrow = box + lane * 128;
        if (p.residual) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 r4 = *reinterpret_cast<const uint4*>(myrow + ((j ^ (lane & 7)) << 4));
            const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2 h2 = *reinterpret_cast<const __half2*>(&w4[q]);
              v[8 * j + 2 * q] += __low2float(h2);
              v[8 * j + 2 * q + 1] += __high2float(h2);
            }
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_f16(v[8 * j], v[8 * j + 1]); o.y = pack_f16(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack_f16(v[8 * j + 4], v[8 * j + 5]); o.w = pack_f16(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(myrow + ((j ^ (lane & 7)) << 4)) = o;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmY, box, c0, x0, y0 + quad * 2, b);
          ptx::bulk_commit();
        }
      }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);     // the pair's MMA issuer lives in the leader CTA
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                     // no CTA leaves (or frees tensor memory) while its peer may still use it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

// out = lateral + bilinear_upsample(src -> (h, w), align_corners=True)   (FPN top-down merge, resnet_fpn.py:108-115)
// NHWC fp16; one thread per 8 channels (16 B).
__global__ void upsample_add_f16_kernel(const uint4* __restrict__ lateral, const uint4* __restrict__ src,
                                         uint4* __restrict__ out, int b, int h, int w, int hs, int ws, int c8,
                                         float ry, float rx) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)b * h * w * c8;
  if (idx >= total) return;
  const int ch = (int)(idx % c8);
  int64_t r = idx / c8;
  const int x = (int)(r % w); r /= w;
  const int y = (int)(r % h);
  const int n = (int)(r / h);
  const float fy = ry * y, fx = rx * x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const uint4* s = src + (int64_t)n * hs * ws * c8;
  const uint4 a = __ldg(s + ((int64_t)y0 * ws + x0) * c8 + ch), bq = __ldg(s + ((int64_t)y0 * ws + x1) * c8 + ch);
  const uint4 cq = __ldg(s + ((int64_t)y1 * ws + x0) * c8 + ch), d = __ldg(s + ((int64_t)y1 * ws + x1) * c8 + ch);
  const uint4 l = __ldg(lateral + idx);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w}, cw[4] = {cq.x, cq.y, cq.z, cq.w},
                 dw[4] = {d.x, d.y, d.z, d.w}, lw[4] = {l.x, l.y, l.z, l.w};
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 a2 = *reinterpret_cast<const __half2*>(&aw[q]), b2 = *reinterpret_cast<const __half2*>(&bw[q]);
    const __half2 c2 = *reinterpret_cast<const __half2*>(&cw[q]), d2 = *reinterpret_cast<const __half2*>(&dw[q]);
    const __half2 l2 = *reinterpret_cast<const __half2*>(&lw[q]);
    const float lo = __low2float(l2) + w00 * __low2float(a2) + w01 * __low2float(b2) + w10 * __low2float(c2) + w11 * __low2float(d2);
    const float hi = __high2float(l2) + w00 * __high2float(a2) + w01 * __high2float(b2) + w10 * __high2float(c2) + w11 * __high2float(d2);
    o[q] = pack_f16(lo, hi);
  }
  out[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// Stem: 7x7 / stride 2 / pad 3 convolution of the 1-channel image + folded BN + ReLU (resnet_fpn.py:58-60,102),
// fp32 image in, NHWC fp16 out.  C_in = 1 makes this a 49-tap FFMA kernel (no tensor-core shape): CTA = 8 x 32
// output pixels x 128 channels, thread = 8 consecutive pixels x 16 channels (128 accumulators), input patch and
// weights in shared memory; per kernel row 21 input LDS + 28 weight LDS.128 feed 896 FFMA.
// Weights: [49 taps][128 channels] fp32 (natural channel order); thread (cg, j) reads the float4 of channels
// j*32 + cg*4 .. +3, so the 8 channel groups of a quarter-warp cover 128 contiguous bytes (conflict-free).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_conv7x7_kernel(const float* __restrict__ img, const float* __restrict__ wperm, const float* __restrict__ bias,
                    __half* __restrict__ out, int h, int w, int ho, int wo) {
  __shared__ float patch[21][72];
  __shared__ __align__(16) float ws[49 * 128];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  for (int e = tid; e < 49 * 128; e += 256) ws[e] = wperm[e];
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  const float* im = img + (int64_t)b * h * w;
  for (int e = tid; e < 21 * 69; e += 256) {
    const int r = e / 69, c = e - r * 69;
    const int y = iy0 + r, x = ix0 + c;
    patch[r][c] = (y >= 0 && y < h && x >= 0 && x < w) ? im[(int64_t)y * w + x] : 0.f;
  }
  __syncthreads();
  const int cg = tid & 7, pg = tid >> 3;
  const int prow = pg >> 2, pcol = (pg & 3) * 8;           // 8 pixels: row prow, columns pcol .. pcol+7 of the tile
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    float in[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) in[i] = patch[prow * 2 + ky][pcol * 2 + i];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float4* wp = reinterpret_cast<const float4*>(ws + (ky * 7 + kx) * 128) + cg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = wp[j * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = in[2 * i + kx];
          acc[i][4 * j] = fmaf(x, w4.x, acc[i][4 * j]); acc[i][4 * j + 1] = fmaf(x, w4.y, acc[i][4 * j + 1]);
          acc[i][4 * j + 2] = fmaf(x, w4.z, acc[i][4 * j + 2]); acc[i][4 * j + 3] = fmaf(x, w4.w, acc[i][4 * j + 3]);
        }
      }
    }
  }
  const int oy = oy0 + prow;
  if (oy < ho) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + j * 32 + cg * 4));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ox = ox0 + pcol + i;
        if (ox < wo) {
          uint2 o;
          o.x = pack_f16(fmaxf(acc[i][4 * j] + b4.x, 0.f), fmaxf(acc[i][4 * j + 1] + b4.y, 0.f));
          o.y = pack_f16(fmaxf(acc[i][4 * j + 2] + b4.z, 0.f), fmaxf(acc[i][4 * j + 3] + b4.w, 0.f));
          *reinterpret_cast<uint2*>(out + (((int64_t)b * ho + oy) * wo + ox) * 128 + j * 32 + cg * 4) = o;
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

// NHWC fp16 map: dims {C, W, H, B}; box {64, box_w, box_h, 1}; 128B swizzle; OOB -> 0 on load, clipped on store
// `stride` > 1: W and H are traversed with that element stride; the box then spans box_w*stride x box_h*stride source
// elements and delivers box_w x box_h of them (cuTensorMapEncodeTiled: ceil(boxDim / elementStride) per dimension).
// box_c = 64 (128-byte rows, 128B swizzle) or 16 (32-byte rows, 32B swizzle: the channel-tail boxes)
static int make_nhwc_tmap(CUtensorMap* m, const void* base, int c, int w, int h, int b, int box_w, int box_h, int stride = 1,
                          int box_c = 64) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return gf_set_error(GF_ERR_DRIVER, "gf_init() was not called");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * w, (cuuint64_t)c * 2 * w * h};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled(nhwc) failed");
  return GF_OK;
}

// weight map for the channel-tail boxes: dims {K, rows}; box {16 elements = 32 B, box_rows}; 32B swizzle
static int make_wtail_tmap(CUtensorMap* m, const void* base, int64_t k, int64_t rows, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return gf_set_error(GF_ERR_DRIVER, "gf_init() was not called");
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)k * 2, (cuuint64_t)k * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return gf_set_error(GF_ERR_DRIVER, "cuTensorMapEncodeTiled(weight tail) failed");
  return GF_OK;
}

template <int BN, int MT = 1>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& txt,
                       const CUtensorMap& twt, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, MT>;
  auto kern = conv3x3_tc_kernel<BN, MT>;
  GF_SMEM_OPTIN(kern, Cfg::kSmem);
  const int tiles = p.batch * p.tiles_x * p.tiles_y;
  const int pairs = (tiles + 1) / 2, max_clusters = num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (pairs < max_clusters ? pairs : max_clusters));
  cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = Cfg::kSmem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, txt, twt, p) != cudaSuccess) return gf_set_error(GF_ERR_LAUNCH, cudaGetErrorString(cudaGetLastError()));
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

}  // namespace gf

using namespace gf;

// x NHWC fp16 [b,h,w,cin_p]; wt fp16 [cout_p][taps][cin_k] (taps = ksize^2, cin_k = cin_p rounded up to 64, zero padded);
// bias fp32 [cout_p]; residual / y NHWC fp16 [b,ho,wo,cout_p], ho = (h - 1) / stride + 1.  ksize 3 (pad 1) or 1 (pad 0),
// stride 1 or 2; cin_p, cout_p multiples of 8; cout_p <= 256.
extern "C" int gf_conv_f16(const void* x, const void* wt, const float* bias, const void* residual, void* y, int batch,
                            int h, int w, int cin_p, int cout_p, int cin_k, int ksize, int stride, int act,
                            gf_stream_t stream) {
  if (batch <= 0 || h <= 0 || w <= 0 || (cin_p % 8) || (cout_p % 8) || cout_p > 256 || (cin_k % 64) || cin_k < cin_p ||
      !(ksize == 1 || ksize == 3) || !(stride == 1 || stride == 2) || bias == nullptr)
    return gf_set_error(GF_ERR_ARG, "gf_conv_f16: channels % 8, cout <= 256, cin_k % 64 == 0, ksize in {1,3}, stride in {1,2}");
  const int taps = ksize * ksize;
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const int BN = cout_p <= 128 ? 128 : (cout_p <= 208 ? 208 : 256);
  CUtensorMap tx, tw, ty;
  int rc;
  if ((rc = make_nhwc_tmap(&tx, x, cin_p, w, h, batch, 16, 8, stride))) return rc;
  if ((rc = make_nhwc_tmap(&ty, y, cout_p, wo, ho, batch, 16, 2))) return rc;
  if ((rc = make_tmap(&tw, wt, 2, taps * (int64_t)cin_k, cout_p, 1, taps * (int64_t)cin_k, 0, BN / 2))) return rc;   // half a weight box per CTA of the cluster
  ConvParams p{};
  p.bias = bias; p.residual = (const __half*)residual; p.batch = batch; p.h = ho; p.w = wo; p.cout_p = cout_p;
  p.kbc = cin_k / 64; p.taps = taps; p.stride = stride; p.pad = ksize / 2;
  p.last_k = (cin_p - (p.kbc - 1) * 64 + 15) / 16;
  if (p.last_k < 1 || p.last_k > 4) return gf_set_error(GF_ERR_ARG, "gf_conv_f16: cin_k must be cin_p rounded up to 64");
  p.tail = (p.last_k == 1 && getenv("GF_CONV_NOTAIL") == nullptr) ? 1 : 0;
  CUtensorMap txt = tx, twt = tw;           // only dereferenced when p.tail
  if (p.tail) {
    if ((rc = make_nhwc_tmap(&txt, x, cin_p, w, h, batch, 16, 8, stride, 16))) return rc;
    if ((rc = make_wtail_tmap(&twt, wt, taps * (int64_t)cin_k, cout_p, BN / 2))) return rc;
  }
  p.tiles_x = gf_cdiv(wo, 16); p.tiles_y = gf_cdiv(ho, 8); p.act = act;
  if (BN == 128 && ho >= 32 && getenv("GF_CONV_MT1") == nullptr) {     // two stacked pixel tiles per weight box
    p.tiles_y = gf_cdiv(ho, 16);
    return launch_conv<128, 2>(tx, tw, ty, txt, twt, p, (cudaStream_t)stream);
  }
  if (BN == 128) return launch_conv<128>(tx, tw, ty, txt, twt, p, (cudaStream_t)stream);
  if (BN == 208) return launch_conv<208>(tx, tw, ty, txt, twt, p, (cudaStream_t)stream);
  return launch_conv<256>(tx, tw, ty, txt, twt, p, (cudaStream_t)stream);
}

extern "C" int gf_conv3x3_f16(const void* x, const void* wt, const float* bias, const void* residual, void* y,
                               int batch, int h, int w, int cin_p, int cout_p, int cin_k, int act, gf_stream_t stream) {
  return gf_conv_f16(x, wt, bias, residual, y, batch, h, w, cin_p, cout_p, cin_k, 3, 1, act, stream);
}

extern "C" int gf_upsample_add_f16(const void* lateral, const void* src, void* out, int batch, int h, int w, int hs,
                                    int ws, int c, gf_stream_t stream) {
  if (batch <= 0 || h <= 1 || w <= 1 || hs <= 0 || ws <= 0 || (c % 8)) return gf_set_error(GF_ERR_ARG, "gf_upsample_add_f16: bad shape");
  const int64_t total = (int64_t)batch * h * w * (c / 8);
  const float ry = (float)(hs - 1) / (float)(h - 1), rx = (float)(ws - 1) / (float)(w - 1);
  upsample_add_f16_kernel<<<gf_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)lateral, (const uint4*)src, (uint4*)out, batch, h, w, hs, ws, c / 8, ry, rx);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}

// img fp32 [b,1,h,w]; wperm fp32 [49 taps][128 channels]; bias fp32 [128]; out NHWC fp16 [b, ceil(h/2), ceil(w/2), 128].
extern "C" int gf_stem_conv7x7_f16(const float* img, const float* wperm, const float* bias, void* out, int batch, int h,
                                    int w, gf_stream_t stream) {
  if (batch <= 0 || h <= 0 || w <= 0) return gf_set_error(GF_ERR_ARG, "gf_stem_conv7x7_f16: bad shape");
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  stem_conv7x7_kernel<<<dim3(gf_cdiv(wo, 32), gf_cdiv(ho, 8), batch), 256, 0, (cudaStream_t)stream>>>(
      img, wperm, bias, (__half*)out, h, w, ho, wo);
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
