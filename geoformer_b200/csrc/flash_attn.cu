// Fused geo self-attention on tcgen05 (model/geo_transformer/transformer.py:111-124 + geo_attention.py:72-97):
// every token of an image attends to that image's anchor (RANSAC-inlier) tokens, full softmax, heads x 64.
//
// One CTA = 128 queries of one (sample, head).  Key tiles of 64 anchors stream through shared memory by TMA.
// Operands are fp16 (kind::f16, fp32 accumulate): 11-bit significands are at least as accurate as tf32, and they
// halve the shared-memory operand traffic per MMA — with N = 64 tiles the tf32 version was bound by the 6 KB of
// smem operand reads per 32-cycle instruction (192 B/clk > the 128 B/clk smem port), not by the softmax warps.
// Two passes over the keys avoid rescaling the TMEM accumulator:
//   pass 1:  S = Q K^T (TMEM)  ->  row maxima m_i
//   pass 2:  S again (bit-identical), P = exp(S/sqrt(d) - m_i) -> swizzled smem, O += P V (A = P from smem,
//            B = V^T tile), row sums l_i;  out = O / l_i
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 softmax/epilogue (thread == query row).
// The score matrix never touches HBM (the 3-kernel path moved 4 x L x S_in x 4 B per head).
#include "common.cuh"
#include "ptx.cuh"

#include <atomic>
#include <cuda_fp16.h>

namespace gf {
extern std::atomic<int64_t> g_launches;

namespace fa {
constexpr int kBQ = 128;       // queries per CTA
constexpr int kBK = 64;        // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kRing = 4;       // K / V^T smem stages
constexpr int kQBytes = kBQ * 128;              // 128 rows x 64 halves: one 128-byte k-block
constexpr int kKBytes = kBK * 128;              // 64 keys x 64 dims (fp16): 8 KB
constexpr int kVBytes = kD * 128;               // V^T tile: 64 dim rows x 64 keys (fp16): 8 KB
constexpr int kPBytes = kBQ * 128;              // P tile: 128 rows x 64 keys (fp16): 16 KB
constexpr int kSoftWarps = 8;   // two softmax warpgroups (4 warps each) ping-ponging over the score tiles
constexpr int kThreads = 64 + 32 * kSoftWarps;
constexpr int kSmem = kQBytes + kRing * kKBytes + kRing * kVBytes + 2 * kPBytes + 1024 + 256 + 2048;
constexpr int kTmemCols = 256;                  // S: 2 x 64, O: 64
}  // namespace fa

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(fa::kThreads, 1)
geo_flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                      float* __restrict__ out, const int* __restrict__ anchor_cnt, int n_samples, int l, int heads,
                      float scale) {
  using namespace fa;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kRing * kKBytes;
  uint8_t* sP = sV + kRing * kVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // kRing
  uint64_t* k_empty = k_full + kRing;
  uint64_t* v_full = k_empty + kRing;
  uint64_t* v_empty = v_full + kRing;
  uint64_t* s_full = v_empty + kRing;      // 2
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xchg = reinterpret_cast<float*>(bars + 32);    // [2 halves][128 rows] row max, then [2][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z;
  const int c = heads * kD;
  const int cnt = anchor_cnt[b];
  if (cnt <= 0) {   // no anchors: the layer is skipped for this sample (caller restores the features); emit zeros
    for (int e = threadIdx.x; e < kBQ * (kD / 4); e += blockDim.x) {
      const int r = e / (kD / 4), q = e - r * (kD / 4);
      if (q0 + r < l) reinterpret_cast<float4*>(out + ((int64_t)b * l + q0 + r) * c + h * kD)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int T = (cnt + kBK - 1) / kBK;
  const int hb = h * n_samples + b;        // batch coordinate of the gathered K / V^T tensors

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1);
      for (int i = 0; i < kRing; ++i) {
        ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1);
        ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], kSoftWarps / 2);
        ptx::mbar_init(&p_full[i], kSoftWarps / 2); ptx::mbar_init(&p_empty[i], 1);
      }
      ptx::mbar_init(o_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, kQBytes);
      ptx::tma_load_3d(sQ, &tmQ, q_full, h * kD, q0, b);
      uint32_t gk = 0, gv = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < T; ++t) {
          {
            const int s = gk % kRing; const uint32_t ph = (gk / kRing) & 1; ++gk;
            ptx::mbar_wait(&k_empty[s], ph ^ 1);
            ptx::mbar_expect_tx(&k_full[s], kKBytes);
            ptx::tma_load_3d(sK + s * kKBytes, &tmK, &k_full[s], 0, t * kBK, hb);
          }
          if (pass == 1) {
            const int s = gv % kRing; const uint32_t ph = (gv / kRing) & 1; ++gv;
            ptx::mbar_wait(&v_empty[s], ph ^ 1);
            ptx::mbar_expect_tx(&v_full[s], kVBytes);
            ptx::tma_load_3d(sV + s * kVBytes, &tmV, &v_full[s], t * kBK, 0, hb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0 /*f16*/, kBQ, 64);
      ptx::mbar_wait(q_full, 0);
      ptx::tc_fence_after();
      const uint32_t aQ = ptx::smem_addr(sQ);
      uint32_t gk = 0, gv = 0, gs = 0, gp = 0;
      auto issue_S = [&]() {
        const int ks = gk % kRing; const uint32_t kph = (gk / kRing) & 1; ++gk;
        const int sb = gs & 1; const uint32_t sph = (gs >> 1) & 1; ++gs;
        ptx::mbar_wait(&k_full[ks], kph);
        ptx::mbar_wait(&s_empty[sb], sph ^ 1);
        ptx::tc_fence_after();
        const uint32_t aK = ptx::smem_addr(sK + ks * kKBytes);
        {
          const uint64_t ad = ptx::umma_desc_sw128(aQ);
          const uint64_t bd = ptx::umma_desc_sw128(aK);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma<1>(tmem_base + sb * 64, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
        }
        ptx::umma_commit(&k_empty[ks]);
        ptx::umma_commit(&s_full[sb]);
      };
      for (int t = 0; t < T; ++t) issue_S();                       // pass 1: scores only
      for (int t = 0; t <= T; ++t) {                               // pass 2: scores of tile t, P V of tile t-1
        if (t < T) issue_S();
        if (t > 0) {
          const int pb = gp & 1; const uint32_t pph = (gp >> 1) & 1; ++gp;
          const int vs = gv % kRing; const uint32_t vph = (gv / kRing) & 1; ++gv;
          ptx::mbar_wait(&p_full[pb], pph);
          ptx::mbar_wait(&v_full[vs], vph);
          ptx::tc_fence_after();
          const uint32_t aP = ptx::smem_addr(sP + pb * kPBytes);
          const uint32_t aV = ptx::smem_addr(sV + vs * kVBytes);
          {
            const uint64_t ad = ptx::umma_desc_sw128(aP);
            const uint64_t bd = ptx::umma_desc_sw128(aV);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma<1>(tmem_O, ad + 2 * k, bd + 2 * k, idesc, (t > 1 || k) ? 1u : 0u);
          }
          ptx::umma_commit(&v_empty[vs]);
          ptx::umma_commit(&p_empty[pb]);
        }
      }
      ptx::umma_commit(o_full);
    }
  } else {
    // ------------------------------------ softmax / epilogue (warps 2..9) ------------------------------------
    // Two warpgroups ping-pong over the score tiles (FA3-style): group g owns TMEM score buffer g, i.e. every global
    // score tile gs with (gs & 1) == g.  The serial chain of one tile (barrier wait -> tcgen05.ld -> exp -> smem P ->
    // proxy fence -> arrive) of one group overlaps with the other group's.
    const int quad = warp & 3;                                     // TMEM lane quadrant (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                              // warpgroup id; also the 32-column half in the epilogue
    const int row = quad * 32 + lane;                              // query row inside the CTA tile
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16) + half * 64;     // this group's score buffer
    float m = -INFINITY;
    for (int gs = half; gs < T; gs += 2) {                         // pass 1: row maxima of the raw logits
      const int t = gs;
      ptx::mbar_wait(&s_full[half], (gs >> 1) & 1);
      ptx::tc_fence_after();
      float v[64];
      ptx::tmem_ld_32x32(t_lane, v);
      ptx::tmem_ld_32x32(t_lane + 32, v + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[half]);
      const int live = min(kBK, cnt - t * kBK);
      if (live != kBK) {
#pragma unroll
        for (int j = 0; j < 64; ++j) if (j >= live) v[j] = -INFINITY;
      }
      float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];            // 4 independent chains
#pragma unroll
      for (int j = 4; j < 64; j += 4) {
        m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]);
      }
      m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    // the two groups exchange their maxima (named barrier 1 over the 256 softmax threads)
    xchg[half * kBQ + row] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(m, xchg[(half ^ 1) * kBQ + row]);
    // exp(s*scale - m*scale) = 2^(s*c2 - m*c2) with c2 = scale*log2(e): one FFMA + one MUFU.EX2 per element
    const float c2 = scale * 1.4426950408889634f;
    const float mc2 = m * c2;
    float lsum = 0.f;
    // pass 2: global score tiles gs = T .. 2T-1; first one of this group: smallest gs >= T with (gs & 1) == half
    for (int gs = T + ((T ^ half) & 1); gs < 2 * T; gs += 2) {
      const int t = gs - T;
      ptx::mbar_wait(&s_full[half], (gs >> 1) & 1);
      ptx::tc_fence_after();
      float v[64];
      ptx::tmem_ld_32x32(t_lane, v);
      ptx::tmem_ld_32x32(t_lane + 32, v + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[half]);
      const int live = min(kBK, cnt - t * kBK);
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = ex2_approx(fmaf(v[j], c2, -mc2));
      if (live != kBK) {
#pragma unroll
        for (int j = 0; j < 64; ++j) if (j >= live) v[j] = 0.f;
      }
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j += 4) { a0 += v[j]; a1 += v[j + 1]; a2 += v[j + 2]; a3 += v[j + 3]; }
      lsum += (a0 + a1) + (a2 + a3);
      const int pb = t & 1;
      ptx::mbar_wait(&p_empty[pb], ((t >> 1) & 1) ^ 1);            // the P V MMAs that read this buffer retired
      uint8_t* prow = sP + pb * kPBytes + row * 128;               // 64 probabilities as fp16 = one 128-byte swizzled row
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 o;
        o.x = pack_h2(v[8 * j], v[8 * j + 1]); o.y = pack_h2(v[8 * j + 2], v[8 * j + 3]);
        o.z = pack_h2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_h2(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(prow + ((j ^ (row & 7)) << 4)) = o;
      }
      ptx::fence_proxy_async();                                    // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
    }
    // row sums: add the other half's
    xchg[2 * kBQ + half * kBQ + row] = lsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    lsum += xchg[2 * kBQ + (half ^ 1) * kBQ + row];
    // epilogue: this warp's 32 output columns of O / l -> swizzled 32x32 box (reusing the P buffers) -> TMA store
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    float o[32];
    ptx::tmem_ld_32x32(tmem_O + (uint32_t(quad * 32) << 16) + half * 32, o);
    ptx::tmem_ld_wait();
    const float inv = 1.f / lsum;
    uint8_t* wst = sP + (warp - 2) * 4096;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<float4*>(wst + lane * 128 + ((j ^ (lane & 7)) << 4)) =
          make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_3d(&tmO, wst, h * kD + half * 32, q0 + quad * 32, b);
      ptx::bulk_commit();
      ptx::bulk_wait<0>();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, fa::kTmemCols);
  }
}

}  // namespace gf

using namespace gf;

// fp16 operands from gf_gather_anchor_kv_f16: q16 [n*l, heads*64] (head h at columns [h*64, h*64+64)),
// kg [heads][n][s_pad][64], vt [heads][n][64][s_pad]; out fp32 [n*l, heads*64].
extern "C" int gf_geo_self_attention_tc(const void* q16, int ldq, const void* kg, const void* vt, float* out, int n, int l,
                                        int heads, int dim, int s_pad, const int* anchor_cnt, gf_stream_t stream) {
  const int c = heads * dim;
  if (n <= 0 || l <= 0 || heads <= 0 || dim != 64 || s_pad <= 0 || (s_pad % 8) || ldq < c || (ldq % 8))
    return gf_set_error(GF_ERR_ARG, "gf_geo_self_attention_tc: dim must be 64, s_pad % 8 == 0, ldq % 8 == 0");
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap(&tq, q16, 2, c, l, n, ldq, (int64_t)l * ldq, fa::kBQ))) return rc;
  if ((rc = make_tmap(&tk, kg, 2, dim, s_pad, (int64_t)heads * n, dim, (int64_t)s_pad * dim, fa::kBK))) return rc;
  if ((rc = make_tmap(&tv, vt, 2, s_pad, dim, (int64_t)heads * n, s_pad, (int64_t)dim * s_pad, fa::kD))) return rc;
  if ((rc = make_out_tmap(&to, out, c, l, n, c, (int64_t)l * c))) return rc;
  GF_SMEM_OPTIN(geo_flash_attn_kernel, fa::kSmem);
  geo_flash_attn_kernel<<<dim3(gf_cdiv(l, fa::kBQ), heads, n), fa::kThreads, fa::kSmem, (cudaStream_t)stream>>>(
      tq, tk, tv, to, out, anchor_cnt, n, l, heads, 1.f / sqrtf((float)dim));
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
