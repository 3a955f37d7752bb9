// Fused geo self-attention on tcgen05 (model/geo_transformer/transformer.py:111-124 + geo_attention.py:72-97):
// every token of an image attends to that image's anchor (RANSAC-inlier) tokens, full softmax, heads x 64.
//
// One CTA = 128 queries of one (sample, head).  Key tiles of 64 anchors stream through shared memory by TMA.
// Two passes over the keys avoid rescaling the TMEM accumulator:
//   pass 1:  S = Q K^T (kind::tf32, TMEM)  ->  row maxima m_i
//   pass 2:  S again (bit-identical), P = exp(S/sqrt(d) - m_i) -> swizzled smem, O += P V (A = P from smem,
//            B = V^T tile), row sums l_i;  out = O / l_i
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 softmax/epilogue (thread == query row).
// The score matrix never touches HBM (the 3-kernel path moved 4 x L x S_in x 4 B per head).
#include "common.cuh"
#include "ptx.cuh"

#include <atomic>

namespace gf {
extern std::atomic<int64_t> g_launches;

namespace fa {
constexpr int kBQ = 128;       // queries per CTA
constexpr int kBK = 64;        // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kRing = 3;       // K / V^T smem stages
constexpr int kQBytes = 2 * kBQ * 128;          // 2 k-blocks of 32 floats
constexpr int kKBytes = 2 * kBK * 128;          // 16 KB per key tile
constexpr int kVBytes = 2 * kD * 128;           // V^T tile: 2 k-blocks (32 keys each) x 64 dim rows
constexpr int kPBytes = 2 * kBQ * 128;          // P tile: 2 k-blocks (32 keys each) x 128 rows
constexpr int kSmem = kQBytes + kRing * kKBytes + kRing * kVBytes + 2 * kPBytes + 1024 + 256;
constexpr int kTmemCols = 256;                  // S: 2 x 64, O: 64
}  // namespace fa

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(192, 1)
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
        ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], 4);
        ptx::mbar_init(&p_full[i], 4); ptx::mbar_init(&p_empty[i], 1);
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
      ptx::tma_load_3d(sQ + kBQ * 128, &tmQ, q_full, h * kD + 32, q0, b);
      uint32_t gk = 0, gv = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < T; ++t) {
          {
            const int s = gk % kRing; const uint32_t ph = (gk / kRing) & 1; ++gk;
            ptx::mbar_wait(&k_empty[s], ph ^ 1);
            ptx::mbar_expect_tx(&k_full[s], kKBytes);
            ptx::tma_load_3d(sK + s * kKBytes, &tmK, &k_full[s], 0, t * kBK, hb);
            ptx::tma_load_3d(sK + s * kKBytes + kBK * 128, &tmK, &k_full[s], 32, t * kBK, hb);
          }
          if (pass == 1) {
            const int s = gv % kRing; const uint32_t ph = (gv / kRing) & 1; ++gv;
            ptx::mbar_wait(&v_empty[s], ph ^ 1);
            ptx::mbar_expect_tx(&v_full[s], kVBytes);
            ptx::tma_load_3d(sV + s * kVBytes, &tmV, &v_full[s], t * kBK, 0, hb);
            ptx::tma_load_3d(sV + s * kVBytes + kD * 128, &tmV, &v_full[s], t * kBK + 32, 0, hb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(2 /*tf32*/, kBQ, 64);
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
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ad = ptx::umma_desc_sw128(aQ + kb * kBQ * 128);
          const uint64_t bd = ptx::umma_desc_sw128(aK + kb * kBK * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma<0>(tmem_base + sb * 64, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
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
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ad = ptx::umma_desc_sw128(aP + kb * kBQ * 128);
            const uint64_t bd = ptx::umma_desc_sw128(aV + kb * kD * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma<0>(tmem_O, ad + 2 * k, bd + 2 * k, idesc, (t > 1 || kb || k) ? 1u : 0u);
          }
          ptx::umma_commit(&v_empty[vs]);
          ptx::umma_commit(&p_empty[pb]);
        }
      }
      ptx::umma_commit(o_full);
    }
  } else {
    // ------------------------------------ softmax / epilogue (warps 2..5) ------------------------------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                              // query row inside the CTA tile
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    uint32_t gs = 0, gp = 0;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) {                                  // pass 1: row maxima of the scaled logits
      const int sb = gs & 1; const uint32_t sph = (gs >> 1) & 1; ++gs;
      ptx::mbar_wait(&s_full[sb], sph);
      ptx::tc_fence_after();
      float v[64];
      ptx::tmem_ld_32x32(t_lane + sb * 64, v);
      ptx::tmem_ld_32x32(t_lane + sb * 64 + 32, v + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[sb]);
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
    // exp(s*scale - m*scale) = 2^(s*c2 - m*c2) with c2 = scale*log2(e): one FFMA + one MUFU.EX2 per element
    const float c2 = scale * 1.4426950408889634f;
    const float mc2 = m * c2;
    float lsum = 0.f;
    for (int t = 0; t < T; ++t) {                                  // pass 2: probabilities
      const int sb = gs & 1; const uint32_t sph = (gs >> 1) & 1; ++gs;
      ptx::mbar_wait(&s_full[sb], sph);
      ptx::tc_fence_after();
      float v[64];
      ptx::tmem_ld_32x32(t_lane + sb * 64, v);
      ptx::tmem_ld_32x32(t_lane + sb * 64 + 32, v + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[sb]);
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
      const int pb = gp & 1; const uint32_t pph = (gp >> 1) & 1; ++gp;
      ptx::mbar_wait(&p_empty[pb], pph ^ 1);                       // the P V MMAs that read this buffer retired
      uint8_t* prow = sP + pb * kPBytes + row * 128;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float4*>(prow + kb * kBQ * 128 + ((j ^ (row & 7)) << 4)) =
              make_float4(v[kb * 32 + 4 * j], v[kb * 32 + 4 * j + 1], v[kb * 32 + 4 * j + 2], v[kb * 32 + 4 * j + 3]);
        }
      }
      ptx::fence_proxy_async();                                    // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
    }
    // epilogue: O / l  -> swizzled 32x32 boxes (reusing P buffer 0) -> TMA store
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    float o[64];
    ptx::tmem_ld_32x32(tmem_O + (uint32_t(quad * 32) << 16), o);
    ptx::tmem_ld_32x32(tmem_O + (uint32_t(quad * 32) << 16) + 32, o + 32);
    ptx::tmem_ld_wait();
    const float inv = 1.f / lsum;
    uint8_t* wst = sP + quad * 8192;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<float4*>(wst + half * 4096 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(o[half * 32 + 4 * j] * inv, o[half * 32 + 4 * j + 1] * inv, o[half * 32 + 4 * j + 2] * inv,
                        o[half * 32 + 4 * j + 3] * inv);
      }
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_3d(&tmO, wst, h * kD, q0 + quad * 32, b);
      ptx::tma_store_3d(&tmO, wst + 4096, h * kD + 32, q0 + quad * 32, b);
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

// q: rows of the fused projection buffer (row stride ldq floats, head h at columns [h*64, h*64+64));
// kg [heads][n][s_pad][64], vt [heads][n][64][s_pad] from gf_gather_anchor_kv; out [n*l, heads*64].
extern "C" int gf_geo_self_attention_tc(const float* q, int ldq, const float* kg, const float* vt, float* out, int n,
                                        int l, int heads, int dim, int s_pad, const int* anchor_cnt, gf_stream_t stream) {
  if (n <= 0 || l <= 0 || heads <= 0 || dim != 64 || s_pad <= 0 || (s_pad % 4) || (ldq % 4))
    return gf_set_error(GF_ERR_ARG, "gf_geo_self_attention_tc: dim must be 64, s_pad % 4 == 0");
  CUtensorMap tq, tk, tv, to;
  int rc;
  const int c = heads * dim;
  if ((rc = make_tmap(&tq, q, 4, ldq, l, n, ldq, (int64_t)l * ldq, fa::kBQ))) return rc;
  if ((rc = make_tmap(&tk, kg, 4, dim, s_pad, (int64_t)heads * n, dim, (int64_t)s_pad * dim, fa::kBK))) return rc;
  if ((rc = make_tmap(&tv, vt, 4, s_pad, dim, (int64_t)heads * n, s_pad, (int64_t)dim * s_pad, fa::kD))) return rc;
  if ((rc = make_out_tmap(&to, out, c, l, n, c, (int64_t)l * c))) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(geo_flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::kSmem) != cudaSuccess)
      return gf_set_error(GF_ERR_LAUNCH, "cudaFuncSetAttribute(flash smem) failed");
    attr = true;
  }
  geo_flash_attn_kernel<<<dim3(gf_cdiv(l, fa::kBQ), heads, n), 192, fa::kSmem, (cudaStream_t)stream>>>(
      tq, tk, tv, to, out, anchor_cnt, n, l, heads, 1.f / sqrtf((float)dim));
  g_launches++;
  GF_CHECK_LAUNCH();
  return GF_OK;
}
