// FlashAttention-2 forward for B200 (sm_100a):  O = softmax(Q K^T * scale) V,  [B,H,N,D] fp16, non-causal.
//
// One CTA owns 256 query rows of one (batch, head): two 128-row Q tiles that ping-pong on the tensor core.
//   warp 0        TMA producer: Q tiles once, then a ring of K and V tiles (128 keys each), 128B/64B-swizzled smem
//   warp 1        MMA issuer (one thread):  S_i = Q_i K_j^T   (tcgen05.mma SS, fp32 accumulate, 128x128 into TMEM)
//                                            O_i += P_ij V_j   (tcgen05.mma TS: P read from TMEM, V MN-major smem)
//   warp 2        TMEM allocator (512 columns: S0 S1 | O0 O1)
//   warps 4..7    softmax warpgroup for Q tile 0, one thread per query row: tcgen05.ld the S row, online softmax
//   warps 8..11   softmax warpgroup for Q tile 1       (row max / exp2 / row sum in fp32 registers), P written back
//                 as fp16 over the S columns with tcgen05.st; O is rescaled in TMEM only when the row max grew by
//                 more than 2^8 (lazy rescaling); the same threads normalise and store O at the end (TMA store).
// While one warpgroup does softmax on its tile the tensor core works for the other tile.
//
// Replaces kernels/flash-attn/mma/basic/flash_attn_mma_share_qkv.cu:L45-709 (kernel) / L711-886 (launcher) and the
// other flash_attn_mma_stages_* variants (same math, different Ampere smem strategies).  Numerics follow the
// reference where it matters: scores scaled by 1/sqrt(D), softmax statistics in fp32, P rounded to fp16 before PV
// (share_qkv.cu:L431-478); the MMA accumulators are fp32 here (the reference default is fp16 accumulate).
#include <cmath>

#include "abi_common.cuh"
#include "ptx.cuh"

namespace b200k {

template <int D_, int STAGES_, bool V_DN_ = false>
struct Fa2Cfg {
  static constexpr int D = D_;
  static constexpr bool V_DN = V_DN_;  // V passed transposed as [B,H,D,N] (the reference's *_swizzle_qkv entry points)
  static constexpr int STAGES = STAGES_;
  static constexpr int CW = (D % 64 == 0) ? 64 : 32;  // width of one smem chunk along D (elements)
  static constexpr int NCH = D / CW;
  static constexpr int ROWB = CW * 2;                  // bytes per smem row = swizzle span (128 or 64)
  static constexpr uint32_t SWZ_MODE = (ROWB == 128) ? 2u : 4u;
  static constexpr int BR = 128;                       // query rows per tile (= TMEM lanes)
  static constexpr int BC = 128;                       // keys per KV tile
  static constexpr int CHUNK_BYTES = 128 * ROWB;       // one TMA box: 128 rows x CW elements
  static constexpr int TILE_BYTES = NCH * CHUNK_BYTES; // a 128 x D tile
  static constexpr int BAR_BYTES = 1024;
  static constexpr int SMEM_BYTES = 1024 + BAR_BYTES + 2 * TILE_BYTES + 2 * STAGES * TILE_BYTES;
  static constexpr int S_COL0 = 0, S_COL1 = 128;
  static constexpr int O_COL0 = 256, O_COL1 = 256 + D;
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 384;
  static_assert(D % 32 == 0 && D >= 32 && D <= 128, "head dim");
  static_assert(SMEM_BYTES <= 232448, "smem");
};

// Lazy-rescale threshold in log2 units: P may grow up to 2^8 before the running max is moved (fp16 P and fp32 sums
// have ample range), so the O accumulator is touched only on the first few KV tiles of a row.
constexpr float kRescaleThreshold = 8.0f;

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
fa2_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int N,
                       float scale_log2) {
  constexpr int D = Cfg::D, STAGES = Cfg::STAGES, NCH = Cfg::NCH, CW = Cfg::CW, ROWB = Cfg::ROWB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  // barriers
  const uint32_t bar_q_full = base;                       // 2
  const uint32_t bar_k_full = base + 16;                  // STAGES
  const uint32_t bar_k_empty = bar_k_full + 8 * STAGES;   // STAGES
  const uint32_t bar_v_full = bar_k_empty + 8 * STAGES;   // STAGES
  const uint32_t bar_v_empty = bar_v_full + 8 * STAGES;   // STAGES
  const uint32_t bar_s_full = bar_v_empty + 8 * STAGES;   // 2   S_i ready            (MMA -> softmax i)
  const uint32_t bar_p_full = bar_s_full + 16;            // 2   P_i written, O_i ok  (softmax i -> MMA)
  const uint32_t bar_o_full = bar_p_full + 16;            // 2   last PV_i done       (MMA -> softmax i)
  const uint32_t tmem_slot = bar_o_full + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const uint32_t smem_q = base + Cfg::BAR_BYTES;                  // 2 tiles
  const uint32_t smem_k = smem_q + 2 * Cfg::TILE_BYTES;           // STAGES tiles
  const uint32_t smem_v = smem_k + STAGES * Cfg::TILE_BYTES;      // STAGES tiles

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 256;
  const int T = (N + Cfg::BC - 1) / Cfg::BC;  // KV tiles

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_q_full + 8 * i, 1);
      mbar_init(bar_s_full + 8 * i, 1);
      mbar_init(bar_p_full + 8 * i, 4);
      mbar_init(bar_o_full + 8 * i, 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_k_full + 8 * s, 1);
      mbar_init(bar_k_empty + 8 * s, 1);
      mbar_init(bar_v_full + 8 * s, 1);
      mbar_init(bar_v_empty + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
  if (warp == 0) {
    // ---------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      auto load_tile = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int row0, uint64_t policy) {
        mbar_arrive_expect_tx(bar, Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < NCH; ++c) tma_load_3d(dst + c * Cfg::CHUNK_BYTES, tm, bar, c * CW, row0, bh, policy);
      };
      load_tile(&tmQ, smem_q, bar_q_full, q0, kPolicyEvictFirst);
      for (int j = 0; j < T; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(bar_k_empty + 8 * s, ph ^ 1);
        load_tile(&tmK, smem_k + s * Cfg::TILE_BYTES, bar_k_full + 8 * s, j * Cfg::BC, kPolicyEvictLast);
        if (j == 0) load_tile(&tmQ, smem_q + Cfg::TILE_BYTES, bar_q_full + 8, q0 + 128, kPolicyEvictFirst);
        mbar_wait(bar_v_empty + 8 * s, ph ^ 1);
        if constexpr (Cfg::V_DN) {
          // V^T tile: D rows x 128 keys, as two [D rows x 64 keys] 128B-swizzled boxes (keys contiguous = K-major B)
          const uint32_t dst = smem_v + s * Cfg::TILE_BYTES;
          mbar_arrive_expect_tx(bar_v_full + 8 * s, Cfg::TILE_BYTES);
          tma_load_3d(dst, &tmV, bar_v_full + 8 * s, j * Cfg::BC, 0, bh, kPolicyEvictLast);
          tma_load_3d(dst + D * 128, &tmV, bar_v_full + 8 * s, j * Cfg::BC + 64, 0, bh, kPolicyEvictLast);
        } else {
          load_tile(&tmV, smem_v + s * Cfg::TILE_BYTES, bar_v_full + 8 * s, j * Cfg::BC, kPolicyEvictLast);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, Cfg::BC, true, false, false);  // Q, K both K-major (D contiguous)
      // P from TMEM; V is MN-major ([keys, D], D contiguous) or, for V^T input, K-major ([D, keys], keys contiguous)
      constexpr uint32_t idesc_o = make_idesc_f16(128, D, true, false, !Cfg::V_DN);
      constexpr uint64_t qk_hi = make_smem_desc_hi(16, 8 * ROWB, Cfg::SWZ_MODE);
      constexpr uint64_t v_hi = Cfg::V_DN ? make_smem_desc_hi(16, 1024, kSwizzle128B)
                                          : make_smem_desc_hi(Cfg::CHUNK_BYTES, 8 * ROWB, Cfg::SWZ_MODE);
      constexpr int KSTEPS_PER_CHUNK = CW / 16;

      auto issue_s = [&](int i, int stage) {
        const uint32_t q_addr = smem_q + i * Cfg::TILE_BYTES;
        const uint32_t k_addr = smem_k + stage * Cfg::TILE_BYTES;
        const uint32_t d_tmem = tmem_base + (i ? Cfg::S_COL1 : Cfg::S_COL0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k / KSTEPS_PER_CHUNK) * Cfg::CHUNK_BYTES + (k % KSTEPS_PER_CHUNK) * 32;
          umma_ss<1>(d_tmem, smem_desc(qk_hi, q_addr + off), smem_desc(qk_hi, k_addr + off), idesc_s, k != 0);
        }
      };
      auto issue_pv = [&](int i, int stage, bool accumulate) {
        const uint32_t v_addr = smem_v + stage * Cfg::TILE_BYTES;
        const uint32_t d_tmem = tmem_base + (i ? Cfg::O_COL1 : Cfg::O_COL0);
        const uint32_t p_tmem = tmem_base + (i ? Cfg::S_COL1 : Cfg::S_COL0);
#pragma unroll
        for (int k = 0; k < Cfg::BC / 16; ++k) {
          // 16 keys = 8 packed fp16x2 columns of P; 16 rows of V = 16 * ROWB bytes (V^T: 32 bytes inside a 64-key box)
          const uint32_t v_off = Cfg::V_DN ? uint32_t((k / 4) * (D * 128) + (k % 4) * 32) : uint32_t(k * 16 * ROWB);
          umma_ts<1>(d_tmem, p_tmem + k * 8, smem_desc(v_hi, v_addr + v_off), idesc_o,
                     (accumulate || k != 0) ? 1u : 0u);
        }
      };

      mbar_wait(bar_q_full, 0);
      mbar_wait(bar_k_full, 0);
      tc_fence_after();
      issue_s(0, 0);
      umma_commit(bar_s_full);
      mbar_wait(bar_q_full + 8, 0);
      tc_fence_after();
      issue_s(1, 0);
      umma_commit(bar_s_full + 8);
      umma_commit(bar_k_empty);
      for (int j = 0; j < T; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        const int s1 = (j + 1) % STAGES;
        const uint32_t ph1 = ((j + 1) / STAGES) & 1;
        mbar_wait(bar_v_full + 8 * s, ph);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mbar_wait(bar_p_full + 8 * i, j & 1);
          tc_fence_after();
          issue_pv(i, s, j > 0);
          if (j == T - 1) umma_commit(bar_o_full + 8 * i);
          if (i == 1) umma_commit(bar_v_empty + 8 * s);
          if (j + 1 < T) {
            if (i == 0) {
              mbar_wait(bar_k_full + 8 * s1, ph1);
              tc_fence_after();
            }
            issue_s(i, s1);
            umma_commit(bar_s_full + 8 * i);
            if (i == 1) umma_commit(bar_k_empty + 8 * s1);
          }
        }
      }
    }
    __syncwarp();
  }
  } else {
    // ---------------------------------------------------------------------------------- softmax + epilogue
    const int i = (warp >= 8) ? 1 : 0;                  // which Q tile
    const uint32_t q = warp & 3;                         // TMEM lane quadrant
    const uint32_t lane_base = (q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + (i ? Cfg::S_COL1 : Cfg::S_COL0);
    const uint32_t o_tmem = tmem_base + lane_base + (i ? Cfg::O_COL1 : Cfg::O_COL0);
    float m_ref = -INFINITY;  // reference max, in log2-scaled units
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      mbar_wait(bar_s_full + 8 * i, j & 1);
      tc_fence_after();
      uint32_t sr[128];
      tmem_ld_32x32b_x32(s_tmem, sr);
      tmem_ld_32x32b_x32(s_tmem + 32, sr + 32);
      tmem_ld_32x32b_x32(s_tmem + 64, sr + 64);
      tmem_ld_32x32b_x32(s_tmem + 96, sr + 96);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      if (j == T - 1 && (N % Cfg::BC) != 0) {
        const int valid = N - j * Cfg::BC;
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int c = 4; c < 128; c += 4) {
        mx0 = fmaxf(mx0, s[c]);
        mx1 = fmaxf(mx1, s[c + 1]);
        mx2 = fmaxf(mx2, s[c + 2]);
        mx3 = fmaxf(mx3, s[c + 3]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = mx > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // warp-uniform: rescale this warp's 32 rows of O (rows that did not move use alpha = 1)
          const float m_new = need ? mx : m_ref;
          const float alpha = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l *= alpha;
#pragma unroll
          for (int c = 0; c < D / 16; ++c) {  // 16-column pieces: the whole score row is live in registers here
            uint32_t orr[16];
            tmem_ld_32x32b_x16(o_tmem + c * 16, orr);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) orr[e] = __float_as_uint(__uint_as_float(orr[e]) * alpha);
            tmem_st_32x32b_x16(o_tmem + c * 16, orr);
          }
          tmem_wait_st();
        }
      }
      // P = exp2(s * scale_log2 - m_ref), row sum in fp32, P packed to fp16 pairs in place
      float l0 = 0.f, l1 = 0.f;
      const float neg_m = -m_ref;
#pragma unroll
      for (int c = 0; c < 128; c += 2) {
        const float p0 = fast_exp2(fmaf(s[c], scale_log2, neg_m));
        const float p1 = fast_exp2(fmaf(s[c + 1], scale_log2, neg_m));
        l0 += p0;
        l1 += p1;
        sr[c >> 1] = pack_half2(p0, p1);
      }
      l += l0 + l1;
      tmem_st_32x32b_x32(s_tmem, sr);
      tmem_st_32x32b_x32(s_tmem + 32, sr + 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full + 8 * i);
    }
    // ---- epilogue: O_i / l -> fp16 -> swizzled smem (reusing this tile's Q buffer) -> TMA store
    mbar_wait(bar_o_full + 8 * i, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const uint32_t stage_base = smem_q + i * Cfg::TILE_BYTES + q * 32 * ROWB;  // this warp's 32 rows inside each chunk
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t orr[32];
      tmem_ld_32x32b_x32(o_tmem + c * 32, orr);
      tmem_wait_ld();
      const int chunk = (c * 32) / CW;
      const int sub0 = ((c * 32) % CW) / 8;  // first 16-byte piece inside the smem row
      const uint32_t row_addr = stage_base + chunk * Cfg::CHUNK_BYTES + lane * ROWB;
      const uint32_t xr = (ROWB == 128) ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float* f = reinterpret_cast<const float*>(orr + 8 * g);
        st_shared_v4(row_addr + (((sub0 + g) ^ xr) << 4), pack_half2(f[0] * inv_l, f[1] * inv_l),
                     pack_half2(f[2] * inv_l, f[3] * inv_l), pack_half2(f[4] * inv_l, f[5] * inv_l),
                     pack_half2(f[6] * inv_l, f[7] * inv_l));
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    const int row0 = q0 + i * 128 + int(q) * 32;
    if (lane == 0 && row0 < N) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) tma_store_3d(&tmO, stage_base + c * Cfg::CHUNK_BYTES, c * CW, row0, bh);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, Cfg::TMEM_COLS);
}

template <class Cfg>
static int launch_fa2(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N, float scale,
                      cudaStream_t stream, const DeviceInfo& di) {
  constexpr int D = Cfg::D;
  const uint64_t BH = uint64_t(B) * uint64_t(H);
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_tmap_3d_u16(&tmQ, Q, BH, N, D, uint64_t(N) * D, D, 1, 128, Cfg::CW, Cfg::ROWB))) return rc;
  if ((rc = make_tmap_3d_u16(&tmK, K, BH, N, D, uint64_t(N) * D, D, 1, 128, Cfg::CW, Cfg::ROWB))) return rc;
  if (Cfg::V_DN) {
    if (N % 8) return set_error(B200K_EALIGN, "b200k_fa2_fwd_f16: V as [B,H,D,N] needs N %% 8 == 0 (16-byte rows), got N=%lld", (long long)N);
    rc = make_tmap_3d_u16(&tmV, V, BH, D, N, uint64_t(N) * D, N, 1, D, 64, 128);
  } else {
    rc = make_tmap_3d_u16(&tmV, V, BH, N, D, uint64_t(N) * D, D, 1, 128, Cfg::CW, Cfg::ROWB);
  }
  if (rc) return rc;
  if ((rc = make_tmap_3d_u16(&tmO, O, BH, N, D, uint64_t(N) * D, D, 1, 32, Cfg::CW, Cfg::ROWB))) return rc;
  auto kern = fa2_fwd_tcgen05_kernel<Cfg>;
  static bool attr_set[64] = {};
  if (!attr_set[di.device]) {
    B200K_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[di.device] = true;
  }
  dim3 grid(unsigned((N + 255) / 256), unsigned(BH));
  const float scale_log2 = scale * 1.4426950408889634f;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, int(N), scale_log2);
  B200K_CHECK_CUDA(cudaGetLastError());
  return B200K_OK;
}

}  // namespace b200k

extern "C" int b200k_fa2_fwd_f16(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N,
                                 int64_t D, float scale, int v_is_dn, int variant, void* stream) {
  using namespace b200k;
  (void)variant;
  if (!Q || !K || !V || !O) return set_error(B200K_EARG, "b200k_fa2_fwd_f16: null pointer");
  if (B < 1 || H < 1 || N < 1 || N > INT32_MAX || B * H > 65535)
    return set_error(B200K_ESHAPE, "b200k_fa2_fwd_f16: need B,H,N >= 1 and B*H <= 65535 (got B=%lld H=%lld N=%lld)",
                     (long long)B, (long long)H, (long long)N);
  if (scale <= 0.f) scale = 1.0f / sqrtf(float(D));
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (v_is_dn) {
    switch (D) {
      case 32: return launch_fa2<Fa2Cfg<32, 4, true>>(Q, K, V, O, B, H, N, scale, s, di);
      case 64: return launch_fa2<Fa2Cfg<64, 4, true>>(Q, K, V, O, B, H, N, scale, s, di);
      case 96: return launch_fa2<Fa2Cfg<96, 3, true>>(Q, K, V, O, B, H, N, scale, s, di);
      case 128: return launch_fa2<Fa2Cfg<128, 2, true>>(Q, K, V, O, B, H, N, scale, s, di);
      default: break;
    }
  }
  switch (D) {
    case 32: return launch_fa2<Fa2Cfg<32, 4>>(Q, K, V, O, B, H, N, scale, s, di);
    case 64: return launch_fa2<Fa2Cfg<64, 4>>(Q, K, V, O, B, H, N, scale, s, di);
    case 96: return launch_fa2<Fa2Cfg<96, 3>>(Q, K, V, O, B, H, N, scale, s, di);
    case 128: return launch_fa2<Fa2Cfg<128, 2>>(Q, K, V, O, B, H, N, scale, s, di);
    default: return set_error(B200K_EHEADDIM, "headdim not support! (b200k_fa2_fwd_f16: D=%lld, supported 32/64/96/128)", (long long)D);
  }
}
