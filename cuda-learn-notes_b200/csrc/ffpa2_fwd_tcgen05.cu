// FFPA forward, CTA-pair variant (cta_group::2) for head dims that are multiples of 256 (256, 512, 768, 1024).
//
// Same algorithm as ffpa_fwd_tcgen05.cu ("D-sliced O", double-buffered S in TMEM, P over S, TS-form PV), but two CTAs
// of a 2x1x1 cluster own two ADJACENT 128-row Q tiles of the same (batch, head, slice) and run ONE stream of
// tcgen05.mma.cta_group::2 instructions (M = 256):
//   S  = [Q_0; Q_1] K_j^T   M=256, N=128 keys : each CTA stages only HALF of the K tile (64 keys), read by both SMs
//   O += [P_0; P_1] V_j     M=256, N=256 cols : each CTA stages only its 128 columns of the V slice
// so the shared-memory traffic and the TMA fill per CTA drop by a third (score MMA: 4 KB of Q + 2 KB of K per
// instruction instead of 4 + 4 — the 1-CTA kernel sits on the 128 B/clk smem limit there) and the same ring bytes
// hold twice as many KV tiles of look-ahead.  The leader CTA issues all MMAs; commits are multicast to both CTAs;
// both CTAs' softmax warpgroups arrive on the leader's p_full barrier (remote mbarrier arrive).
#include <cmath>

#include "abi_common.cuh"
#include "ptx.cuh"

namespace b200k {

namespace ffpa2 {
constexpr int BR = 128, BC = 128, CW = 64;
constexpr int QBOX = 128 * 128;   // [128 rows x 64 fp16]
constexpr int KBOX = 64 * 128;    // [64 keys x 64 fp16]: this CTA's half of a K chunk
constexpr int VBOX = 128 * 128;   // [128 keys x 64 cols]
constexpr int STAGE_BYTES = 32768;
constexpr int BAR_BYTES = 1024;
constexpr int MAX_STAGES = 6;
constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256;
constexpr int TMEM_COLS = 512;
constexpr int THREADS = 256;
constexpr float kRescaleThreshold = 8.0f;
}  // namespace ffpa2

// Ring stage contents (32 KB): QK stage, Q resident: 4 K half-chunks [64 x 64]; QK stage, Q streamed: {Q chunk
// [128 x 64], K half-chunk}; V stage: 2 V chunks [128 x 64] = this CTA's 128 columns of the 256-column slice.
template <bool Q_RESIDENT>
__global__ void __launch_bounds__(ffpa2::THREADS, 1)
ffpa2_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKh,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int N, int D,
                         int stages, float scale_log2) {
  using namespace ffpa2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const uint32_t bar_full = base;                          // MAX_STAGES (leader's are used)
  const uint32_t bar_empty = base + 8 * MAX_STAGES;        // MAX_STAGES (each CTA its own, multicast commits)
  const uint32_t bar_q_full = bar_empty + 8 * MAX_STAGES;  // 1 (leader's)
  const uint32_t bar_s_full = bar_q_full + 8;              // 2
  const uint32_t bar_p_full = bar_s_full + 16;             // 2 (leader's; 8 arrivals = 4 warps x 2 CTAs)
  const uint32_t bar_pv_done = bar_p_full + 16;            // 1
  const uint32_t bar_o_full = bar_pv_done + 8;             // 1
  const uint32_t tmem_slot = bar_o_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const int nqk = D / CW;
  const uint32_t smem_q = base + BAR_BYTES;
  const uint32_t smem_ring = smem_q + (Q_RESIDENT ? nqk * QBOX : 0);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);  // warp-uniform for the compiler
  const bool leader = (rank == 0);
  const int bh = blockIdx.z;
  const int q0 = blockIdx.x * BR;             // grid.x is even: CTA pair = Q tiles (2c, 2c+1)
  const int col0 = blockIdx.y * 256;          // this cluster's slice of the head dim (always 256 wide here)
  const int my_col0 = col0 + int(rank) * 128; // the 128 V columns this CTA stages
  const int T = (N + BC - 1) / BC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKh);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_q_full, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_full + 8, 1);
    mbar_init(bar_p_full, 8);
    mbar_init(bar_p_full + 8, 8);
    mbar_init(bar_pv_done, 1);
    mbar_init(bar_o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_slot, TMEM_COLS);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------------------------- TMA producer (both CTAs)
    // one lane is elected once and runs the whole role loop alone (see fa2_fwd_tcgen05.cu / tools/ubench/ubench_attn.cu)
    if (elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == stages) { stage = 0; phase ^= 1; }
    };
    if (Q_RESIDENT) {
      {
        if (leader) mbar_arrive_expect_tx(bar_q_full, 2 * nqk * QBOX);
        const uint32_t qb = mapa(bar_q_full, 0);
        for (int c = 0; c < nqk; ++c) tma_load_3d_2sm(smem_q + c * QBOX, &tmQ, qb, c * CW, q0, bh, kPolicyEvictFirst);
      }
    }
    auto load_qk = [&](int j) {
      if (Q_RESIDENT) {
        for (int c = 0; c < nqk; c += 4) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const int nb = min(4, nqk - c);
          const uint32_t dst = smem_ring + stage * STAGE_BYTES;
          {
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * nb * KBOX);
            const uint32_t fb = mapa(bar_full + 8 * stage, 0);
            for (int b = 0; b < nb; ++b)
              tma_load_3d_2sm(dst + b * KBOX, &tmKh, fb, (c + b) * CW, j * BC + int(rank) * 64, bh, kPolicyEvictLast);
          }
          advance();
        }
      } else {
        for (int c = 0; c < nqk; ++c) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t dst = smem_ring + stage * STAGE_BYTES;
          {
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * (QBOX + KBOX));
            const uint32_t fb = mapa(bar_full + 8 * stage, 0);
            tma_load_3d_2sm(dst, &tmQ, fb, c * CW, q0, bh, kPolicyEvictLast);
            tma_load_3d_2sm(dst + QBOX, &tmKh, fb, c * CW, j * BC + int(rank) * 64, bh, kPolicyEvictLast);
          }
          advance();
        }
      }
    };
    auto load_v = [&](int j) {
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      const uint32_t dst = smem_ring + stage * STAGE_BYTES;
      {
        if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * 2 * VBOX);
        const uint32_t fb = mapa(bar_full + 8 * stage, 0);
        tma_load_3d_2sm(dst, &tmV, fb, my_col0, j * BC, bh, kPolicyEvictLast);
        tma_load_3d_2sm(dst + VBOX, &tmV, fb, my_col0 + CW, j * BC, bh, kPolicyEvictLast);
      }
      advance();
    };
    load_qk(0);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) load_qk(j + 1);
      load_v(j);
    }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer (leader CTA)
    // one lane is elected once and runs the whole role loop alone (see fa2_fwd_tcgen05.cu / tools/ubench/ubench_attn.cu)
    if (elect_one() && leader) {  // this order keeps the MMA issue free of waterfall loops (SASS-checked)
      constexpr uint32_t idesc_s = make_idesc_f16(256, BC, true, false, false);   // M = 256 over the CTA pair
      constexpr uint32_t idesc_o = make_idesc_f16(256, 256, true, false, true);   // N = 256: 128 columns per CTA
      constexpr uint64_t qk_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
      constexpr uint64_t v_hi = make_smem_desc_hi(VBOX, 1024, kSwizzle128B);
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      auto issue_s = [&](int buf) {
        const uint32_t d_tmem = tmem_base + (buf ? S_COL1 : S_COL0);
        if (Q_RESIDENT) {
          for (int c = 0; c < nqk; c += 4) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t kb = smem_ring + stage * STAGE_BYTES;
            const int nb = min(4, nqk - c);
            {
              for (int b = 0; b < nb; ++b) {
                const uint32_t qa = smem_q + (c + b) * QBOX;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ss<2>(d_tmem, smem_desc(qk_hi, qa + k * 32), smem_desc(qk_hi, kb + b * KBOX + k * 32), idesc_s,
                             (c + b + k) != 0 ? 1u : 0u);
              }
              umma_commit_2sm(bar_empty + 8 * stage, 0b11);
              if (c + 4 >= nqk) umma_commit_2sm(bar_s_full + 8 * buf, 0b11);
            }
            advance();
          }
        } else {
          for (int c = 0; c < nqk; ++c) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t qa = smem_ring + stage * STAGE_BYTES;
            const uint32_t kb = qa + QBOX;
            {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss<2>(d_tmem, smem_desc(qk_hi, qa + k * 32), smem_desc(qk_hi, kb + k * 32), idesc_s,
                           (c + k) != 0 ? 1u : 0u);
              umma_commit_2sm(bar_empty + 8 * stage, 0b11);
              if (c + 1 >= nqk) umma_commit_2sm(bar_s_full + 8 * buf, 0b11);
            }
            advance();
          }
        }
      };
      auto issue_pv = [&](int buf, bool accumulate, bool last_tile) {
        const uint32_t p_tmem = tmem_base + (buf ? S_COL1 : S_COL0);
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t va = smem_ring + stage * STAGE_BYTES;
        {
#pragma unroll
          for (int k = 0; k < BC / 16; ++k)
            umma_ts<2>(tmem_base + O_COL, p_tmem + k * 8, smem_desc(v_hi, va + k * 16 * 128), idesc_o,
                       (accumulate || k != 0) ? 1u : 0u);
          umma_commit_2sm(bar_empty + 8 * stage, 0b11);
          umma_commit_2sm(bar_pv_done, 0b11);
          if (last_tile) umma_commit_2sm(bar_o_full, 0b11);
        }
        advance();
      };
      if (Q_RESIDENT) {
        mbar_wait(bar_q_full, 0);
        tc_fence_after();
      }
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s((j + 1) & 1);
        mbar_wait(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        issue_pv(j & 1, j > 0, j == T - 1);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------------- softmax + epilogue
    const uint32_t q = warp & 3;
    const uint32_t lane_base = (q * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_base + O_COL;
    float m_ref = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      const int buf = j & 1;
      const uint32_t s_tmem = tmem_base + lane_base + (buf ? S_COL1 : S_COL0);
      mbar_wait(bar_s_full + 8 * buf, (j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[128];
      tmem_ld_32x32b_x32(s_tmem, sr);
      tmem_ld_32x32b_x32(s_tmem + 32, sr + 32);
      tmem_ld_32x32b_x32(s_tmem + 64, sr + 64);
      tmem_ld_32x32b_x32(s_tmem + 96, sr + 96);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      if (j == T - 1 && (N % BC) != 0) {
        asm volatile("" ::: "memory");  // keep this a real branch (see fa2_fwd_tcgen05.cu)
        const int valid = N - j * BC;
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      const float mx = row_max<128>(s) * scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = mx > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once PV_{j-1} has completed (pv_done has completed exactly j-1 or j phases here)
          mbar_wait(bar_pv_done, (j - 1) & 1);
          tc_fence_after();
          const float m_new = need ? mx : m_ref;
          const float alpha = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l *= alpha;
          for (int c = 0; c < 256 / 16; ++c) {
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
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      const float neg_m = -m_ref;
      const float2 scale2 = make_float2(scale_log2, scale_log2), negm2 = make_float2(neg_m, neg_m);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) {
        float2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = ffma2(make_float2(s[c0 + 2 * e], s[c0 + 2 * e + 1]), scale2, negm2);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[e].x = fast_exp2(x[e].x);
          x[e].y = fast_exp2(x[e].y);
        }
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          acc0 = fadd2(acc0, x[e]);
          acc1 = fadd2(acc1, x[e + 1]);
          sr[(c0 >> 1) + e] = pack_half2(x[e].x, x[e].y);
          sr[(c0 >> 1) + e + 1] = pack_half2(x[e + 1].x, x[e + 1].y);
        }
      }
      l += (acc0.x + acc0.y) + (acc1.x + acc1.y);
      tmem_st_32x32b_x32(s_tmem, sr);
      tmem_st_32x32b_x32(s_tmem + 32, sr + 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa(bar_p_full + 8 * buf, 0));   // P is in TMEM (wait::st above): nothing else to publish
    }
    // ---- epilogue: this CTA's 128 rows x 256 columns
    mbar_wait(bar_o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const uint32_t stage_base = smem_ring + q * 32 * 128;  // ring memory is idle now: [chunk][128 rows][128 B]
    for (int c = 0; c < 256 / 32; ++c) {
      uint32_t orr[32];
      tmem_ld_32x32b_x32(o_tmem + c * 32, orr);
      tmem_wait_ld();
      const int chunk = c >> 1;
      const int sub0 = (c & 1) * 4;
      const uint32_t row_addr = stage_base + chunk * QBOX + lane * 128;
      const uint32_t xr = lane & 7;
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
    const int row0 = q0 + int(q) * 32;
    if (lane == 0 && row0 < N) {
      for (int c = 0; c < 4; ++c) tma_store_3d(&tmO, stage_base + c * QBOX, col0 + c * CW, row0, bh);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tmem_base, ffpa2::TMEM_COLS);
}

// Host launcher, called from b200k_ffpa_fwd_f16 (ffpa_fwd_tcgen05.cu) when D % 256 == 0.
int launch_ffpa_2cta(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N, int64_t D,
                     float scale, cudaStream_t s) {
  const uint64_t BH = uint64_t(B) * uint64_t(H);
  CUtensorMap tmQ, tmKh, tmV, tmO;
  int rc;
  if ((rc = make_tmap_3d_u16(&tmQ, Q, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmKh, K, BH, N, D, uint64_t(N) * D, D, 1, 64, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmV, V, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmO, O, BH, N, D, uint64_t(N) * D, D, 1, 32, 64, 128))) return rc;
  int device = 0;
  B200K_CHECK_CUDA(cudaGetDevice(&device));
  struct { int device; } di = {device};
  const bool q_resident = (D <= 512);
  const int q_bytes = q_resident ? int(D / 64) * ffpa2::QBOX : 0;
  int stages = (232448 - 1024 - ffpa2::BAR_BYTES - q_bytes) / ffpa2::STAGE_BYTES;
  if (stages > ffpa2::MAX_STAGES) stages = ffpa2::MAX_STAGES;
  const int smem = 1024 + ffpa2::BAR_BYTES + q_bytes + stages * ffpa2::STAGE_BYTES;
  const unsigned qtiles = unsigned((N + 127) / 128);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((qtiles + 1) / 2 * 2, unsigned(D / 256), unsigned(BH));
  cfg.blockDim = dim3(ffpa2::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float scale_log2 = scale * 1.4426950408889634f;
  if (q_resident) {
    auto kern = ffpa2_fwd_tcgen05_kernel<true>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, smem)) return rc;
    B200K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQ, tmKh, tmV, tmO, int(N), int(D), stages, scale_log2));
  } else {
    auto kern = ffpa2_fwd_tcgen05_kernel<false>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, smem)) return rc;
    B200K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQ, tmKh, tmV, tmO, int(N), int(D), stages, scale_log2));
  }
  return B200K_OK;
}

}  // namespace b200k
