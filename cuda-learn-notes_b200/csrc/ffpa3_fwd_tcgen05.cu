// FFPA forward without recomputing S, for head dims 256 and 512: O^T accumulation.
//
// The D-sliced kernels (ffpa_fwd / ffpa2_fwd) give every 256-column slice of O its own CTA (pair) and recompute
// S = Q K^T per slice, because a 128-row fp32 O of D columns needs D of the 512 TMEM columns.  Here the accumulator is
// TRANSPOSED: O^T[D, rows] += V^T[D, keys] * P^T[keys, rows] puts the head dim on the TMEM LANES (M = 256 over a CTA
// pair, D / 256 accumulators) and the 128 query rows of the pair's Q tile on the columns: 128 columns per accumulator,
// 256 for D = 512, which leaves room for a double-buffered S.  One pass over S per KV tile, whatever D.
//   cluster of 2 CTAs = one Q tile of 128 rows (64 per CTA), one (batch, head)
//   S  = Q K_j^T      tcgen05.mma.cta_group::2, M = 128 (64 rows per CTA), N = 256 keys (128 staged per CTA), K = D.
//                     An M = 128 pair accumulator is FOLDED (tools/ubench/probe_pair_m128.cu): lanes 0-63 of a CTA hold
//                     its 64 rows for keys [0,N/2), lanes 64-127 the same rows for keys [N/2,N): 128 TMEM columns per S.
//   softmax           128 threads per CTA, thread L owns row L % 64, keys 128 (L / 64) ... +128: two threads per row, partial
//                     maxima swapped through shared memory; P (fp16) is written to SHARED memory, K-major, swizzled
//   O^T += V_j^T P^T  cta_group::2, M = 256 (each CTA supplies 128 head-dim columns of V as an MN-major A operand straight
//                     from the [keys, D] tile), N = 128 rows (each CTA supplies its 64 rows of P as the B operand), K = 256 keys
// O^T's columns are the rows of BOTH CTAs, so the lazy-rescale factor of a row must reach the peer before the next PV -
// without a cluster-scope memory fence on the per-tile path (mbarrier.arrive.release.cluster = MEMBAR.ALL.GPU, 1-1.5 k
// cycles per arrive, measured).  Every tile each row's factor (1.0 when the row max did not move) goes to both CTAs by
// st.async, completing on the receiving CTA's decision barrier of that tile; each first-half-row warp sends its "moved"
// flag to the leader the same way on p_full, with its arrive.  The MMA thread reads the flags after p_full(j): none set
// (the steady state) -> PV(j) is issued at once; any set -> the verdict "rescale" goes out (st.async again, so a decision
// barrier completes only when the verdict AND all 128 factors have landed) and it waits until the softmax warps of both
// CTAs have scaled their lanes of O^T.  The softmax warps look at decision(j-1) at the END of tile j, so no cross-CTA
// wait sits on the per-tile softmax chain.
// The epilogue divides column c by the row sum l[c] (exchanged the same way) and stores O transposed.
#include <cmath>

#include "abi_common.cuh"
#include "ptx.cuh"

namespace b200k {

namespace ffpa3 {
constexpr int BC = 256, CW = 64;   // 256 keys per KV tile: the fixed latencies of the softmax chain are paid once per 256 keys
constexpr int QBOX = 64 * 128;     // [64 rows x 64 fp16]: one D chunk of this CTA's 64 Q rows
constexpr int KBOX = 128 * 128;    // [128 keys x 64 fp16]: this CTA's half of a K chunk
constexpr int VBOX = 128 * 128;    // [128 keys x 64 head-dim columns]
constexpr int PBOX = 64 * 128;     // [64 rows x 64 keys]
constexpr int STAGE_BYTES = 32768; // 2 K half-chunks, or the 2 V boxes of one accumulator for 128 keys
// Every byte of shared memory counts: the operand ring must keep ~64 B/clk of K and V in flight against an L2 latency of
// ~1.5 us, and with Q (64 KB) and P (32 KB) resident a fourth 32 KB stage only fits if the bookkeeping stays inside 3 KB and
// the dynamic shared memory needs no alignment slack (it starts 1 KB into the window; the kernel checks).
constexpr int BAR_BYTES = 256;
constexpr int MISC_BYTES = 2816;   // xchg[2][128] f32 (2nd half doubles as 1/l at the end) | alpha[3][128] f32 | flags[4][4] | decision[4]
constexpr int MAX_STAGES = 6;
constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256;
constexpr int TMEM_COLS = 512;
constexpr int THREADS = 256;
constexpr float kRescaleThreshold = 8.0f;
}  // namespace ffpa3

__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// Arrives that publish only (a) this CTA's own shared memory to its own tensor core and (b) words that travel with
// st.async below: release at CTA scope is all they need.  The .release.cluster form costs a GPU-wide memory fence
// (MEMBAR.ALL.GPU + ERRBAR, ~1-1.5 k cycles measured) and sat on the softmax -> PV chain of every tile.
__device__ __forceinline__ void mbar_arrive_cluster_release_cta(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster_release_cta(uint32_t cluster_bar, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster_relaxed(uint32_t cluster_bar, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(tx) : "memory");
}
// One 32-bit word into a CTA of the cluster, completing 4 bytes on an mbarrier of THAT CTA: whoever sees the barrier phase
// complete sees the word - no fence on either side.
__device__ __forceinline__ void st_async_u32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v), "r"(cluster_bar)
               : "memory");
}
// wait with cluster-scope acquire: data written by the peer CTA before its release.cluster arrive is visible afterwards
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > B200K_SPIN_LIMIT_CYCLES) mbar_timeout_trap(bar, parity);
  }
}

template <int NACC>   // D = 256 * NACC
__global__ void __launch_bounds__(ffpa3::THREADS, 1)
ffpa3_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmKh,
                         const __grid_constant__ CUtensorMap tmV, __half* __restrict__ O, int N, int stages, float scale_log2) {
  using namespace ffpa3;
  constexpr int D = 256 * NACC;
  constexpr int nqk = D / CW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();   // the swizzled operand tiles need 1024-byte alignment and there is no slack to realign
  uint8_t* base_ptr = smem_raw;

  const uint32_t bar_full = base;                          // MAX_STAGES (leader's are used)
  const uint32_t bar_empty = base + 8 * MAX_STAGES;        // MAX_STAGES (each CTA its own, multicast commits)
  const uint32_t bar_q_full = bar_empty + 8 * MAX_STAGES;  // 1 (leader's)
  const uint32_t bar_s_full = bar_q_full + 8;              // 2 (each CTA, multicast commit)
  const uint32_t bar_s_free = bar_s_full + 16;             // 2 (leader's; 8 arrivals = 4 warps x 2 CTAs)
  const uint32_t bar_p_full = bar_s_free + 16;             // 2 (leader's; 8 arrivals + 4 flag words by st.async)
  const uint32_t bar_pv_done = bar_p_full + 16;            // 2 (each CTA, multicast commit): P buffer b free / O^T stable
  const uint32_t bar_decision = bar_pv_done + 16;          // 2 (each CTA; 1 arrival + 516 bytes): verdict(j) and the 128 factors are readable
  const uint32_t bar_rsdone = bar_decision + 16;           // 1 (leader's; 8 arrivals): a requested rescale has been applied
  const uint32_t bar_o_full = bar_rsdone + 8;              // 1 (each CTA, multicast commit)
  const uint32_t tmem_slot = bar_o_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const uint32_t misc = base + BAR_BYTES;
  const uint32_t xchg = misc;                 // [2][128] f32: partial row maxima / row sums of the two half-row threads
  const uint32_t alpha_buf = misc + 1024;     // [3][128] f32: rescale factor of every row of the pair, tile j in slot j % 3
  const uint32_t linv_buf = misc + 512;       // [128] f32: 1 / row sum (the second xchg slot, dead by then)
  const uint32_t flag_buf = misc + 2560;      // [4][4] u32: "a row of warp w of CTA c moved its max" (c*2 + w)
  const uint32_t decision_buf = misc + 2624;  // [4] u32: the MMA thread's verdict for tile j (slot j & 3)
  const uint32_t smem_q = misc + MISC_BYTES;
  const uint32_t smem_p = smem_q + nqk * QBOX;            // ONE buffer of 4 chunks x PBOX (64 rows x 256 keys)
  const uint32_t smem_ring = smem_p + 4 * PBOX;

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const uint32_t peer = rank ^ 1u;
  const bool leader = (rank == 0);
  const int bh = blockIdx.y;
  const int q0 = int(blockIdx.x >> 1) * 128;  // the pair's Q tile; this CTA stages rows q0 + 64 rank ... + 64
  const int T = (N + BC - 1) / BC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQh);
    tma_prefetch_desc(&tmKh);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_q_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_s_full + 8 * b, 1);
      mbar_init(bar_s_free + 8 * b, 8);
      mbar_init(bar_p_full + 8 * b, 8);
      mbar_init(bar_pv_done + 8 * b, 1);
      mbar_init(bar_decision + 8 * b, 1);
    }
    mbar_init(bar_rsdone, 8);
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
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      {
        if (leader) mbar_arrive_expect_tx(bar_q_full, 2 * nqk * QBOX);
        const uint32_t qb = mapa(bar_q_full, 0);
        for (int c = 0; c < nqk; ++c)
          tma_load_3d_2sm(smem_q + c * QBOX, &tmQh, qb, c * CW, q0 + int(rank) * 64, bh, kPolicyEvictFirst);
      }
      auto load_k = [&](int j) {
        for (int c = 0; c < nqk; c += 2) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t dst = smem_ring + stage * STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * 2 * KBOX);
          const uint32_t fb = mapa(bar_full + 8 * stage, 0);
          for (int b = 0; b < 2; ++b)
            tma_load_3d_2sm(dst + b * KBOX, &tmKh, fb, (c + b) * CW, j * BC + int(rank) * 128, bh, kPolicyEvictLast);
          advance();
        }
      };
      auto load_v = [&](int j) {
        for (int a = 0; a < NACC; ++a) {
          for (int kh = 0; kh < 2; ++kh) {   // keys [128 kh, 128 kh + 128) of the tile
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t dst = smem_ring + stage * STAGE_BYTES;
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * 2 * VBOX);
            const uint32_t fb = mapa(bar_full + 8 * stage, 0);
            const int col = a * 256 + int(rank) * 128;   // this CTA's 128 head-dim columns of accumulator a
            tma_load_3d_2sm(dst, &tmV, fb, col, j * BC + kh * 128, bh, kPolicyEvictLast);
            tma_load_3d_2sm(dst + VBOX, &tmV, fb, col + CW, j * BC + kh * 128, bh, kPolicyEvictLast);
            advance();
          }
        }
      };
      load_k(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer (leader CTA)
    if (elect_one() && leader) {
      constexpr uint32_t idesc_s = make_idesc(128, BC, 0, false, false);       // M = 128 over the pair: 64 rows per CTA
      constexpr uint32_t idesc_o = make_idesc(256, 128, 0, /*a_mn=*/true, false);  // O^T: M = head dim, N = 128 rows
      constexpr uint64_t k_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);     // Q, K, P: K-major, 128-byte rows
      constexpr uint64_t v_hi = make_smem_desc_hi(VBOX, 1024, kSwizzle128B);   // V as A: MN-major, next 64 columns one box on
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      auto issue_s = [&](int buf) {
        const uint32_t d_tmem = tmem_base + (buf ? S_COL1 : S_COL0);
        for (int c = 0; c < nqk; c += 2) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t kb = smem_ring + stage * STAGE_BYTES;
          for (int b = 0; b < 2; ++b) {
            const uint32_t qa = smem_q + (c + b) * QBOX;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<2>(d_tmem, smem_desc(k_hi, qa + k * 32), smem_desc(k_hi, kb + b * KBOX + k * 32), idesc_s,
                         (c + b + k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(bar_empty + 8 * stage, 0b11);
          if (c + 2 >= nqk) umma_commit_2sm(bar_s_full + 8 * buf, 0b11);
          advance();
        }
      };
      auto issue_pv = [&](bool accumulate, bool last_tile) {
        for (int a = 0; a < NACC; ++a) {
          for (int kh = 0; kh < 2; ++kh) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t va = smem_ring + stage * STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 8; ++k)   // 16 keys per MMA: V rows 16 k of this stage, P chunk 2 kh + k / 4, 32 bytes per step
              umma_ss<2>(tmem_base + O_COL + a * 128, smem_desc(v_hi, va + k * 16 * 128),
                         smem_desc(k_hi, smem_p + (2 * kh + (k >> 2)) * PBOX + (k & 3) * 32), idesc_o,
                         (accumulate || kh != 0 || k != 0) ? 1u : 0u);
            umma_commit_2sm(bar_empty + 8 * stage, 0b11);
            if (a == NACC - 1 && kh == 1) {
              umma_commit_2sm(bar_pv_done, 0b11);
              if (last_tile) umma_commit_2sm(bar_o_full, 0b11);
            }
            advance();
          }
        }
      };
      mbar_wait(bar_q_full, 0);
      tc_fence_after();
      issue_s(0);
      uint32_t rs_phase = 0;
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          const int b = (j + 1) & 1;
          if (j + 1 >= 2) {  // the softmax warps of both CTAs hold S(j-1) in registers
            mbar_wait(bar_s_free + 8 * b, ((j - 1) >> 1) & 1);
            tc_fence_after();
          }
          issue_s(b);
        }
        mbar_wait_cluster(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
        {
          // did any row of the pair move its reference max in tile j?  (flags written by all 8 warps before their arrive)
          const uint32_t fb = flag_buf + (j & 3) * 16;
          const uint32_t f = ld_shared_u32(fb) | ld_shared_u32(fb + 4) | ld_shared_u32(fb + 8) | ld_shared_u32(fb + 12);
          // the verdict and the 128 rescale factors of tile j (st.async by the softmax threads) complete on the same barrier
          const uint32_t d_addr = decision_buf + (j & 3) * 4;
          for (uint32_t c = 0; c < 2; ++c) {
            const uint32_t db = mapa(bar_decision + 8 * (j & 1), c);
            mbar_arrive_expect_tx_cluster_relaxed(db, 4 + 128 * 4);
            st_async_u32(mapa(d_addr, c), f, db);
          }
          if (f != 0) {
            mbar_wait_cluster(bar_rsdone, rs_phase);
            rs_phase ^= 1;
          }
        }
        tc_fence_after();
        issue_pv(j > 0, j == T - 1);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------------- softmax + epilogue
    const uint32_t q = warp & 3;
    const uint32_t L = q * 32 + lane;            // TMEM lane of this thread
    const uint32_t r = L & 63;                   // row inside this CTA's 64
    const uint32_t h = L >> 6;                   // which 64 keys of the tile
    const uint32_t grow = rank * 64 + r;         // row inside the pair's 128 = column of O^T
    const uint32_t lane_base = (q * 32) << 16;
    float m_ref = -INFINITY;
    float l = 0.f;
    auto apply_decision = [&](int t) {   // t = tile whose factors are in slot t % 3
      mbar_wait_cluster(bar_decision + 8 * (t & 1), (t >> 1) & 1);
      if (ld_shared_u32(decision_buf + (t & 3) * 4) != 0) {
        if (t >= 1) {  // O^T must be stable: PV(t-1) complete (PV(t) is held back by the MMA thread until this is done)
          mbar_wait(bar_pv_done, (t - 1) & 1);
          tc_fence_after();
        }
        for (int a = 0; a < NACC; ++a) {
          const uint32_t o_tmem = tmem_base + lane_base + O_COL + a * 128;
          for (int c = 0; c < 128 / 16; ++c) {
            uint32_t orr[16];
            tmem_ld_32x32b_x16(o_tmem + c * 16, orr);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e)
              orr[e] = __float_as_uint(__uint_as_float(orr[e]) * ld_shared_f32(alpha_buf + ((t % 3) * 128 + c * 16 + e) * 4));
            tmem_st_32x32b_x16(o_tmem + c * 16, orr);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(bar_rsdone, 0));
      }
    };
    for (int j = 0; j < T; ++j) {
      const int buf = j & 1;
      const uint32_t jp = (j >> 1) & 1;
      const uint32_t s_tmem = tmem_base + lane_base + (buf ? S_COL1 : S_COL0);
      mbar_wait(bar_s_full + 8 * buf, jp);
      tc_fence_after();
      uint32_t sr[128];
      tmem_ld_32x32b_x32(s_tmem, sr);
      tmem_ld_32x32b_x32(s_tmem + 32, sr + 32);
      tmem_ld_32x32b_x32(s_tmem + 64, sr + 64);
      tmem_ld_32x32b_x32(s_tmem + 96, sr + 96);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa(bar_s_free + 8 * buf, 0));
      float* s = reinterpret_cast<float*>(sr);
      if (j == T - 1 && (N % BC) != 0) {
        asm volatile("" ::: "memory");
        const int valid = N - j * BC - int(h) * 128;
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      // row max: this thread's 64 keys, then the partner thread's (the other half of the row)
      const float mx_half = row_max<128>(s) * scale_log2;
      st_shared_f32(xchg + (buf * 128 + L) * 4, mx_half);
      named_bar_sync(1, 128);
      const float mx = fmaxf(mx_half, ld_shared_f32(xchg + (buf * 128 + (L ^ 64u)) * 4));
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_ref = mx;
      } else {
        need = mx > m_ref + kRescaleThreshold;
        if (need) {
          alpha = fast_exp2(m_ref - mx);
          m_ref = mx;
          l *= alpha;
        }
      }
      // publish this row's factor (one of the two threads of the row) to both CTAs: st.async completing on the decision
      // barrier of tile j, which every reader of the factors waits on.  Nobody waits here.
      if (h == 0) {
        const uint32_t a_addr = alpha_buf + ((j % 3) * 128 + grow) * 4;
        for (uint32_t c = 0; c < 2; ++c) st_async_u32(mapa(a_addr, c), __float_as_uint(alpha), mapa(bar_decision + 8 * buf, c));
      }
      const uint32_t any = __any_sync(0xffffffffu, need) ? 1u : 0u;
      // P = exp2(s * scale - m_ref) for this thread's 128 keys
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
      // The verdict on tile j-1 (long since there in the steady state): if a row moved, scale that column of O^T now - the
      // MMA thread holds PV(j-1) back until this is done, and the P buffer below is only free once PV(j-1) has run.
      if (j >= 1) apply_decision(j - 1);
      // the (single) P buffer was last read by PV(j-1); in the steady state that finished long ago
      if (j >= 1) mbar_wait(bar_pv_done, (j - 1) & 1);
      {
        // this thread's keys 128 h ... + 128 = chunks 2 h and 2 h + 1; row r of a chunk: 8 x 16-byte units, unit u at u ^ (r & 7)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const uint32_t row_addr = smem_p + (2 * h + cc) * PBOX + r * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            st_shared_v4(row_addr + ((u ^ (r & 7)) << 4), sr[32 * cc + 4 * u], sr[32 * cc + 4 * u + 1], sr[32 * cc + 4 * u + 2],
                         sr[32 * cc + 4 * u + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t pf = mapa(bar_p_full + 8 * buf, 0);
        if (q < 2) {   // the two threads of a row agree on "moved": the warps of the first half-row carry the flag
          mbar_arrive_expect_tx_cluster_release_cta(pf, 4);
          st_async_u32(mapa(flag_buf + ((j & 3) * 4 + rank * 2 + q) * 4, 0), any, pf);
        } else {
          mbar_arrive_cluster_release_cta(pf);
        }
      }
    }
    apply_decision(T - 1);
    // ---- epilogue: row sums of the two half-row threads -> 1 / l for every row of the pair, in both CTAs
    named_bar_sync(1, 128);                 // every local thread is done with the exchange slots of the last tile
    st_shared_f32(xchg + L * 4, l);
    named_bar_sync(1, 128);
    mbar_wait(bar_o_full, 0);               // the last PV has run: both CTAs are past every read of the second exchange slot
    tc_fence_after();
    if (h == 0) {
      const float inv = 1.0f / (l + ld_shared_f32(xchg + (L ^ 64u) * 4));
      const uint32_t a_addr = linv_buf + grow * 4;
      st_shared_f32(a_addr, inv);
      st_shared_cluster_u32(mapa(a_addr, peer), __float_as_uint(inv));
    }
    // cluster barrier #1 (all threads of both CTAs take part, the other warps below): every 1 / l written by the peer is visible
    cluster_sync();
    // O[q0 + c, a*256 + rank*128 + L] = O^T[L][c] / l[c]
    __half* obase = O + (size_t(bh) * size_t(N) + size_t(q0)) * size_t(D);
    for (int a = 0; a < NACC; ++a) {
      const uint32_t o_tmem = tmem_base + lane_base + O_COL + a * 128;
      const int dcol = a * 256 + int(rank) * 128 + int(L);
      for (int c = 0; c < 128 / 32; ++c) {
        uint32_t orr[32];
        tmem_ld_32x32b_x32(o_tmem + c * 32, orr);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int row = c * 32 + e;
          if (q0 + row < N)
            obase[size_t(row) * D + dcol] = __float2half_rn(__uint_as_float(orr[e]) * ld_shared_f32(linv_buf + row * 4));
        }
      }
    }
  }
  if (warp < 4) cluster_sync();   // cluster barrier #1 for the warps that do not run the epilogue

  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tmem_base, ffpa3::TMEM_COLS);
}

// Host launcher, called from b200k_ffpa_fwd_f16 (ffpa_fwd_tcgen05.cu): the default for D = 512, variant bit 0x200 for D = 256.
int launch_ffpa_otrans(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N, int64_t D,
                       float scale, cudaStream_t s) {
  const uint64_t BH = uint64_t(B) * uint64_t(H);
  CUtensorMap tmQh, tmKh, tmV;
  int rc;
  if ((rc = make_tmap_3d_u16(&tmQh, Q, BH, N, D, uint64_t(N) * D, D, 1, 64, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmKh, K, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmV, V, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  int device = 0;
  B200K_CHECK_CUDA(cudaGetDevice(&device));
  const int nacc = int(D / 256);
  const int fixed = ffpa3::BAR_BYTES + ffpa3::MISC_BYTES + int(D / 64) * ffpa3::QBOX + 4 * ffpa3::PBOX;
  int stages = (232448 - fixed) / ffpa3::STAGE_BYTES;
  if (stages > ffpa3::MAX_STAGES) stages = ffpa3::MAX_STAGES;
  const int smem = fixed + stages * ffpa3::STAGE_BYTES;
  const unsigned qtiles = unsigned((N + 127) / 128);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(qtiles * 2, unsigned(BH), 1);
  cfg.blockDim = dim3(ffpa3::THREADS);
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
  __half* Op = static_cast<__half*>(O);
  if (nacc == 2) {
    auto kern = ffpa3_fwd_tcgen05_kernel<2>;
    if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), device, smem)) return rc2;
    B200K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQh, tmKh, tmV, Op, int(N), stages, scale_log2));
  } else {
    auto kern = ffpa3_fwd_tcgen05_kernel<1>;
    if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), device, smem)) return rc2;
    B200K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQh, tmKh, tmV, Op, int(N), stages, scale_log2));
  }
  return B200K_OK;
}

}  // namespace b200k
