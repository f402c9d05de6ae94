// FFPA forward (large head dim, D = 256 .. 1024 step 64) for B200 (sm_100a): O = softmax(Q K^T / sqrt(D)) V.
//
// The reference's FFPA keeps SRAM O(1) in D by tiling QK^T and PV at MMA granularity
// (ffpa-attn-mma/csrc/cuffpa/ffpa_attn_templates_L1.cuh:L7-593, numerics include/cuffpa/prefill.cuh:L273-533).
// On Blackwell the binding resource is tensor memory: a 128-row fp32 O accumulator needs D columns and TMEM has 512.
// Design here ("D-sliced O"):
//   * one CTA = one 128-row Q tile x one slice of the head dim (<= 256 output columns); grid = (N/128, slices, B*H);
//   * every CTA computes the full scores S = Q K^T over all of D (streamed in 64-wide chunks, O(1) smem in D) into a
//     DOUBLE-BUFFERED S in TMEM, so QK^T of tile j+1 runs on the tensor core while the softmax warps work on tile j;
//   * P (fp16) goes back into TMEM over S and feeds the TS-form MMA  O_slice += P V[:, slice];
//   * TMEM: S0 | S1 | O slice = 128 + 128 + 256 columns.
// The price is recomputing S once per slice (1.5x tensor work at D=512); in exchange nothing crosses CTAs and the
// accumulators are fp32 (the reference keeps the running O in fp16 for D > 64, launch_templates.cuh:L72-80).
// Q stays resident in shared memory when it fits (D <= 512: 128 KB); above that Q chunks are re-streamed with K.
//
// Warp roles (256 threads): 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..7 softmax/correction/epilogue.
#include <cmath>

#include "abi_common.cuh"
#include "ptx.cuh"

namespace b200k {

namespace ffpa {
constexpr int BR = 128, BC = 128, CW = 64;
constexpr int BOX_BYTES = 128 * 128;      // [128 rows x 64 fp16], 128B-swizzled
constexpr int STAGE_BYTES = 2 * BOX_BYTES;  // a ring stage holds two boxes
constexpr int BAR_BYTES = 3072;  // 1 KB of mbarriers + 2 KB exchange buffer of the split-row softmax
constexpr int MAX_STAGES = 6;
constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256;
constexpr int TMEM_COLS = 512;
constexpr float kRescaleThreshold = 8.0f;
}  // namespace ffpa

// Q_RESIDENT: the whole [128 x D] Q tile is loaded once; ring stages then carry {K chunk c, K chunk c+1} for QK^T.
// Otherwise ring stages carry {Q chunk c, K chunk c}.  PV stages carry {V chunk c, V chunk c+1} of this CTA's slice.
// SPLIT = 2: two softmax warpgroups share every query row (columns 0..63 / 64..127 of the score tile): warps 4..7 and
// 8..11 with the same TMEM lane quadrant work on the same 32 rows.  Partial row maxima are exchanged through shared
// memory (one named barrier of 64 threads per quadrant and tile); row sums stay partial until the epilogue.  This halves
// the softmax time per tile, which is what bounds the kernel when D is small (D = 128: QK^T + PV need ~1150 cycles per
// tile, one warpgroup of exponentials ~2250).
template <bool Q_RESIDENT, int SPLIT>
__global__ void __launch_bounds__(128 + 128 * SPLIT, 1)
ffpa_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int N, int D,
                        int slice_cols, int stages, float scale_log2, int serial) {
  const int dbg_skip = serial >> 1;  // debugging: 1 = skip the exponentials, 2 = skip all softmax work (timing probes)
  serial &= 1;
  using namespace ffpa;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const uint32_t bar_full = base;                        // MAX_STAGES
  const uint32_t bar_empty = base + 8 * MAX_STAGES;      // MAX_STAGES
  const uint32_t bar_q_full = bar_empty + 8 * MAX_STAGES;  // 1 (resident Q)
  const uint32_t bar_s_full = bar_q_full + 8;            // 2
  const uint32_t bar_p_full = bar_s_full + 16;           // 2   one per S/P buffer: the softmax warps may run one tile
                                                         //     ahead of the MMA thread, a single barrier could alias phases
  const uint32_t bar_pv_done = bar_p_full + 16;          // 1   completes once per PV_j (guards the lazy O rescale)
  const uint32_t bar_o_full = bar_pv_done + 8;           // 1   completes once, after the last PV (epilogue)
  const uint32_t tmem_slot = bar_o_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  float* xch = reinterpret_cast<float*>(base_ptr + 1024);  // [2 tile parities][2 halves][128 rows] fp32 (SPLIT = 2)
  const int nqk = (D + CW - 1) / CW;                      // 64-wide chunks of the head dim; for D % 64 == 32 the last box is half
                                                          // outside the tensor: TMA zero-fills it (adds 0 to S, 0-columns to O)
  const uint32_t smem_q = base + BAR_BYTES;               // resident Q: nqk boxes
  const uint32_t smem_ring = smem_q + (Q_RESIDENT ? nqk * BOX_BYTES : 0);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const uint32_t lane = threadIdx.x & 31;
  const int bh = blockIdx.z;
  const int q0 = blockIdx.x * BR;
  const int col0 = blockIdx.y * 256;                      // first output column of this slice
  const int ncols = min(slice_cols, D - col0);            // 32 .. 256, multiple of 32
  const int nv = (ncols + CW - 1) / CW;                   // V chunks in the slice (the TMA store clips a half-outside chunk)
  const int T = (N + BC - 1) / BC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_q_full, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_full + 8, 1);
    mbar_init(bar_p_full, 4 * SPLIT);
    mbar_init(bar_p_full + 8, 4 * SPLIT);
    mbar_init(bar_pv_done, 1);
    mbar_init(bar_o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_slot, TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------------------------- TMA producer
    // one lane is elected once and runs the whole role loop alone (see fa2_fwd_tcgen05.cu / tools/ubench/ubench_attn.cu)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      if (Q_RESIDENT) {
        {
          mbar_arrive_expect_tx(bar_q_full, nqk * BOX_BYTES);
          for (int c = 0; c < nqk; ++c)
            tma_load_3d(smem_q + c * BOX_BYTES, &tmQ, bar_q_full, c * CW, q0, bh, kPolicyEvictFirst);
        }
      }
      auto load_qk = [&](int j) {
        if (Q_RESIDENT) {
          for (int c = 0; c < nqk; c += 2) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const int nb = min(2, nqk - c);
            const uint32_t dst = smem_ring + stage * STAGE_BYTES;
            {
              mbar_arrive_expect_tx(bar_full + 8 * stage, nb * BOX_BYTES);
              for (int b = 0; b < nb; ++b)
                tma_load_3d(dst + b * BOX_BYTES, &tmK, bar_full + 8 * stage, (c + b) * CW, j * BC, bh, kPolicyEvictLast);
            }
            advance();
          }
        } else {
          for (int c = 0; c < nqk; ++c) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t dst = smem_ring + stage * STAGE_BYTES;
            {
              mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * BOX_BYTES);
              tma_load_3d(dst, &tmQ, bar_full + 8 * stage, c * CW, q0, bh, kPolicyEvictLast);
              tma_load_3d(dst + BOX_BYTES, &tmK, bar_full + 8 * stage, c * CW, j * BC, bh, kPolicyEvictLast);
            }
            advance();
          }
        }
      };
      auto load_v = [&](int j) {
        for (int c = 0; c < nv; c += 2) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const int nb = min(2, nv - c);
          const uint32_t dst = smem_ring + stage * STAGE_BYTES;
          {
            mbar_arrive_expect_tx(bar_full + 8 * stage, nb * BOX_BYTES);
            for (int b = 0; b < nb; ++b)
              tma_load_3d(dst + b * BOX_BYTES, &tmV, bar_full + 8 * stage, col0 + (c + b) * CW, j * BC, bh, kPolicyEvictLast);
          }
          advance();
        }
      };
      load_qk(0);
      for (int j = 0; j < T; ++j) {
        if (!serial && j + 1 < T) load_qk(j + 1);
        load_v(j);
        if (serial && j + 1 < T) load_qk(j + 1);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer
    // one lane is elected once and runs the whole role loop alone (see fa2_fwd_tcgen05.cu / tools/ubench/ubench_attn.cu)
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, BC, true, false, false);
      constexpr uint64_t qk_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
      constexpr uint64_t v_hi = make_smem_desc_hi(BOX_BYTES, 1024, kSwizzle128B);
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == stages) { stage = 0; phase ^= 1; }
      };
      auto issue_s = [&](int buf) {
        const uint32_t d_tmem = tmem_base + (buf ? S_COL1 : S_COL0);
        if (Q_RESIDENT) {
          for (int c = 0; c < nqk; c += 2) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t kb = smem_ring + stage * STAGE_BYTES;
            const int nb = min(2, nqk - c);
            {
              for (int b = 0; b < nb; ++b) {
                const uint32_t qa = smem_q + (c + b) * BOX_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ss<1>(d_tmem, smem_desc(qk_hi, qa + k * 32), smem_desc(qk_hi, kb + b * BOX_BYTES + k * 32),
                             idesc_s, (c + b + k) != 0 ? 1u : 0u);
              }
              umma_commit(bar_empty + 8 * stage);
              if (c + 2 >= nqk) umma_commit(bar_s_full + 8 * buf);
            }
            advance();
          }
        } else {
          for (int c = 0; c < nqk; ++c) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t qa = smem_ring + stage * STAGE_BYTES;
            const uint32_t kb = qa + BOX_BYTES;
            {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss<1>(d_tmem, smem_desc(qk_hi, qa + k * 32), smem_desc(qk_hi, kb + k * 32), idesc_s,
                           (c + k) != 0 ? 1u : 0u);
              umma_commit(bar_empty + 8 * stage);
              if (c + 1 >= nqk) umma_commit(bar_s_full + 8 * buf);
            }
            advance();
          }
        }
      };
      auto issue_pv = [&](int buf, bool accumulate, bool last_tile) {
        const uint32_t p_tmem = tmem_base + (buf ? S_COL1 : S_COL0);
        for (int c = 0; c < nv; c += 2) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const int nb = min(2, nv - c);
          const uint32_t va = smem_ring + stage * STAGE_BYTES;
          const uint32_t idesc_o = make_idesc_f16(128, uint32_t(nb * CW), true, false, true);
          const uint32_t d_tmem = tmem_base + O_COL + c * CW;
          {
#pragma unroll
            for (int k = 0; k < BC / 16; ++k)
              umma_ts<1>(d_tmem, p_tmem + k * 8, smem_desc(v_hi, va + k * 16 * 128), idesc_o,
                         (accumulate || k != 0) ? 1u : 0u);
            umma_commit(bar_empty + 8 * stage);
            if (c + 2 >= nv) {  // whole PV_j issued
              umma_commit(bar_pv_done);
              if (last_tile) umma_commit(bar_o_full);
            }
          }
          advance();
        }
      };
      if (Q_RESIDENT) {
        mbar_wait(bar_q_full, 0);
        tc_fence_after();
      }
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (!serial && j + 1 < T) issue_s((j + 1) & 1);
        mbar_wait(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        issue_pv(j & 1, j > 0, j == T - 1);
        if (serial && j + 1 < T) issue_s((j + 1) & 1);  // debugging order: no QK^T / softmax overlap
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------------------------- softmax + epilogue
    constexpr int COLS = 128 / SPLIT;                    // score columns handled by one thread
    const uint32_t q = warp & 3;                         // TMEM lane quadrant = which 32 query rows
    const uint32_t half = (SPLIT == 2) ? ((warp - 4) >> 2) : 0u;  // which column half of the score tile
    const uint32_t row = q * 32 + lane;
    const uint32_t lane_base = (q * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_base + O_COL;
    float m_ref = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      const int buf = j & 1;
      const uint32_t s_tmem = tmem_base + lane_base + (buf ? S_COL1 : S_COL0);
      mbar_wait(bar_s_full + 8 * buf, (j >> 1) & 1);
      tc_fence_after();
      if (dbg_skip == 2) {  // timing probe: hand the (garbage) P back at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_full + 8 * buf);
        l = 1.f;
        continue;
      }
      uint32_t sr[COLS];
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) tmem_ld_32x32b_x32(s_tmem + half * COLS + c * 32, sr + c * 32);
      tmem_wait_ld();
      float* s = reinterpret_cast<float*>(sr);
      if (j == T - 1 && (N % BC) != 0) {
        // ragged last tile only.  The empty asm keeps this a real (warp-uniform) branch: if-converted, the 2 x BC
        // compare/select instructions would run on every tile.
        asm volatile("" ::: "memory");
        const int valid = N - j * BC - int(half) * COLS;
#pragma unroll
        for (int c = 0; c < COLS; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      float mx = row_max<COLS>(s);
      if constexpr (SPLIT == 2) {
        // row max over both halves: both threads of a row end up with the identical value
        float* slot = xch + (j & 1) * 256;
        slot[half * 128 + row] = mx;
        named_bar_sync(1 + q, 64);
        mx = fmaxf(mx, slot[(half ^ 1) * 128 + row]);
      }
      mx *= scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = mx > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O may only be touched once PV_{j-1} has completed.  (S_j was issued BEFORE PV_{j-1}, so s_full says
          // nothing about it.)  S_j complete implies PV_{j-2} complete (in-order pipe) and PV_j cannot start before
          // this thread arrives on p_full, so bar_pv_done has completed exactly j-1 or j phases here: no aliasing.
          mbar_wait(bar_pv_done, (j - 1) & 1);
          tc_fence_after();
          const float m_new = need ? mx : m_ref;
          const float alpha = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l *= alpha;
          // with SPLIT = 2 the two threads of a row rescale alternate 64-column chunks of O
          for (int cc = int(half); cc < nv; cc += SPLIT) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t orr[16];
              tmem_ld_32x32b_x16(o_tmem + cc * 64 + c * 16, orr);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) orr[e] = __float_as_uint(__uint_as_float(orr[e]) * alpha);
              tmem_st_32x32b_x16(o_tmem + cc * 64 + c * 16, orr);
            }
          }
          tmem_wait_st();
        }
      }
      // processed in blocks of 16 with packed fp32x2 arithmetic (FFMA2 / FADD2: one issue slot per two elements);
      // all FFMA2s of a block, then its MUFU.EX2s, then sums / packs, so 16 independent exponentials are in flight
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      float neg_m = -m_ref;
      const float2 scale2 = make_float2(scale_log2, scale_log2), negm2 = make_float2(neg_m, neg_m);
#pragma unroll
      for (int c0 = 0; c0 < COLS; c0 += 16) {
        float2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = ffma2(make_float2(s[c0 + 2 * e], s[c0 + 2 * e + 1]), scale2, negm2);
        if (dbg_skip != 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            x[e].x = fast_exp2(x[e].x);
            x[e].y = fast_exp2(x[e].y);
          }
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
#pragma unroll
      for (int c = 0; c < COLS / 64; ++c) tmem_st_32x32b_x32(s_tmem + half * (COLS / 2) + c * 32, sr + c * 32);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full + 8 * buf);
    }
    // ---- epilogue
    if constexpr (SPLIT == 2) {
      // total row sum = the two partial sums (the slot of parity T&1 was last used at tile T-2: free again)
      float* slot = xch + (T & 1) * 256;
      slot[half * 128 + row] = l;
      named_bar_sync(1 + q, 64);
      l += slot[(half ^ 1) * 128 + row];
    }
    mbar_wait(bar_o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const uint32_t stage_base = smem_ring + q * 32 * 128;  // ring memory is idle now: [chunk][128 rows][128 B]
    for (int cc = int(half); cc < nv; cc += SPLIT) {        // 64-column chunks of O owned by this thread
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t orr[32];
        tmem_ld_32x32b_x32(o_tmem + cc * 64 + h2 * 32, orr);
        tmem_wait_ld();
        const int sub0 = h2 * 4;
        const uint32_t row_addr = stage_base + cc * BOX_BYTES + lane * 128;
        const uint32_t xr = lane & 7;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float* f = reinterpret_cast<const float*>(orr + 8 * g);
          st_shared_v4(row_addr + (((sub0 + g) ^ xr) << 4), pack_half2(f[0] * inv_l, f[1] * inv_l),
                       pack_half2(f[2] * inv_l, f[3] * inv_l), pack_half2(f[4] * inv_l, f[5] * inv_l),
                       pack_half2(f[6] * inv_l, f[7] * inv_l));
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    const int row0 = q0 + int(q) * 32;
    if (lane == 0 && row0 < N) {
      for (int cc = int(half); cc < nv; cc += SPLIT) tma_store_3d(&tmO, stage_base + cc * BOX_BYTES, col0 + cc * CW, row0, bh);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, ffpa::TMEM_COLS);
}

// CTA-pair variant for D % 256 == 0 (ffpa2_fwd_tcgen05.cu)
int launch_ffpa_2cta(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N, int64_t D,
                     float scale, cudaStream_t s);
int launch_ffpa_otrans(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N, int64_t D,
                       float scale, cudaStream_t s);

}  // namespace b200k

extern "C" int b200k_ffpa_fwd_f16(const void* Q, const void* K, const void* V, void* O, int64_t B, int64_t H, int64_t N,
                                  int64_t D, float scale, int variant, void* stream) {
  using namespace b200k;
  if (D == 32 || D == 64 || D == 96 || (D == 128 && !(variant & 16)))  // variant bit 16: run D=128 on this kernel (experiment)
    return b200k_fa2_fwd_f16(Q, K, V, O, B, H, N, D, scale, 0, variant, stream);
  if (!Q || !K || !V || !O) return set_error(B200K_EARG, "b200k_ffpa_fwd_f16: null pointer");
  // 160, 224, ... (D % 64 == 32) are the reference's ENABLE_FFPA_ALL_HEADDIM rungs (launch_templates.cuh:L483-552)
  if (D < 128 || D > 1024 || (D % 32) != 0)
    return set_error(B200K_EHEADDIM, "headdim not support! (b200k_ffpa_fwd_f16: D=%lld; supported 32 .. 1024 step 32)", (long long)D);
  if (B < 1 || H < 1 || N < 1 || N > INT32_MAX || B * H > 65535)
    return set_error(B200K_ESHAPE, "b200k_ffpa_fwd_f16: need B,H,N >= 1 and B*H <= 65535");
  if (scale <= 0.f) scale = 1.0f / sqrtf(float(D));
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // variant bits: 1 = force the single-CTA kernel, 2 = stream Q even when it fits, 4 = serial debug order
  // CTA-pair kernel for D = 512 / 768 / 1024 (measured on (1,32,4096,D): D=512 994 vs 836 TFLOP/s; at D=256 the
  // single-CTA kernel is ahead, 1.21 vs 1.16 PFLOP/s — profiles/r01_ffpa_check.jsonl); variant bit 8 forces it.
  // D = 512 (BASELINE config #4): the O^T kernel (ffpa3_fwd_tcgen05.cu), which computes S once per KV tile instead of once
  // per 256-column slice of O: 1.40 vs 0.93-1.01 PFLOP/s on (1,32,4096,512).  Variant bit 0x200 forces it (also for D = 256,
  // where the S-recompute-free single-CTA kernel is faster: 1.32-1.40 vs 1.11), bit 0x400 keeps the D-sliced pair kernel.
  if (((variant & 0x200) && (D == 256 || D == 512)) || (D == 512 && !(variant & (0x400 | 7 | 8))))
    return launch_ffpa_otrans(Q, K, V, O, B, H, N, D, scale, s);
  if ((D % 256) == 0 && (D >= 512 || (variant & 8)) && !(variant & 7))
    return launch_ffpa_2cta(Q, K, V, O, B, H, N, D, scale, s);
  const uint64_t BH = uint64_t(B) * uint64_t(H);
  CUtensorMap tmQ, tmK, tmV, tmO;
  if ((rc = make_tmap_3d_u16(&tmQ, Q, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmK, K, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmV, V, BH, N, D, uint64_t(N) * D, D, 1, 128, 64, 128))) return rc;
  if ((rc = make_tmap_3d_u16(&tmO, O, BH, N, D, uint64_t(N) * D, D, 1, 32, 64, 128))) return rc;
  const bool q_resident = (D <= 512) && !(variant & 2);  // variant 2 forces the streaming-Q path (testing)
  const int q_bytes = q_resident ? int((D + 63) / 64) * ffpa::BOX_BYTES : 0;
  int stages = (232448 - 1024 - ffpa::BAR_BYTES - q_bytes) / ffpa::STAGE_BYTES;
  if (stages > ffpa::MAX_STAGES) stages = ffpa::MAX_STAGES;
  const int smem = 1024 + ffpa::BAR_BYTES + q_bytes + stages * ffpa::STAGE_BYTES;
  const int slices = int((D + 255) / 256);
  dim3 grid(unsigned((N + 127) / 128), unsigned(slices), unsigned(BH));
  const float scale_log2 = scale * 1.4426950408889634f;
  // Two softmax warpgroups per tile (two threads per query row) is an option (variant bit 32).  It paid off before the
  // row max moved to FMNMX3 and the role loops to a single elected lane; since then one thread per row is faster:
  // D = 256 1312 vs 1234 TFLOP/s, D = 320 793 vs 748, D = 384 837 vs 786 (same-process round robin, (1,32,8192,D)).
  // The kernel is bound by the tensor pipe's operand feed here, not by the softmax: with the softmax skipped entirely
  // (probe bit 128) D = 256 runs at 1270.
  const bool split = (variant & 32) != 0;
  const int serial = ((variant & 4) ? 1 : 0) | (((variant >> 6) & 3) << 1);  // bits 6,7: timing probes
  if (q_resident && split) {
    auto kern = ffpa_fwd_tcgen05_kernel<true, 2>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, smem)) return rc;
    kern<<<grid, 384, smem, s>>>(tmQ, tmK, tmV, tmO, int(N), int(D), 256, stages, scale_log2, serial);
  } else if (q_resident) {
    auto kern = ffpa_fwd_tcgen05_kernel<true, 1>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, smem)) return rc;
    kern<<<grid, 256, smem, s>>>(tmQ, tmK, tmV, tmO, int(N), int(D), 256, stages, scale_log2, serial);
  } else {
    auto kern = ffpa_fwd_tcgen05_kernel<false, 1>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, smem)) return rc;
    kern<<<grid, 256, smem, s>>>(tmQ, tmK, tmV, tmO, int(N), int(D), 256, stages, scale_log2, serial);
  }
  B200K_CHECK_CUDA(cudaGetLastError());
  return B200K_OK;
}
