// HGEMM for B200 (sm_100a): C[M,N] = A[M,K] * B, fp16 in/out, fp32 accumulate in tensor memory.
//
// One persistent, warp-specialised kernel:
//   warp 0      TMA producer     cp.async.bulk.tensor 2-D loads of A and B k-blocks (64 wide) into a ring of
//                                128B-swizzled smem stages, completion on "full" mbarriers
//   warp 1      MMA issuer       one thread issues tcgen05.mma.kind::f16 (4 per k-block), tcgen05.commit releases the
//                                smem stage ("empty" mbarrier) and, after the last k-block, publishes the accumulator
//   warp 2      TMEM allocator   tcgen05.alloc / dealloc of 2 accumulator buffers (epilogue of tile i overlaps the
//                                main loop of tile i+1)
//   warps 4..7  epilogue         tcgen05.ld 32 lanes x 64 columns -> fp16 -> swizzled smem -> TMA store, per warp
// CG = 2 pairs two CTAs (cta_group::2, cluster 2x1x1) on one 256 x BN tile: each CTA loads its 128 rows of A and
// its half of B; the leader CTA issues the MMAs for both and multicasts the commits.
// Round 2: MT = 2 gives each CTA 256 rows (a 512 x 256 pair tile filling TMEM: a quarter less operand traffic per flop,
// which on this power-limited part is clock frequency - see GemmCfg), a stream-K remainder round (GemmPlan), a pipelined
// epilogue, A stored [K,M] as an MN-major operand (A_MN), and a %globaltimer trace hook.
//
// This replaces the reference's cp.async + ldmatrix + mma.sync.m16n8k16 kernels
//   kernels/hgemm/mma/basic/hgemm_mma_stage.cu:L590-1023 (kernel), L2380-2454 (launcher)  and their NN/TN siblings;
// the reference's `stages`, `swizzle`, `swizzle_stride` arguments map onto the ring depth (fixed per variant), the
// hardware 128B swizzle (always on) and the grouped tile rasterisation (GROUP_M) below.
#include <map>
#include <mutex>
#include <utility>

#include "abi_common.cuh"
#include "ptx.cuh"

namespace b200k {

// DT: 0 = fp16, 1 = bf16 (both kind::f16, 16-bit output), 2 = tf32 (kind::tf32 on fp32 operands, fp32 output)
// MT: 128-row accumulator blocks per CTA.  MT = 1: two accumulator buffers of BN columns, the epilogue of tile i runs
// under the main loop of tile i+1.  MT = 2 (with CG = 2: a 512 x 256 pair tile): both 256-column accumulators of a tile
// fill the 512 TMEM columns, every B stage is used by two MMAs, so a CTA takes in 48 KB of operands per 1024 tensor
// cycles instead of 32 KB per 512 - a quarter less L2 -> SM traffic per flop.  On this power-limited part that traffic
// is clock frequency: ncu on 8192^3 shows cuBLAS (nvjet 256x256 per CTA, same arrangement) moving 6.44 GB from L2 into
// the SMs at 1.44 GHz where the 256 x 256 pair tile moves 8.61 GB at 1.32 GHz with a busier tensor pipe.
template <int CG_, int BN_, bool B_MN_, int STAGES_, int DT_ = 0, int MT_ = 1, bool A_MN_ = false>
struct GemmCfg {
  // A_MN: A is stored transposed, [K,M] row-major (M contiguous) - the "NT" / "TT" BLAS cases.  Like an
  // [K,N] B it is consumed in place as an MN-major UMMA operand: boxes of BK k-rows x 64 m-elements.
  static constexpr bool A_MN = A_MN_;
  static constexpr int CG = CG_;
  static constexpr int BN = BN_;
  static constexpr bool B_MN = B_MN_;  // true: B is [K,N] row-major (N contiguous) = "NN"; false: B^T [N,K] = "TN"
  static constexpr int STAGES = STAGES_;
  static constexpr int MT = MT_;
  static constexpr int NACC = 2 / MT_;  // accumulator buffers in flight
  static constexpr int BM_CTA = 128 * MT_;
  static constexpr int BM = BM_CTA * CG;
  static constexpr int DT = DT_;
  static constexpr int ELEM = (DT_ == 2) ? 4 : 2;       // bytes per operand / output element
  static constexpr int BK = 128 / ELEM;                // one 128-byte swizzle row of K per stage: 64 halves or 32 floats
  static constexpr int UMMA_K = 32 / ELEM;             // K of one tcgen05.mma: 32 bytes
  static constexpr int ROW_ELEMS = 128 / ELEM;         // elements of a 128-byte shared-memory row (MN-major boxes, C staging)
  static constexpr int BN_CTA = BN / CG;  // rows of B (N index) each CTA stages
  static constexpr int A_BYTES = BM_CTA * 128;
  static constexpr int B_BYTES = BN_CTA * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARP_BYTES = 2 * 32 * 128;  // two 32-row x 128-byte buffers per epilogue warp
  static constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // NACC buffers of MT accumulators of BN fp32 columns
  static_assert(MT_ == 1 || MT_ == 2, "MT");
  static_assert(!(A_MN_ && (MT_ != 1 || DT_ == 2)), "A^T storage is built for the 16-bit types and MT = 1");
  static constexpr int BAR_BYTES = 1024;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + BAR_BYTES + STAGES * STAGE_BYTES + EPI_BYTES;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
  static_assert(BN_CTA % 64 == 0 && BN_CTA <= 256, "B box");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

// Grouped rasterisation: consecutive tile ids walk GROUP_M tile-rows down before moving one tile-column right, so
// the ~74 tiles in flight share few A row-panels and few B column-panels (L2 reuse).  This is the B200 counterpart of
// the reference's block-swizzle stride (hgemm_mma_stage.cu:L2339-2343, hgemm.py:L71-81).
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int* tm, int* tn) {
  const int per_group = group_m * tiles_n;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, tiles_m - first_m);
  const int local = t - g * per_group;
  *tm = first_m + local % gsz;
  *tn = local / gsz;
}

// Work decomposition.  Tiles [sk_tiles, T) are data-parallel: cluster c owns tiles sk_tiles + c + i*G.  When T is not a
// multiple of the G resident clusters the remainder r = T mod G would cost a whole extra round on r clusters while G - r
// idle (8192^3: 1024 tiles on 74 pairs = 13.84 rounds paid as 14; 4096^3: 3.46 paid as 4; 2048^3: 64 tiles on 74
// pairs).  Those r tiles are done stream-K instead: their r * KB k-blocks are cut into G contiguous, equal ranges, one
// per cluster, processed BEFORE the cluster's data-parallel tiles.  A range is shorter than one tile (r < G), so it
// touches at most two tiles: the tail k-blocks of tile j ("writer": fp32 partial sums go to a workspace slot owned by
// the cluster, then a per-warp flag is raised) and the head k-blocks of tile j+1 ("finisher": it
// holds k-block 0, runs second, and its epilogue adds the partials of the clusters that follow it before converting and
// storing).  Sums are added in a fixed order, so results are deterministic.  All clusters are co-resident (grid <= #SMs,
// one CTA per SM), which is what lets a finisher wait for its writers.
struct GemmPlan {
  int sk_tiles = 0;       // r: tiles [0, r) are stream-K
  int units_lo = 0;       // every cluster gets units_lo k-blocks of the r * KB, the first units_rem get one more
  int units_rem = 0;
  float* partials = nullptr;   // [cluster][cta rank][epilogue warp][32 rows x BN] fp32, layout private to the kernel
  uint32_t* flags = nullptr;   // [cluster][cta rank][epilogue warp]: 0 = empty, 1 = partial published; the reader lowers it again
  unsigned long long* trace = nullptr;  // debugging: globaltimer stamps, 128 per cluster (b200k_debug_set_hgemm_trace)
};

struct WorkItem {
  int tile, kb0, kb1;
  int kind;       // 0 = whole tile, 1 = writer (partial sums to the workspace), 2 = finisher (adds the writers' partials)
  int last_writer;  // finisher: clusters (cluster_id, last_writer] hold the rest of this tile
};

// The schedule functions are host-callable too: b200k_debug_hgemm_schedule() replays them for tests/test_abi.py, which checks
// coverage (every k-block of every tile exactly once) and the wait-for order (no cycle) over many shapes without a GPU.
__host__ __device__ __forceinline__ int sk_unit_begin(const GemmPlan& p, int c) { return c * p.units_lo + (c < p.units_rem ? c : p.units_rem); }
__host__ __device__ __forceinline__ int sk_unit_owner(const GemmPlan& p, int x) {
  const int big = p.units_rem * (p.units_lo + 1);
  return x < big ? x / (p.units_lo + 1) : p.units_rem + (x - big) / p.units_lo;
}
// i-th work item of cluster c (same sequence in the producer, the MMA issuer and the epilogue warps)
__host__ __device__ __forceinline__ WorkItem get_work(const GemmPlan& p, int c, int G, int num_tiles, int num_kb, int i) {
  WorkItem w;
  w.kind = -1;  // -1: no more work
  int n_sk = 0;
  if (p.sk_tiles > 0) {
    const int ub = sk_unit_begin(p, c), ue = sk_unit_begin(p, c + 1);
    if (ub < ue) {
      const int j0 = ub / num_kb, k0 = ub - j0 * num_kb;
      const int k1 = (k0 + (ue - ub) < num_kb) ? k0 + (ue - ub) : num_kb;
      const bool two = ue > (j0 + 1) * num_kb;
      n_sk = two ? 2 : 1;
      if (i < n_sk) {
        if (i == 0) { w.tile = j0; w.kb0 = k0; w.kb1 = k1; }
        else { w.tile = j0 + 1; w.kb0 = 0; w.kb1 = ue - (j0 + 1) * num_kb; }
        w.kind = (w.kb0 > 0) ? 1 : (w.kb1 < num_kb ? 2 : 0);
        w.last_writer = (w.kind == 2) ? sk_unit_owner(p, (w.tile + 1) * num_kb - 1) : c;
        return w;
      }
    }
  }
  const int t = p.sk_tiles + c + (i - n_sk) * G;
  w.tile = t; w.kb0 = 0; w.kb1 = num_kb; w.last_writer = c;
  w.kind = (t < num_tiles) ? 0 : -1;
  return w;
}

__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <class Cfg>
__global__ void __launch_bounds__(256, 1)
hgemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int tiles_m, int tiles_n,
                     int group_m, uint64_t policy_a, uint64_t policy_b, const GemmPlan plan) {
  constexpr int CG = Cfg::CG, BN = Cfg::BN, STAGES = Cfg::STAGES;
  constexpr bool B_MN = Cfg::B_MN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const uint32_t bar_full = base;                  // STAGES x 8 B
  const uint32_t bar_empty = base + 8 * STAGES;    // STAGES x 8 B
  const uint32_t bar_tfull = base + 16 * STAGES;   // 2 x 8 B   accumulator ready   (MMA -> epilogue)
  const uint32_t bar_tempty = bar_tfull + 16;      // 2 x 8 B   accumulator drained (epilogue -> MMA)
  const uint32_t tmem_slot = bar_tempty + 16;      // 4 B
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (tmem_slot - base));
  const uint32_t smem_tiles = base + Cfg::BAR_BYTES;
  const uint32_t smem_epi = smem_tiles + STAGES * Cfg::STAGE_BYTES;

  // warp index via shuffle: known warp-uniform to the compiler, so role code stays on the uniform datapath
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int num_kb = (K + Cfg::BK - 1) / Cfg::BK;
  const int num_tiles = tiles_m * tiles_n;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  unsigned long long* const trace = plan.trace ? plan.trace + size_t(cluster_id) * 128 : nullptr;
  if (trace && leader && threadIdx.x == 32) trace[0] = global_timer();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4 * CG);  // one arrival per epilogue warp of every CTA in the pair
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (trace && leader && threadIdx.x == 32) trace[1] = global_timer();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of a pair)
    // The whole warp runs the loop convergently (waits and address arithmetic stay warp-uniform); one elected lane
    // issues the expect_tx and the TMA instructions.
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const WorkItem w = get_work(plan, cluster_id, num_clusters, num_tiles, num_kb, it);
        if (w.kind < 0) break;
        int tm, tn;
        tile_coords(w.tile, tiles_m, tiles_n, group_m, &tm, &tn);
        const int m0 = tm * Cfg::BM + int(cta_rank) * Cfg::BM_CTA;
        const int n0 = tn * BN + int(cta_rank) * Cfg::BN_CTA;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t fb_local = bar_full + 8 * stage;
          const uint32_t sa = smem_tiles + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const int k0 = kb * Cfg::BK;
          if (elect_one()) {
          if constexpr (CG == 2) {
            // all bytes of both CTAs are accounted on the leader's barrier
            if (leader) mbar_arrive_expect_tx(fb_local, 2 * Cfg::STAGE_BYTES);
            const uint32_t fb = mapa(fb_local, 0);
            if constexpr (Cfg::A_MN) {
#pragma unroll
              for (int j = 0; j < Cfg::BM_CTA / Cfg::ROW_ELEMS; ++j)
                tma_load_2d_2sm(sa + j * Cfg::BK * 128, &tmA, fb, m0 + j * Cfg::ROW_ELEMS, k0, policy_a);
            } else {
              tma_load_2d_2sm(sa, &tmA, fb, k0, m0, policy_a);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int j = 0; j < Cfg::BN_CTA / Cfg::ROW_ELEMS; ++j)
                tma_load_2d_2sm(sb + j * Cfg::BK * 128, &tmB, fb, n0 + j * Cfg::ROW_ELEMS, k0, policy_b);
            } else {
              tma_load_2d_2sm(sb, &tmB, fb, k0, n0, policy_b);
            }
          } else {
            mbar_arrive_expect_tx(fb_local, Cfg::STAGE_BYTES);
            if constexpr (Cfg::A_MN) {
#pragma unroll
              for (int j = 0; j < Cfg::BM_CTA / Cfg::ROW_ELEMS; ++j)
                tma_load_2d(sa + j * Cfg::BK * 128, &tmA, fb_local, m0 + j * Cfg::ROW_ELEMS, k0, policy_a);
            } else {
              tma_load_2d(sa, &tmA, fb_local, k0, m0, policy_a);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int j = 0; j < Cfg::BN_CTA / Cfg::ROW_ELEMS; ++j)
                tma_load_2d(sb + j * Cfg::BK * 128, &tmB, fb_local, n0 + j * Cfg::ROW_ELEMS, k0, policy_b);
            } else {
              tma_load_2d(sb, &tmB, fb_local, k0, n0, policy_b);
            }
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // Whole warp convergent; tcgen05.mma / commit issued by one elected lane.
    if (leader) {
      constexpr uint32_t idesc = make_idesc(128 * CG, BN, /*operand format: f16, bf16, tf32*/ Cfg::DT, /*a_mn=*/Cfg::A_MN, /*b_mn=*/B_MN);
      // A: K-major, rows 128 B apart, 8-row groups 1024 B apart.
      constexpr uint64_t a_hi = Cfg::A_MN ? make_smem_desc_hi(Cfg::BK * 128, 1024, kSwizzle128B)
                                          : make_smem_desc_hi(16, 1024, kSwizzle128B);
      constexpr uint32_t a_kstep = Cfg::A_MN ? uint32_t(Cfg::UMMA_K) * 128u : 32u;
      // B (TN): same K-major layout.  B (NN): MN-major, one 128 B row = 64 (32 for tf32) N-elements, 8 K-rows per
      // 1024 B atom (SBO), the next row-full of N-elements one TMA box (BK rows x 128 B) further (LBO).
      // A 32-bit MN-major operand uses the "128B swizzle, 32-byte atoms" layout (UMMA layout type 1, TMA
      // SWIZZLE_128B_ATOM_32B): atoms are 4 K-rows tall, so the two atoms of one K = 8 step are SBO = 512 B apart.
      constexpr bool MN32 = B_MN && Cfg::DT == 2;
      constexpr uint64_t b_hi = MN32   ? make_smem_desc_hi(Cfg::BK * 128, 512, 1)
                                : B_MN ? make_smem_desc_hi(Cfg::BK * 128, 1024, kSwizzle128B)
                                       : make_smem_desc_hi(16, 1024, kSwizzle128B);
      constexpr uint32_t b_kstep = B_MN ? uint32_t(Cfg::UMMA_K) * 128u : 32u;  // bytes per UMMA K step
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const WorkItem w = get_work(plan, cluster_id, num_clusters, num_tiles, num_kb, it);
        if (w.kind < 0) break;
        const int acc = it % Cfg::NACC;
        const uint32_t acc_phase = (it / Cfg::NACC) & 1;
        const uint32_t d_tmem = tmem_base + acc * Cfg::MT * BN;
        // one k-block: the MMAs of accumulator blocks [mt0, mt1) on the operand stage `st`
        auto issue_kblock = [&](int st, int kb, int mt0, int mt1) {
          const uint32_t sa = smem_tiles + st * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::BK / Cfg::UMMA_K; ++k) {
            const uint64_t bdesc = smem_desc(b_hi, sb + k * b_kstep);
#pragma unroll
            for (int mt = 0; mt < Cfg::MT; ++mt) {   // rows [mt*128, mt*128+128) of this CTA's A stage, same B
              if (mt < mt0 || mt >= mt1) continue;
              const uint64_t adesc = smem_desc(a_hi, sa + mt * (128 * 128) + k * a_kstep);
              umma_ss<CG, (Cfg::DT == 2)>(d_tmem + mt * BN, adesc, bdesc, idesc, (kb > w.kb0 || k != 0) ? 1u : 0u);
            }
          }
        };
        auto commit_kblock = [&](int st, int kb) {
          if constexpr (CG == 2) umma_commit_2sm(bar_empty + 8 * st, 0b11);
          else umma_commit(bar_empty + 8 * st);
          if (kb == w.kb1 - 1) {  // accumulator complete: publish it to the epilogue warps
            if constexpr (CG == 2) umma_commit_2sm(bar_tfull + 8 * acc, 0b11);
            else umma_commit(bar_tfull + 8 * acc);
          }
        };
        int kb = w.kb0;
        if constexpr (Cfg::MT == 2) {
          // Both accumulators of a tile fill TMEM, so the epilogue of tile i cannot hide behind tile i+1 as a whole.  But
          // the epilogue drains block 0 first and hands it back on its own barrier: the first PRE k-blocks of the next
          // tile (the operand stages the producer has already filled) are issued for block 0 alone while block 1 is
          // still being read out, then for block 1, releasing the stages; the rest of the tile runs both blocks per stage.
          const int pre = min(STAGES, w.kb1 - w.kb0);
          mbar_wait(bar_tempty + 0, acc_phase ^ 1);
          tc_fence_after();
          int st = stage;
          uint32_t ph = phase;
          for (int p = 0; p < pre; ++p) {
            mbar_wait(bar_full + 8 * st, ph);
            tc_fence_after();
            if (trace && it == 0 && p == 0 && lane == 0) trace[2] = global_timer();
            if (elect_one()) issue_kblock(st, kb + p, 0, 1);
            __syncwarp();
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
          mbar_wait(bar_tempty + 8, acc_phase ^ 1);
          tc_fence_after();
          for (int p = 0; p < pre; ++p, ++kb) {
            if (elect_one()) {
              issue_kblock(stage, kb, 1, 2);
              commit_kblock(stage, kb);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
        }
        for (; kb < w.kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (trace && it == 0 && kb == w.kb0 && lane == 0) trace[2] = global_timer();
          if (elect_one()) {
            issue_kblock(stage, kb, 0, Cfg::MT);
            commit_kblock(stage, kb);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (trace && it < 56 && lane == 0) trace[8 + it] = global_timer();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp16 -> smem -> TMA store
    const uint32_t q = warp & 3;  // TMEM lane quadrant this warp may read
    const uint32_t epi = smem_epi + q * Cfg::EPI_WARP_BYTES;
    uint32_t nbuf = 0;
    for (int it = 0;; ++it) {
      const WorkItem w = get_work(plan, cluster_id, num_clusters, num_tiles, num_kb, it);
      if (w.kind < 0) break;
      int tm, tn;
      tile_coords(w.tile, tiles_m, tiles_n, group_m, &tm, &tn);
      const int acc = it % Cfg::NACC;
      const uint32_t acc_phase = (it / Cfg::NACC) & 1;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      // one chunk = 32 rows x 128 bytes of C per warp: 64 columns of 16-bit output or 32 columns of fp32
      constexpr int CC = Cfg::ROW_ELEMS;
      // stream-K: this warp's 32 x BN slice of a cluster's partial tile, stored as [chunk][16-byte unit][lane] so that
      // every warp-wide access is one contiguous 512-byte segment
      constexpr int WARP_PARTIAL = 32 * BN * Cfg::MT;  // floats
      constexpr int NCHUNK = Cfg::MT * (BN / CC);
      const size_t my_slot = (size_t(cluster_id) * CG + cta_rank) * 4 + q;
      if (w.kind == 2) {
        // finisher: wait until every writer of this tile has published this warp's slice, then lower the flag again (each
        // slot has exactly one reader and is written at most once per launch), so that every launch - and every replay
        // of a CUDA graph that captured one - starts from all-zero flags without anything host-side
        for (int cc = cluster_id + 1; cc <= w.last_writer; ++cc) {
          volatile uint32_t* f = plan.flags + (size_t(cc) * CG + cta_rank) * 4 + q;
          const long long t0 = clock64();
          while (*f == 0u) {
            if (clock64() - t0 > B200K_SPIN_LIMIT_CYCLES) __trap();
          }
          __syncwarp();   // every lane has seen the flag
          if (lane == 0) *f = 0u;
        }
        __threadfence();
      }
      // Two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is converted, staged and stored
      // (an unpipelined loop exposed the full TMEM read latency eight times per 512 x 256 tile, which has no second
      // accumulator buffer to hide its epilogue behind).
      auto load_chunk = [&](int c, uint32_t* r) {
        const uint32_t taddr = tmem_addr(tmem_base, q * 32, acc * Cfg::MT * BN + c * CC);
        tmem_ld_32x32b_x32(taddr, r);
        if constexpr (CC == 64) tmem_ld_32x32b_x32(taddr + 32, r + 32);
      };
      auto process_chunk = [&](int c, uint32_t* r) {
        const int row0 = tm * Cfg::BM + int(cta_rank) * Cfg::BM_CTA + (c / (BN / CC)) * 128 + int(q) * 32;
        if (w.kind == 1) {
          // writer: raw fp32 partial sums to this cluster's workspace slot
          uint4* dst = reinterpret_cast<uint4*>(plan.partials + my_slot * WARP_PARTIAL) + size_t(c) * (CC / 4) * 32 + lane;
#pragma unroll
          for (int j = 0; j < CC / 4; ++j) dst[j * 32] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          return;
        }
        if (w.kind == 2) {
          for (int cc = cluster_id + 1; cc <= w.last_writer; ++cc) {   // fixed order: deterministic sums
            const uint4* src = reinterpret_cast<const uint4*>(plan.partials + ((size_t(cc) * CG + cta_rank) * 4 + q) * WARP_PARTIAL) +
                               size_t(c) * (CC / 4) * 32 + lane;
#pragma unroll
            for (int j = 0; j < CC / 4; ++j) {
              const uint4 v = __ldcg(src + j * 32);
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + __uint_as_float(v.x));
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + __uint_as_float(v.y));
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + __uint_as_float(v.z));
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + __uint_as_float(v.w));
            }
          }
        }
        const uint32_t buf = epi + (nbuf & 1) * 4096;
        ++nbuf;
        if (lane == 0) tma_store_wait_read<1>();  // the store that last used this buffer has read it
        __syncwarp();
        const uint32_t row_addr = buf + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if constexpr (Cfg::DT == 2) {
            st_shared_v4(row_addr + ((j ^ (lane & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          } else {
            const float* f = reinterpret_cast<const float*>(r + 8 * j);
            if constexpr (Cfg::DT == 1)
              st_shared_v4(row_addr + ((j ^ (lane & 7)) << 4), pack_bf162(f[0], f[1]), pack_bf162(f[2], f[3]),
                           pack_bf162(f[4], f[5]), pack_bf162(f[6], f[7]));
            else
              st_shared_v4(row_addr + ((j ^ (lane & 7)) << 4), pack_half2(f[0], f[1]), pack_half2(f[2], f[3]),
                           pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int col0 = tn * BN + (c % (BN / CC)) * CC;
        if (lane == 0) {
          // One bulk group per chunk, also for chunks that lie outside C (ragged N or M): the wait_read<1> above counts
          // groups, and an uncounted chunk would let the chunk after it overwrite a buffer whose store is still reading.
          if (row0 < M && col0 < N) tma_store_2d(&tmC, buf, col0, row0);
          tma_store_commit();
        }
      };
      auto release_tmem = [&](int which) {
        // accumulator (MT = 2: accumulator block `which`) fully read: hand the TMEM columns back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster_relaxed(mapa(bar_tempty + 8 * which, 0));   // only TMEM reads, already complete
          else mbar_arrive(bar_tempty + 8 * which);
        }
      };
      static_assert(NCHUNK % 2 == 0, "chunk pairs");
      uint32_t ra[CC], rb[CC];
      load_chunk(0, ra);
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += 2) {
        tmem_wait_ld();            // chunk c is in ra
        load_chunk(c + 1, rb);
        process_chunk(c, ra);
        tmem_wait_ld();            // chunk c+1 is in rb
        if (Cfg::MT == 2 && c + 2 == NCHUNK / 2) release_tmem(0);   // block 0 is out: the next tile may start on it
        if (c + 2 < NCHUNK) load_chunk(c + 2, ra);
        else release_tmem(Cfg::MT == 2 ? 1 : acc);
        process_chunk(c + 1, rb);
      }
      if (w.kind == 1) {
        // publish: all lanes' partial stores, then the flag (the finisher's matching warp polls it)
        __threadfence();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(plan.flags + my_slot) = 1u;
      }
      if (trace && leader && warp == 4 && it < 56 && lane == 0) trace[64 + it] = global_timer();
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
    if (trace && leader && warp == 4 && lane == 0) trace[3] = global_timer();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
}

// Stream-K workspace: one per (device, stream), allocated on first use and kept (like a BLAS handle's workspace), so
// that launches on different streams never share partial sums.  Flags are zeroed at allocation and lowered by their
// reader inside the kernel, so nothing has to be cleared between launches and a captured launch can be replayed.  First
// use on a stream (or a larger problem than any before on it) calls cudaMalloc: warm up on the capture stream, at the
// largest size, before capturing a CUDA graph.
struct SkWorkspace {
  float* partials = nullptr;
  uint32_t* flags = nullptr;
  size_t partial_bytes = 0;
};
static unsigned long long* g_hgemm_trace = nullptr;  // b200k_debug_set_hgemm_trace()

static int get_sk_workspace(int device, cudaStream_t stream, size_t partial_bytes, SkWorkspace* out) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SkWorkspace> table;
  constexpr size_t kFlagBytes = 4096;  // 128 clusters x 2 CTAs x 4 warps x 4 B
  std::lock_guard<std::mutex> lock(mu);
  SkWorkspace& w = table[{device, stream}];
  if (w.partial_bytes < partial_bytes) {
    if (w.partials) B200K_CHECK_CUDA(cudaFree(w.partials));
    w = SkWorkspace();
    void* p = nullptr;
    B200K_CHECK_CUDA(cudaMalloc(&p, partial_bytes + kFlagBytes));
    B200K_CHECK_CUDA(cudaMemset(static_cast<char*>(p) + partial_bytes, 0, kFlagBytes));
    w.partials = static_cast<float*>(p);
    w.flags = reinterpret_cast<uint32_t*>(static_cast<char*>(p) + partial_bytes);
    w.partial_bytes = partial_bytes;
  }
  *out = w;   // a copy taken under the lock
  return B200K_OK;
}

// Decides whether the remainder round of this problem runs stream-K and, if so, fills the schedule part of the plan
// (sk_tiles, units_lo, units_rem); the caller then runs max_clusters clusters and attaches the workspace.
static bool plan_stream_k(int nacc, int64_t num_tiles, int num_kb, int max_clusters, int tune, GemmPlan* plan) {
  const int64_t rem_tiles = num_tiles % max_clusters;
  const bool force_sk = ((tune >> 22) & 1) != 0;   // experiments: stream-K also for a single partial round
  if (!(nacc == 2 && rem_tiles != 0 && (num_tiles > max_clusters || force_sk) && num_kb >= 8 && max_clusters <= 128 &&
        !((tune >> 20) & 1)))
    return false;
  const int64_t units = rem_tiles * num_kb;
  plan->sk_tiles = int(rem_tiles);
  plan->units_lo = int(units / max_clusters);
  plan->units_rem = int(units % max_clusters);
  return true;
}

template <class Cfg>
static int launch_hgemm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, cudaStream_t stream,
                        const DeviceInfo& di, int tune) {
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if (Cfg::A_MN) rc = make_tmap_2d(&tmA, A, K, M, M, Cfg::BK, Cfg::ROW_ELEMS, Cfg::ELEM);
  else rc = make_tmap_2d(&tmA, A, M, K, K, Cfg::BM_CTA, Cfg::BK, Cfg::ELEM);
  if (rc) return rc;
  if (Cfg::B_MN) rc = make_tmap_2d(&tmB, B, K, N, N, Cfg::BK, Cfg::ROW_ELEMS, Cfg::ELEM, /*atom32=*/Cfg::DT == 2);
  else rc = make_tmap_2d(&tmB, B, N, K, K, Cfg::BN_CTA, Cfg::BK, Cfg::ELEM);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tmC, C, M, N, N, 32, Cfg::ROW_ELEMS, Cfg::ELEM))) return rc;

  const int tiles_m = int((M + Cfg::BM - 1) / Cfg::BM);
  const int tiles_n = int((N + Cfg::BN - 1) / Cfg::BN);
  const int64_t num_tiles = int64_t(tiles_m) * tiles_n;
  const int max_clusters = di.sm_count / Cfg::CG;
  int clusters = int(num_tiles < max_clusters ? num_tiles : max_clusters);
  // Stream-K for the remainder round (see GemmPlan); tune bit 20 switches it off (A/B measurements).
  GemmPlan plan;
  plan.trace = g_hgemm_trace;
  const int num_kb = int((K + Cfg::BK - 1) / Cfg::BK);
  // Only when at least one data-parallel round follows: the finisher's fix-up (wait for the writers, read their partials
  // back from L2) then runs under the next tile's main loop.  With nothing to hide behind it costs more than it saves
  // (2048^3, 64 tiles on 74 pairs: 27.3 us stream-K vs 19.0 us plain, measured).
  // Not with MT = 2 either: a single accumulator buffer means the finisher's fix-up stalls the MMA stream (8192^3:
  // 701 us with the remainder round stream-K, 677 us without - ncu, profiles/r02_hgemm_tile_sweep_ncu.txt).
  if (plan_stream_k(Cfg::NACC, num_tiles, num_kb, max_clusters, tune, &plan)) {
    SkWorkspace ws;
    if ((rc = get_sk_workspace(di.device, stream, size_t(max_clusters) * Cfg::CG * Cfg::BM_CTA * Cfg::BN * sizeof(float), &ws))) return rc;
    clusters = max_clusters;
    plan.partials = ws.partials;
    plan.flags = ws.flags;
  }
  // tune (experiments): bits [8,16) override GROUP_M, bits [16,20) select the L2 eviction hints of the TMA loads.
  int group_m = 8;
  if ((tune >> 8) & 0xff) group_m = (tune >> 8) & 0xff;
  uint64_t policy_a = kPolicyEvictNormal, policy_b = kPolicyEvictNormal;
  switch ((tune >> 16) & 0xf) {
    case 1: policy_a = kPolicyEvictLast; policy_b = kPolicyEvictFirst; break;  // A panels are re-used by the next wave
    case 2: policy_a = kPolicyEvictLast; policy_b = kPolicyEvictNormal; break;
    case 3: policy_a = kPolicyEvictLast; policy_b = kPolicyEvictLast; break;
    default: break;
  }

  auto kern = hgemm_tcgen05_kernel<Cfg>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), di.device, Cfg::SMEM_BYTES)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * Cfg::CG);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg::CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, int(M), int(N), int(K), tiles_m, tiles_n, group_m, policy_a, policy_b, plan));
  return B200K_OK;
}

}  // namespace b200k

namespace b200k {
template <int DT>
static int gemm_dispatch(const char* who, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
                         int b_is_nk, int variant, void* stream, int a_is_km = 0) {
  constexpr int PACK = (DT == 2) ? 4 : 8;  // elements per 16 bytes
  if (!A || !B || !C) return set_error(B200K_EARG, "%s: null pointer", who);
  if (M < 1 || N < 1 || K < 1 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
    return set_error(B200K_ESHAPE, "%s: M,N,K must be in [1, 2^31) (got %lld,%lld,%lld)", who, (long long)M, (long long)N,
                     (long long)K);
  if ((K % PACK) || (N % PACK))
    return set_error(B200K_ESHAPE, "%s: K and N must be multiples of %d (16-byte rows), got K=%lld N=%lld", who, PACK,
                     (long long)K, (long long)N);
  if (a_is_km && DT == 2) return set_error(B200K_EDTYPE, "%s: A stored as [K,M] is built for f16 / bf16 only", who);
  if (a_is_km && (M % PACK))
    return set_error(B200K_ESHAPE, "%s: M must be a multiple of %d when A is stored as [K,M]", who, PACK);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int tune = variant & ~0xff;
  variant &= 0xff;
  if (variant == B200K_HGEMM_AUTO) {
    // 256x256 pair tiles (best smem/L2 traffic per flop) unless they would leave most of the machine idle; then the
    // 1-CTA 128x256 tile doubles the number of work units.  (The 256x128 pair tile is L2-bandwidth bound: measured
    // 0.59x of the 256x256 tile at 2048^3, profiles/r01_hgemm_bringup_check.jsonl.)
    const int64_t t256 = ((M + 255) / 256) * ((N + 255) / 256);
    const int64_t t512 = ((M + 511) / 512) * ((N + 255) / 256);
    variant = (t256 * 4 >= di.sm_count) ? B200K_HGEMM_2CTA_256x256 : B200K_HGEMM_1CTA_128x256;
    // 512 x 256 pair tiles once there are at least ~3 rounds of them (less operand traffic per flop: higher clocks under
    // the power cap); tune bit 21 keeps the 256 x 256 tile (A/B measurements).
    if (t512 * 2 >= 3 * di.sm_count && !((tune >> 21) & 1)) variant = B200K_HGEMM_2CTA_512x256;
  }
  const bool nn = (b_is_nk == 0);
  if (a_is_km) {
    // A^T storage: one build (256 x 256 pair tile), 16-bit types
    if constexpr (DT == 2) {
      return set_error(B200K_EDTYPE, "%s: A stored as [K,M] is built for f16 / bf16 only", who);
    } else {
      if (M % PACK) return set_error(B200K_ESHAPE, "%s: M must be a multiple of %d when A is stored as [K,M]", who, PACK);
      return nn ? launch_hgemm<GemmCfg<2, 256, true, 6, DT, 1, true>>(A, B, C, M, N, K, s, di, tune)
                : launch_hgemm<GemmCfg<2, 256, false, 6, DT, 1, true>>(A, B, C, M, N, K, s, di, tune);
    }
  }
  switch (variant) {
    case B200K_HGEMM_1CTA_128x256:
      return nn ? launch_hgemm<GemmCfg<1, 256, true, 4, DT>>(A, B, C, M, N, K, s, di, tune)
                : launch_hgemm<GemmCfg<1, 256, false, 4, DT>>(A, B, C, M, N, K, s, di, tune);
    case B200K_HGEMM_2CTA_256x256:
      return nn ? launch_hgemm<GemmCfg<2, 256, true, 6, DT>>(A, B, C, M, N, K, s, di, tune)
                : launch_hgemm<GemmCfg<2, 256, false, 6, DT>>(A, B, C, M, N, K, s, di, tune);
    case B200K_HGEMM_2CTA_512x256:
      return nn ? launch_hgemm<GemmCfg<2, 256, true, 4, DT, 2>>(A, B, C, M, N, K, s, di, tune)
                : launch_hgemm<GemmCfg<2, 256, false, 4, DT, 2>>(A, B, C, M, N, K, s, di, tune);
    case B200K_HGEMM_2CTA_256x128:
      return nn ? launch_hgemm<GemmCfg<2, 128, true, 8, DT>>(A, B, C, M, N, K, s, di, tune)
                : launch_hgemm<GemmCfg<2, 128, false, 8, DT>>(A, B, C, M, N, K, s, di, tune);
    default:
      return set_error(B200K_EARG, "%s: unknown variant %d", who, variant);
  }
}
}  // namespace b200k

// Debug hook (not part of the drop-in surface): device buffer of 128 uint64 per cluster (74 x 128 for the pair kernels)
// that every following GEMM launch fills with %globaltimer stamps: [0] kernel entry, [1] set-up done, [2] first operand
// stage landed, [3] last store done, [8+i] MMAs of work item i issued, [64+i] epilogue of item i done.  nullptr = off.
extern "C" int b200k_debug_set_hgemm_trace(void* dev_u64_buffer) {
  b200k::g_hgemm_trace = static_cast<unsigned long long*>(dev_u64_buffer);
  return B200K_OK;
}

// Debug hook (host only, no device needed): the work-item schedule the 256 x 256 pair kernel would run for `num_tiles` output
// tiles of `num_kb` k-blocks on `clusters` resident CTA pairs - the same plan_stream_k() / get_work() the launcher and the
// kernel use.  Rows of 7 int32: cluster, item index, tile, kb0, kb1, kind (0 whole tile, 1 writer, 2 finisher),
// last_writer.  Returns the number of rows (also when it exceeds `cap`; only `cap` rows are written), < 0 on bad arguments.
extern "C" int64_t b200k_debug_hgemm_schedule(int64_t num_tiles, int num_kb, int clusters, int tune, int32_t* rows, int64_t cap) {
  using namespace b200k;
  if (num_tiles < 1 || num_tiles > INT32_MAX || num_kb < 1 || clusters < 1 || (!rows && cap > 0)) return -1;
  GemmPlan plan;
  const bool sk = plan_stream_k(2, num_tiles, num_kb, clusters, tune, &plan);
  const int G = sk ? clusters : int(num_tiles < clusters ? num_tiles : clusters);
  int64_t n = 0;
  for (int c = 0; c < G; ++c) {
    for (int i = 0;; ++i) {
      const WorkItem w = get_work(plan, c, G, int(num_tiles), num_kb, i);
      if (w.kind < 0) break;
      if (n < cap) {
        int32_t* r = rows + n * 7;
        r[0] = c; r[1] = i; r[2] = w.tile; r[3] = w.kb0; r[4] = w.kb1; r[5] = w.kind; r[6] = w.last_writer;
      }
      ++n;
    }
  }
  return n;
}

extern "C" int b200k_hgemm_f16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int b_is_nk,
                               int variant, void* stream) {
  return b200k::gemm_dispatch<0>("b200k_hgemm_f16", A, B, C, M, N, K, b_is_nk, variant, stream);
}

extern "C" int b200k_gemm_ex(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int a_is_km, int b_is_nk,
                             int dtype, int variant, void* stream) {
  switch (dtype) {
    case B200K_F16: return b200k::gemm_dispatch<0>("b200k_gemm_ex(f16)", A, B, C, M, N, K, b_is_nk, variant, stream, a_is_km);
    case B200K_BF16: return b200k::gemm_dispatch<1>("b200k_gemm_ex(bf16)", A, B, C, M, N, K, b_is_nk, variant, stream, a_is_km);
    case B200K_F32: return b200k::gemm_dispatch<2>("b200k_gemm_ex(tf32)", A, B, C, M, N, K, b_is_nk, variant, stream, a_is_km);
    default: return b200k::set_error(B200K_EDTYPE, "b200k_gemm_ex: dtype %d not supported (f16, bf16, f32-as-tf32)", dtype);
  }
}

extern "C" int b200k_gemm(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int b_is_nk, int dtype,
                          int variant, void* stream) {
  switch (dtype) {
    case B200K_F16: return b200k::gemm_dispatch<0>("b200k_gemm(f16)", A, B, C, M, N, K, b_is_nk, variant, stream);
    case B200K_BF16: return b200k::gemm_dispatch<1>("b200k_gemm(bf16)", A, B, C, M, N, K, b_is_nk, variant, stream);
    case B200K_F32: return b200k::gemm_dispatch<2>("b200k_gemm(tf32)", A, B, C, M, N, K, b_is_nk, variant, stream);
    default: return b200k::set_error(B200K_EDTYPE, "b200k_gemm: dtype %d not supported (f16, bf16, f32-as-tf32)", dtype);
  }
}
