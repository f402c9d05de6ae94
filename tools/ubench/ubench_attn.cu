// Micro-benchmark: which part of the attention pipeline costs the cycles?  One CTA per SM, D = 128, 128x128 tiles.
//   mode 0: bare MMA stream (8 SS + 8 TS per tile)                                -> tensor floor
//   mode 1: + per-tile handshake MMA warp <-> a 128-thread consumer group (s_full / p_full mbarriers), FFPA order
//   mode 2: + TMA ring: K and V tiles (32 KB each) streamed from global memory through a 5-stage smem ring
//   mode 3: mode 2 with the consumer group also reading S (tcgen05.ld 128 cols) and writing P (tcgen05.st 64 cols)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cuda-learn-notes_b200/csrc ubench_attn.cu -o ubench_attn -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "ptx.cuh"
using namespace b200k;

constexpr int STAGES = 5, STAGE_BYTES = 32768;

template <bool ONCE>
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tmK, int mode, int tiles, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = base, bar_empty = base + 64, bar_s = base + 128, bar_p = base + 144, bar_done = base + 160, slot = base + 176;
  const uint32_t sq = base + 1024, ring = sq + 32768;
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1); mbar_init(bar_p, 4); mbar_init(bar_p + 8, 4); mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc<1>(slot, 512); tmem_relinquish<1>(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  const long long t0 = clock64();
  if (warp == 0 && mode >= 2) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < 2 * tiles; ++i) {  // K(0), then K(j+1), V(j) ... same count: 2 stage fills per tile
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
        const int row = ((i * 37 + blockIdx.x * 11) % 60) * 128;
        tma_load_3d(ring + stage * STAGE_BYTES, &tmK, bar_full + 8 * stage, 0, row, 0, kPolicyEvictLast);
        tma_load_3d(ring + stage * STAGE_BYTES + 16384, &tmK, bar_full + 8 * stage, 64, row, 0, kPolicyEvictLast);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && (!ONCE || elect_one())) {
    const uint32_t idesc_s = make_idesc_f16(128, 128, true, false, false);
    const uint32_t idesc_o = make_idesc_f16(128, 128, true, false, true);
    const uint64_t qk_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
    const uint64_t v_hi = make_smem_desc_hi(16384, 1024, kSwizzle128B);
    int stage = 0; uint32_t phase = 0;
    auto next_stage = [&]() -> uint32_t {
      uint32_t addr = ring + stage * STAGE_BYTES;
      if (mode >= 2) { mbar_wait(bar_full + 8 * stage, phase); tc_fence_after(); }
      return addr;
    };
    auto release = [&]() {
      if (mode >= 2) umma_commit(bar_empty + 8 * stage);
    };
    auto adv = [&]() { if (++stage == STAGES) { stage = 0; phase ^= 1; } };
    auto issue_s = [&](int buf) {
      const uint32_t kb = next_stage();
      if (ONCE || elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk / 4) * 16384 + (kk % 4) * 32;
          umma_ss<1>(tmem_base + buf * 128, smem_desc(qk_hi, sq + off), smem_desc(qk_hi, kb + off), idesc_s, kk != 0);
        }
        release();
        if (mode >= 1) umma_commit(bar_s + 8 * buf);
      }
      if (!ONCE) __syncwarp();
      adv();
    };
    issue_s(0);
    for (int j = 0; j < tiles; ++j) {
      if (j + 1 < tiles) issue_s((j + 1) & 1);
      if (mode >= 1) { mbar_wait(bar_p + 8 * (j & 1), (j >> 1) & 1); tc_fence_after(); }
      const uint32_t vb = next_stage();
      if (ONCE || elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts<1>(tmem_base + 256, tmem_base + (j & 1) * 128 + kk * 8, smem_desc(v_hi, vb + kk * 2048), idesc_o, 1);
        release();
        if (j == tiles - 1) umma_commit(bar_done);
      }
      if (!ONCE) __syncwarp();
      adv();
    }
    mbar_wait(bar_done, 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (warp >= 4 && mode >= 1) {
    const uint32_t q = warp & 3;
    for (int j = 0; j < tiles; ++j) {
      const int buf = j & 1;
      mbar_wait(bar_s + 8 * buf, (j >> 1) & 1);
      tc_fence_after();
      if (mode >= 3) {
        uint32_t r[128];
        const uint32_t ta = tmem_base + ((q * 32) << 16) + buf * 128;
        tmem_ld_32x32b_x32(ta, r); tmem_ld_32x32b_x32(ta + 32, r + 32); tmem_ld_32x32b_x32(ta + 64, r + 64); tmem_ld_32x32b_x32(ta + 96, r + 96);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 64; ++c) r[c] = r[2 * c] ^ r[2 * c + 1];
        tmem_st_32x32b_x32(ta, r); tmem_st_32x32b_x32(ta + 32, r + 32);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * buf);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  void* buf; cudaMalloc(&buf, 8192 * 128 * 2); cudaMemset(buf, 0, 8192 * 128 * 2);
  CUtensorMap tm;
  cuuint64_t dims[3] = {128, 8192, 1}; cuuint64_t strides[2] = {256, 8192 * 256}; cuuint32_t box[3] = {64, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int smem = 1024 + 1024 + 32768 + STAGES * STAGE_BYTES;
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 256;
  for (int once = 0; once < 2; ++once)
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (once) k<true><<<148, 256, smem>>>(tm, mode, tiles, d); else k<false><<<148, 256, smem>>>(tm, mode, tiles, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      if (rep == 1) printf("elect-once=%d mode %d: %.1f clk per 128x128x128 tile (tensor floor 1024)  %s\n", once, mode, double(cyc) / tiles, cudaGetErrorString(e));
    }
  }
  return 0;
}
