// Probe: where does tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) put its accumulator in tensor memory?
// A[R, 0] = R, A[R, 1] = 1, B[n, 0] = 1, B[n, 1] = n / 256  =>  D[R, n] = R + n / 256 (exact in fp32), so every TMEM cell
// read back with tcgen05.ld.32x32b names the (row, column) it holds.  Cells never written keep -1.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../cuda-learn-notes_b200/csrc -I../../include probe_pair_m128.cu -o probe_pair_m128
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "ptx.cuh"
using namespace b200k;

constexpr int N = 128;   // UMMA N
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* p = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + 8192, bar = base + 16384, slot = base + 16384 + 64;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(p + 16384 + 64);
  const uint32_t rank = cluster_ctarank();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill A (64 rows x 64 k, 128B swizzle, K-major) and B (64 n-rows x 64 k)
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    const uint32_t off = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
    const int R = rank * 64 + r;
    __half a = __float2half(k == 0 ? float(R) : (k == 1 ? 1.0f : 0.0f));
    __half b = __float2half(k == 0 ? 1.0f : (k == 1 ? float(R) / 256.0f : 0.0f));   // here R plays the role of n
    *reinterpret_cast<__half*>(p + off) = a;
    *reinterpret_cast<__half*>(p + 8192 + off) = b;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc<2>(slot, 128);
    tmem_relinquish<2>();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;
  // pre-fill this CTA's 128 lanes x 128 columns with -1
  {
    uint32_t m1[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) m1[c] = __float_as_uint(-1.0f);
    for (int c = 0; c < 4; ++c) tmem_st_32x32b_x32(tmem_base + ((warp * 32) << 16) + c * 32, m1);
    tmem_wait_st();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, N, 0, false, false);
      constexpr uint64_t hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
      umma_ss<2>(tmem_base, smem_desc(hi, sA), smem_desc(hi, sB), idesc, 0u);
      umma_commit_2sm(bar, 0b11);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t r[128];
  for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_base + ((warp * 32) << 16) + c * 32, r + c * 32);
  tmem_wait_ld();
  for (int c = 0; c < 128; ++c) out[(rank * 128 + threadIdx.x) * 128 + c] = __uint_as_float(r[c]);
  tc_fence_before();
  cluster_sync();
  if (warp == 0) tmem_dealloc<2>(tmem_base, 128);
}

int main() {
  float* d;
  cudaMalloc(&d, 2 * 128 * 128 * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<2, 128, 32768>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  static float h[2 * 128 * 128];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int cta = 0; cta < 2; ++cta) {
    printf("== CTA %d: for lanes 0,1,15,16,31,32,33,63,64,65,95,96,127: cells as (row,col) at columns 0,1,31,32,63,64,65,127\n", cta);
    const int lanes[] = {0, 1, 15, 16, 31, 32, 33, 63, 64, 65, 95, 96, 127};
    const int cols[] = {0, 1, 31, 32, 63, 64, 65, 127};
    for (int li = 0; li < 13; ++li) {
      printf("lane %3d:", lanes[li]);
      for (int ci = 0; ci < 8; ++ci) {
        const float v = h[(cta * 128 + lanes[li]) * 128 + cols[ci]];
        if (v < 0) printf("  [c%3d: untouched]", cols[ci]);
        else printf("  [c%3d: r%3d n%3d]", cols[ci], int(v), int((v - int(v)) * 256.0f + 0.5f));
      }
      printf("\n");
    }
    int touched = 0;
    for (int i = 0; i < 128 * 128; ++i) touched += h[cta * 128 * 128 + i] >= 0;
    printf("touched cells: %d of %d\n", touched, 128 * 128);
  }
  return 0;
}
