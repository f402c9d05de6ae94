// Micro-benchmark: tcgen05.ld / tcgen05.st cost as the softmax warps use them.
//   mode 0: per tile 4 x ld.32x32b.x32 (128 fp32 columns = 16 KB per warp), wait::ld, consume
//   mode 1: per tile 2 x st.32x32b.x32 (64 packed columns = 8 KB per warp), wait::st
//   mode 2: 8 x ld.32x32b.x16 with a wait after each (the O-rescale access pattern)
// with 4 warps (one per SM sub-partition) or 8 warps (two per sub-partition) active.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cuda-learn-notes_b200/csrc -I../../include ubench_tmem.cu -o ubench_tmem
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace b200k;

__global__ void __launch_bounds__(256, 1) k(int mode, int tiles, uint32_t* out, long long* clk) {
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc<1>(smem_u32(&slot), 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const uint32_t q = warp & 3;
  const uint32_t col0 = (warp >> 2) * 256;  // second group of warps uses other columns
  const uint32_t t = tmem_base + ((q * 32) << 16) + col0;
  uint32_t acc = threadIdx.x;
  uint32_t r[128];
#pragma unroll
  for (int c = 0; c < 128; ++c) r[c] = acc + c;
  // initialise the columns we read
  for (int c = 0; c < 8; ++c) tmem_st_32x32b_x32(t + c * 32, r + (c & 3) * 32);
  tmem_wait_st();
  __syncthreads();
  const long long t0 = clock64();
  for (int j = 0; j < tiles; ++j) {
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(t + c * 32, r + c * 32);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 128; c += 8) acc ^= r[c];
    } else if (mode == 1) {
#pragma unroll
      for (int c = 0; c < 64; c += 8) r[c] = acc + j + c;
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_st_32x32b_x32(t + 128 + c * 32, r + c * 32);
      tmem_wait_st();
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32b_x16(t + c * 16, r + c * 16);
        tmem_wait_ld();
        acc ^= r[c * 16];
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) clk[blockIdx.x * 8 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ r[5];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_base, 512);
}

int main() {
  uint32_t* out;
  long long* clk;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&clk, 148 * 8 * 8);
  const int tiles = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8}) {
      k<<<148, warps * 32>>>(mode, tiles, out, clk);
      cudaDeviceSynchronize();
      k<<<148, warps * 32>>>(mode, tiles, out, clk);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[8];
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      printf("mode %d, %d warps: %7.1f clk per tile per warp (%s)\n", mode, warps, double(h[0]) / tiles, cudaGetErrorString(e));
    }
  return 0;
}
