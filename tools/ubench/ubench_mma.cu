// Micro-benchmark: issue rate / execution time of tcgen05.mma shapes on one SM (or one CTA pair), no TMA traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cuda-learn-notes_b200/csrc ubench_mma.cu -o ubench_mma
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace b200k;

// mode: 0 = SS (A,B smem), 1 = TS (A tmem, B smem).  CG = 1 or 2.  N, then MN-major B flag.
template <int CG>
__global__ void __launch_bounds__(128, 1) k(int mode, int M, int N, int b_mn, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16;
  const uint32_t sa = base + 1024, sb = sa + 65536;
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc<CG>(slot, 512); tmem_relinquish<CG>(); }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(M, N, true, false, b_mn != 0);
    const uint64_t a_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
    const uint64_t b_hi = b_mn ? make_smem_desc_hi(16384, 1024, kSwizzle128B) : make_smem_desc_hi(16, 1024, kSwizzle128B);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const int k = i & 3;
          const uint64_t ad = smem_desc(a_hi, sa + (i & 4) * 4096 + k * 32);
          const uint64_t bd = smem_desc(b_hi, sb + (i & 4) * 4096 + (b_mn ? k * 2048 : k * 32));
          if (mode == 0) umma_ss<CG>(tmem_base + (i & 1) * 0, ad, bd, idesc, 1);
          else if (mode == 1) umma_ts<CG>(tmem_base + 256, tmem_base + k * 8, bd, idesc, 1);
          else {
            // mode 2: stage the A slice smem -> TMEM with tcgen05.cp (128 lanes x 256 bit), then TS-form MMA on it
            const uint32_t a_t = tmem_base + 384 + (i & 7) * 8;
            if (CG == 1) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_t), "l"(ad) : "memory");
            else asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(a_t), "l"(ad) : "memory");
            umma_ts<CG>(tmem_base, a_t, bd, idesc, 1);
          }
        }
        if (CG == 2) umma_commit_2sm(bar, 0b01); else umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

// Attention-like MMA pattern on one SM, no TMA / softmax: per "tile" 8 SS MMAs [128 x 128 x 16] (S = Q K^T, D = 128)
// into alternating S buffers, then 8 TS MMAs [128 x 128 x 16] (O += P V) — operand addresses as in the real kernel.
__global__ void __launch_bounds__(128, 1) k_attn(int tiles, int d, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16;
  const uint32_t sq = base + 1024, sk = sq + 32768, sv = sk + 65536;  // Q 32 KB, K 2 x 32 KB, V 2 x 32 KB
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc<1>(slot, 512); tmem_relinquish<1>(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1) {
    const uint32_t idesc_s = make_idesc_f16(128, 128, true, false, false);
    const uint32_t idesc_o = make_idesc_f16(128, d, true, false, true);
    const uint64_t qk_hi = make_smem_desc_hi(16, 1024, kSwizzle128B);
    const uint64_t v_hi = make_smem_desc_hi(16384, 1024, kSwizzle128B);
    const int ksteps = d / 16;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int t = 0; t < tiles; ++t) {
          const uint32_t kb = sk + (t & 1) * 32768, vb = sv + (t & 1) * 32768;
          const uint32_t s_t = tmem_base + (t & 1) * 128;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
            umma_ss<1>(s_t, smem_desc(qk_hi, sq + off), smem_desc(qk_hi, kb + off), idesc_s, k != 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts<1>(tmem_base + 256, s_t + k * 8, smem_desc(v_hi, vb + k * 2048), idesc_o, 1);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 1024 + 1024 + 2 * 65536 + 32768;
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  {
    const int smem_a = 1024 + 1024 + 32768 + 2 * 65536;
    cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
    for (int dd : {64, 128}) {
      k_attn<<<148, 128, smem_a>>>(256, dd, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("attention MMA pattern D=%d: %.1f clk per 128x128 tile (floor %d)  %s\n", dd, double(cyc) / 256, dd / 16 * 64 + 8 * dd / 2, cudaGetErrorString(e));
    }
  }
  const int iters = 2048;
  struct C { int cg, mode, M, N, bmn; const char* name; } cases[] = {
    {1, 0, 128, 256, 0, "1cta SS M128 N256 K-major B"}, {1, 0, 128, 128, 0, "1cta SS M128 N128"}, {1, 0, 128, 64, 0, "1cta SS M128 N64"},
    {1, 0, 128, 256, 1, "1cta SS M128 N256 MN-major B"}, {1, 0, 128, 128, 1, "1cta SS M128 N128 MN-major B"},
    {1, 1, 128, 128, 1, "1cta TS M128 N128 (PV, D=128)"}, {1, 1, 128, 64, 1, "1cta TS M128 N64 (PV, D=64)"}, {1, 1, 128, 256, 1, "1cta TS M128 N256"},
    {1, 2, 128, 256, 0, "1cta cp+TS M128 N256"}, {1, 2, 128, 128, 0, "1cta cp+TS M128 N128"}, {1, 2, 128, 64, 0, "1cta cp+TS M128 N64"},
    {2, 2, 256, 256, 0, "2cta cp+TS M256 N256"}, {2, 2, 256, 192, 0, "2cta cp+TS M256 N192"}, {2, 2, 256, 128, 0, "2cta cp+TS M256 N128"},
    {2, 0, 256, 256, 0, "2cta SS M256 N256"}, {2, 0, 256, 128, 0, "2cta SS M256 N128"}, {2, 1, 256, 256, 1, "2cta TS M256 N256"}, {2, 1, 256, 128, 1, "2cta TS M256 N128"},
  };
  for (auto& c : cases) {
    for (int grid_full = 0; grid_full < 2; ++grid_full) {
      const int ctas = grid_full ? 148 : c.cg;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = c.cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaError_t e;
      if (c.cg == 1) e = cudaLaunchKernelEx(&cfg, k<1>, c.mode, c.M, c.N, c.bmn, iters, d);
      else e = cudaLaunchKernelEx(&cfg, k<2>, c.mode, c.M, c.N, c.bmn, iters, d);
      cudaError_t e2 = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      const double per = double(cyc) / iters;
      const double floor = double(c.M > 128 ? 128 : c.M) * c.N * 16 / 4096.0 * (c.M / 128) / c.cg;  // 4096 MAC/clk/SM
      printf("%-34s grid=%3d: %7.1f clk/MMA (floor %.0f)  %s %s\n", c.name, ctas, per, floor, cudaGetErrorString(e), cudaGetErrorString(e2));
    }
  }
  return 0;
}
