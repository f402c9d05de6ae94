// Micro-benchmark: issue rate of the instructions the softmax loop is made of, per SM sub-partition.
// For each op: 8 independent dependency chains per thread, 1 / 2 / 4 warps per sub-partition; prints the clocks one
// warp-instruction occupies its pipe (clk * warps_per_smsp / instructions per warp).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cuda-learn-notes_b200/csrc -I../../include ubench_pipes.cu -o ubench_pipes
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace b200k;

enum Op { EX2 = 0, F2FP, FFMA2_, FADD2_, FMNMX3_, FFMA_, LEA_, MIX };
static const char* kNames[] = {"MUFU.EX2", "F2FP.PACK_AB", "FFMA2", "FADD2", "FMNMX3", "FFMA", "LEA(shl+add)", "softmax mix"};

template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, float seed, float* out, long long* clk) {
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = seed + float(threadIdx.x + e) * 1e-3f;
  float2 b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) b[e] = make_float2(a[2 * e], a[2 * e + 1]);
  const float2 c2 = make_float2(1.0001f, 0.9999f), d2 = make_float2(1e-3f, -1e-3f);
  uint32_t u[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) u[e] = __float_as_uint(a[e]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (OP == EX2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = fast_exp2(a[e]);
      } else if (OP == F2FP) {
#pragma unroll
        for (int e = 0; e < 8; ++e) u[e] = pack_half2(__uint_as_float(u[e]), __uint_as_float(u[(e + 1) & 7] ^ 0x1000));
      } else if (OP == FFMA2_) {
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = ffma2(b[e], c2, d2);
      } else if (OP == FADD2_) {
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = fadd2(b[e], d2);
      } else if (OP == FMNMX3_) {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = fmax3(a[e], a[(e + 3) & 7] * 0.0f + float(it), seed);
      } else if (OP == FFMA_) {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = fmaf(a[e], 1.0001f, seed);
      } else if (OP == LEA_) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          asm volatile("{\n\t.reg .b32 t;\n\tshl.b32 t, %1, 23;\n\tadd.s32 %0, %0, t;\n\t}" : "+r"(u[e]) : "r"(u[(e + 1) & 7]));
      } else if (OP == MIX) {
        // the per-16-scores mix of the softmax loop: 8 FFMA2, 16 EX2, 8 FADD2, 8 F2FP
        float2 x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = ffma2(make_float2(a[e], a[(e + 1) & 7]), c2, d2);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[e].x = fast_exp2(x[e].x);
          x[e].y = fast_exp2(x[e].y);
        }
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          b[0] = fadd2(b[0], x[e]);
          b[1] = fadd2(b[1], x[e + 1]);
          u[e] ^= pack_half2(x[e].x, x[e].y);
          u[e + 1] ^= pack_half2(x[e + 1].x, x[e + 1].y);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += 1e-7f;
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  float r = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) r += a[e] + __uint_as_float(u[e] & 0xff);
#pragma unroll
  for (int e = 0; e < 4; ++e) r += b[e].x + b[e].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
static void run(float* out, long long* clk) {
  const int iters = 512;
  // instructions per warp per outer r-iteration
  const int per_r = (OP == FFMA2_ || OP == FADD2_) ? 4 : (OP == MIX ? 48 : 8);
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(iters, 0.5f, out, clk);
    cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(iters, 0.5f, out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double n = double(iters) * 8 * per_r;
    printf("%-14s %d warp(s)/SMSP: %6.2f clk per warp-instruction (pipe time %6.2f)  %s\n", kNames[OP], warps / 4,
           double(h) / n, double(h) / n / (warps / 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&clk, 148 * 8);
  run<EX2>(out, clk);
  run<F2FP>(out, clk);
  run<FFMA2_>(out, clk);
  run<FADD2_>(out, clk);
  run<FMNMX3_>(out, clk);
  run<FFMA_>(out, clk);
  run<LEA_>(out, clk);
  run<MIX>(out, clk);
  return 0;
}
