// Host-side helpers shared by every translation unit behind the C ABI (include/b200k.h):
// error reporting, driver-entry-point lookup for cuTensorMapEncodeTiled (so the library has no link-time
// dependency on libcuda.so and can be dlopen'ed on a GPU-less box), device properties cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/b200k.h"

namespace b200k {

int set_error(int code, const char* fmt, ...);  // returns `code`
#define B200K_CHECK_CUDA(expr)                                                                      \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::b200k::set_error(B200K_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                                \
  } while (0)

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int max_smem_optin = 0;
};
// Properties of the current device; fails (B200K_EARCH) unless it is compute capability 10.x.
int get_device_info(DeviceInfo* out);

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and per device.  One mutex-protected table for every
// launcher in the library; a (function, device) pair is recorded only after the attribute call succeeded, and the
// recorded size only grows, so a transient failure is retried by the next launch instead of poisoning it.
int ensure_dynamic_smem(const void* func, int device, int bytes);

// 2-D row-major fp16/any-16-bit tensor map: global [rows, cols] (cols contiguous, row pitch `pitch_elems`),
// box [box_rows, box_cols], 128B swizzle when box_cols*2 == 128, else no swizzle.
int make_tmap_2d_u16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                     uint32_t box_rows, uint32_t box_cols, bool swizzle128);
// same, for 2-byte (f16 / bf16) or 4-byte (fp32) elements; always SWIZZLE_128B
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                 uint32_t box_cols, int elem_bytes, bool atom32 = false);
// 3-D variant: [d2, d1, d0] with d0 contiguous, strides in elements; swizzle_bytes in {0, 32, 64, 128}.
int make_tmap_3d_u16(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint64_t stride2,
                     uint64_t stride1, uint32_t box2, uint32_t box1, uint32_t box0, int swizzle_bytes);

}  // namespace b200k
