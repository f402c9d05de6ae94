// Error reporting, device query and TMA tensor-map encoding for libb200k.so (see abi_common.cuh).
#include "abi_common.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace b200k {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int get_device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  int dev = 0;
  B200K_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return set_error(B200K_ECUDA, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  DeviceInfo& c = cache[dev];
  if (c.device != dev) {
    DeviceInfo d;
    B200K_CHECK_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    B200K_CHECK_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    B200K_CHECK_CUDA(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    B200K_CHECK_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    d.device = dev;
    c = d;
  }
  if (c.cc_major != 10)
    return set_error(B200K_EARCH, "libb200k needs a compute-capability 10.x device (B200, sm_100a); device %d is %d.%d",
                     dev, c.cc_major, c.cc_minor);
  *out = c;
  return B200K_OK;
}

int ensure_dynamic_smem(const void* func, int device, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> have;
  std::lock_guard<std::mutex> lock(mu);
  auto it = have.find({func, device});
  if (it != have.end() && it->second >= bytes) return B200K_OK;
  B200K_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  have[{func, device}] = bytes;
  return B200K_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      return set_error(B200K_ECUDA, "cuTensorMapEncodeTiled not available from the driver (%s)",
                       cudaGetErrorString(e));
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return B200K_OK;
}

int make_tmap_2d_u16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                     uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((pitch_elems * 2) & 15))
    return set_error(B200K_EALIGN, "TMA needs a 16-byte aligned base (%p) and row pitch (%llu bytes)", base,
                     (unsigned long long)(pitch_elems * 2));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200K_ECUDA, "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu box=%ux%u) failed with CUresult %d",
                     (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols, (int)r);
  return B200K_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows,
                 uint32_t box_cols, int elem_bytes, bool atom32) {
  if (elem_bytes == 2) return make_tmap_2d_u16(out, base, rows, cols, pitch_elems, box_rows, box_cols, true);
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((pitch_elems * 4) & 15))
    return set_error(B200K_EALIGN, "TMA needs a 16-byte aligned base (%p) and row pitch (%llu bytes)", base,
                     (unsigned long long)(pitch_elems * 4));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200K_ECUDA, "cuTensorMapEncodeTiled(2d f32 rows=%llu cols=%llu box=%ux%u) failed with CUresult %d",
                     (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols, (int)r);
  return B200K_OK;
}

int make_tmap_3d_u16(CUtensorMap* out, const void* base, uint64_t d2, uint64_t d1, uint64_t d0, uint64_t stride2,
                     uint64_t stride1, uint32_t box2, uint32_t box1, uint32_t box0, int swizzle_bytes) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((stride1 * 2) & 15) || ((stride2 * 2) & 15))
    return set_error(B200K_EALIGN, "TMA needs 16-byte aligned base and strides");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                        : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200K_ECUDA, "cuTensorMapEncodeTiled(3d %llux%llux%llu box=%ux%ux%u) failed with CUresult %d",
                     (unsigned long long)d2, (unsigned long long)d1, (unsigned long long)d0, box2, box1, box0, (int)r);
  return B200K_OK;
}

}  // namespace b200k

extern "C" {
int b200k_abi_version(void) { return B200K_ABI_VERSION; }
const char* b200k_last_error(void) { return b200k::g_err; }
int b200k_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  b200k::DeviceInfo d;
  int rc = b200k::get_device_info(&d);
  if (rc) return rc;
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return B200K_OK;
}
}
