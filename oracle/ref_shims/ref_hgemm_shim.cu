// TEST INFRASTRUCTURE — not part of the product.
// Builds the UNMODIFIED reference HGEMM kernel (kernels/hgemm/mma/basic/hgemm_mma_stage.cu, compiled in its own
// "bin" mode, i.e. without torch headers) into a tiny C-ABI shared object so that tests / bench.py can run the
// reference's own mma.sync kernel on the B200 next to ours.  The reference source is #included from where it lies
// under /root/reference at build time (oracle/build_ref.py passes the -I paths); nothing is copied into this repo.
#define main b200k_ref_hgemm_unused_main  // the reference file carries its own benchmark main()
#include "hgemm_mma_stage.cu"              // -> lanunch_hgemm_mma_m16n8k16_nn<K_STAGE, BLOCK_SWIZZLE_STRIDE>  (L1966-1992)
#undef main

// hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem with block swizzle; stride choices follow hgemm.py:L71-81
// (N/2 for N<=4096, N/4 above, floor 256 -> no swizzle).
extern "C" int ref_hgemm_mma_stages_dsmem_nn(void* a, void* b, void* c, int M, int N, int K, int stages,
                                             int swizzle_stride) {
  half* A = reinterpret_cast<half*>(a);
  half* B = reinterpret_cast<half*>(b);
  half* C = reinterpret_cast<half*>(c);
#define B200K_REF_CASE(S, STRIDE)                                   \
  if (stages == S && swizzle_stride == STRIDE) {                    \
    lanunch_hgemm_mma_m16n8k16_nn<S, STRIDE>(A, B, C, M, N, K);     \
    return cudaGetLastError() == cudaSuccess ? 0 : -1;              \
  }
  B200K_REF_CASE(2, 512)
  B200K_REF_CASE(2, 1024)
  B200K_REF_CASE(2, 2048)
  B200K_REF_CASE(2, 4096)
  B200K_REF_CASE(3, 512)
  B200K_REF_CASE(3, 1024)
  B200K_REF_CASE(3, 2048)
  B200K_REF_CASE(3, 4096)
  B200K_REF_CASE(4, 1024)
  B200K_REF_CASE(4, 2048)
  B200K_REF_CASE(4, 4096)
#undef B200K_REF_CASE
  return -2;  // combination not instantiated
}
