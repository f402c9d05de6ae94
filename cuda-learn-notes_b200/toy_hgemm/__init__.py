"""Drop-in `toy_hgemm` module: the 38 names exported by the reference's pybind block
(kernels/hgemm/pybind/hgemm.cc:L58-107), same positional signatures, result written into the caller's `c`.

Every HGEMM name routes to the single tcgen05 kernel behind ``b200k_hgemm_f16`` ("38 names, one kernel"):
the reference's per-kernel knobs are Ampere implementation details (`stages` = cp.async ring depth, `swizzle` /
`swizzle_stride` = thread-block rasterisation) that the B200 kernel fixes internally (TMA ring depth per tile
variant, hardware 128B swizzle, grouped tile order), so they are accepted and ignored.  The `*_tn*` names take
`b` as the reference does: a [K,N]-shaped tensor whose storage is B^T [N,K] row-major
(kernels/hgemm/tools/utils.py:L135-140, hgemm.py:L317).

`hgemm.py` of the reference imports this module if it is importable (tools/utils.py:L116-121), so putting
`cuda-learn-notes_b200/` on PYTHONPATH makes the unmodified script run on these kernels.
"""
from __future__ import annotations

import torch

from b200k import ops as _ops

__version__ = "0.1.0+b200"

_NN_3ARG = [
    # kernels/hgemm/naive/hgemm.cu
    "hgemm_naive_f16", "hgemm_sliced_k_f16", "hgemm_t_8x8_sliced_k_f16x4", "hgemm_t_8x8_sliced_k_f16x4_pack",
    "hgemm_t_8x8_sliced_k_f16x4_bcf", "hgemm_t_8x8_sliced_k_f16x4_pack_bcf", "hgemm_t_8x8_sliced_k_f16x8_pack_bcf",
    "hgemm_t_8x8_sliced_k_f16x8_pack_bcf_dbuf",
    # kernels/hgemm/naive/hgemm_async.cu
    "hgemm_t_8x8_sliced_k16_f16x8_pack_dbuf", "hgemm_t_8x8_sliced_k16_f16x8_pack_dbuf_async",
    "hgemm_t_8x8_sliced_k32_f16x8_pack_dbuf", "hgemm_t_8x8_sliced_k32_f16x8_pack_dbuf_async",
    "hgemm_t_16x8_sliced_k32_f16x8_pack_dbuf", "hgemm_t_16x8_sliced_k32_f16x8_pack_dbuf_async",
    # kernels/hgemm/cublas/hgemm_cublas.cu (NN)
    "hgemm_cublas_tensor_op_nn",
    # kernels/hgemm/wmma/hgemm_wmma.cu
    "hgemm_wmma_m16n16k16_naive", "hgemm_wmma_m16n16k16_mma4x2", "hgemm_wmma_m16n16k16_mma4x2_warp2x4",
    "hgemm_wmma_m16n16k16_mma4x2_warp2x4_dbuf_async", "hgemm_wmma_m32n8k16_mma2x4_warp2x4_dbuf_async",
    # kernels/hgemm/mma/basic/hgemm_mma.cu
    "hgemm_mma_m16n8k16_naive", "hgemm_mma_m16n8k16_mma2x4_warp4x4",
]
_TN_3ARG = ["hgemm_cublas_tensor_op_tn"]
_NN_STAGED = [
    # kernels/hgemm/wmma/hgemm_wmma_stage.cu
    "hgemm_wmma_m16n16k16_mma4x2_warp2x4_stages", "hgemm_wmma_m16n16k16_mma4x2_warp2x4_stages_dsmem",
    "hgemm_wmma_m16n16k16_mma4x2_warp4x4_stages_dsmem", "hgemm_wmma_m16n16k16_mma4x4_warp4x4_stages_dsmem",
    # kernels/hgemm/mma/basic/hgemm_mma_stage.cu (the flagship NN family)
    "hgemm_mma_m16n8k16_mma2x4_warp4x4_stages", "hgemm_mma_m16n8k16_mma2x4_warp4x4_stages_dsmem",
    "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem", "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem_x4",
    "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem_rr",
    # kernels/hgemm/mma/swizzle/hgemm_mma_stage_swizzle.cu
    "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem_swizzle",
]
_TN_STAGED = [
    "hgemm_mma_m16n8k16_mma2x4_warp4x4_stages_dsmem_tn",  # mma/basic/hgemm_mma_stage_tn.cu
    "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem_tn_swizzle_x4",  # mma/swizzle/..._tn_swizzle_x4.cu
    "hgemm_mma_stages_block_swizzle_tn_cute",  # cutlass/hgemm_mma_stage_tn_cute.cu
]


def _make3(name: str, tn: bool):
    def fn(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> None:
        _ops.hgemm(a, b, c, tn=tn)

    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = name
    return fn


def _make6(name: str, tn: bool):
    def fn(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, stages: int, swizzle: bool, swizzle_stride: int) -> None:
        _ops.hgemm(a, b, c, tn=tn)

    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = name
    return fn


def _make_cublas(name: str, tn: bool):
    """The two `hgemm_cublas_tensor_op_*` names are the reference's COMPARATOR rows (hgemm_cublas.cu:L41-84: cublasGemmEx,
    CUBLAS_COMPUTE_16F).  They stay a comparator here: the call goes to cuBLAS through torch.matmul, never to the
    tcgen05 kernel, so the "(cublas)" row of the reference's hgemm.py is what it says it is."""
    def fn(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> None:
        if a.dtype != torch.float16 or b.dtype != torch.float16 or c.dtype != torch.float16:
            raise RuntimeError("values must be torch::kHalf")
        if not (a.is_cuda and b.is_cuda and c.is_cuda):
            raise RuntimeError("b200k: tensors must live on a CUDA device (there is no CPU path)")
        if a.size(1) != b.size(0) or c.size(0) != a.size(0) or c.size(1) != b.size(1):
            raise RuntimeError("Tensor size mismatch!")
        if tn and b.is_contiguous():
            # as_col_major(): a [K,N]-shaped contiguous tensor whose memory is B^T [N,K] row-major (tools/utils.py:L135-140)
            b = b.reshape(-1).view(b.size(1), b.size(0)).t()
        torch.matmul(a, b, out=c)

    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = name + " (cuBLAS comparator)"
    return fn


for _n in _NN_3ARG:
    globals()[_n] = _make_cublas(_n, False) if "cublas" in _n else _make3(_n, False)
for _n in _TN_3ARG:
    globals()[_n] = _make_cublas(_n, True) if "cublas" in _n else _make3(_n, True)
for _n in _NN_STAGED:
    globals()[_n] = _make6(_n, False)
for _n in _TN_STAGED:
    globals()[_n] = _make6(_n, True)


def init_cublas_handle() -> None:
    """The reference creates a global cuBLAS handle here (hgemm_cublas.cu:L13-26).  Nothing to set up."""


def destroy_cublas_handle() -> None:
    """Counterpart of init_cublas_handle (hgemm_cublas.cu:L28-38).  Nothing to tear down."""


HGEMM_NAMES = _NN_3ARG + _TN_3ARG + _NN_STAGED + _TN_STAGED
__all__ = HGEMM_NAMES + ["init_cublas_handle", "destroy_cublas_handle"]
