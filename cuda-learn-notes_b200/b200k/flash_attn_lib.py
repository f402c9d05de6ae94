"""Drop-in for the reference's JIT module `flash_attn_lib` (kernels/flash-attn/pybind/flash_attn.cc:L182-216):
25 default + 3 `--build-others` entry points, all `fn(Q, K, V, O, stages) -> None`, result written into O.

All names route to the tcgen05 attention kernels: head dims 32/64/96/128 to b200k_fa2_fwd_f16, larger ones (the
tiling_qk / tiling_qkv families accept D up to 1024, flash_attn_mma.py:L436-506) to b200k_ffpa_fwd_f16.
`stages` (the reference's cp.async ring depth) is accepted and ignored: the TMA ring depth is fixed per head dim.
The three `*_swizzle_qkv` entry points of share_kv / share_qkv / tiling_qk receive V transposed as [B,H,D,N]
(flash_attn_mma.py:L378, L542-543, L553-554, L564-565); that layout is detected from the shape.
"""
from __future__ import annotations

import torch

from . import ops as _ops

_FAMILIES = {
    # name -> max head dim the reference dispatches (flash_attn_mma.py:L436-506)
    "split_kv": 128, "split_q": 128, "split_q_shared_kv": 256, "split_q_shared_qkv": 256,
    "split_q_tiling_qk": 1024, "split_q_tiling_qkv": 1024,
    "split_q_shared_kv_acc_f32": 256, "split_q_shared_qkv_acc_f32": 256,
    "split_q_tiling_qk_acc_f32": 1024, "split_q_tiling_qkv_acc_f32": 1024,
    "split_q_shared_kv_swizzle_q": 256, "split_q_shared_kv_swizzle_qk": 256, "split_q_shared_kv_swizzle_qkv": 256,
    "split_q_shared_qkv_swizzle_q": 256, "split_q_shared_qkv_swizzle_qk": 256, "split_q_shared_qkv_swizzle_qkv": 256,
    "split_q_tiling_qk_swizzle_q": 1024, "split_q_tiling_qk_swizzle_qk": 1024, "split_q_tiling_qk_swizzle_qkv": 1024,
    "split_q_tiling_qkv_swizzle_q": 1024, "split_q_tiling_qkv_swizzle_qk": 1024, "split_q_tiling_qkv_swizzle_qkv": 1024,
    "split_q_tiling_qkv_acc_f32_swizzle_q": 1024, "split_q_tiling_qkv_acc_f32_swizzle_qk": 1024,
    "split_q_tiling_qkv_acc_f32_swizzle_qkv": 1024,
    # BUILD_FLASH_ATTN_MMA_OTHERS
    "split_q_shared_qkv_Os2g": 256, "split_q_shared_kv_acc_f32_rr": 256, "split_q_shared_qkv_acc_f32_rr": 256,
}
_V_TRANSPOSED = {"split_q_shared_kv_swizzle_qkv", "split_q_shared_qkv_swizzle_qkv", "split_q_tiling_qk_swizzle_qkv"}


def _make(short: str, max_d: int):
    name = "flash_attn_mma_stages_" + short

    def fn(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, O: torch.Tensor, stages: int) -> None:
        D = Q.size(-1)
        if D > max_d:
            raise RuntimeError("headdim not support!")
        v_is_dn = short in _V_TRANSPOSED and V.dim() == 4 and V.size(-1) == Q.size(-2) and V.size(-2) == D \
            and not (Q.size(-2) == D)
        if short in _V_TRANSPOSED and Q.size(-2) == D:
            v_is_dn = True  # square [N == D] case: these entry points always take the transposed layout
        if D <= 128:
            _ops.fa2_fwd(Q, K, V, O, v_is_dn=v_is_dn)
        else:
            if v_is_dn:
                # the large-head-dim kernel consumes V as [B,H,N,D]: transpose with the library's own kernel
                Vt = torch.empty(Q.shape, dtype=V.dtype, device=V.device)
                _ops.transpose_16bit_batched(V, Vt)
                V = Vt
            _ops.ffpa_fwd(Q, K, V, O)

    fn.__name__ = fn.__qualname__ = name
    fn.__doc__ = name
    return name, fn


NAMES = []
for _short, _maxd in _FAMILIES.items():
    _n, _f = _make(_short, _maxd)
    globals()[_n] = _f
    NAMES.append(_n)
__all__ = list(NAMES)
