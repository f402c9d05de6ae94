"""Tensor-level host wrappers over the C ABI (include/b200k.h).

These are the single place where torch tensors are turned into raw pointers + sizes + the current CUDA stream.
Argument checks reproduce the reference bindings' behaviour and error strings
(kernels/hgemm/mma/basic/hgemm_mma_stage.cu:L2078-2087 CHECK_TORCH_TENSOR_DTYPE / _SHAPE,
kernels/flash-attn/utils/utils.h:L66-78, kernels/flash-attn/mma/basic/flash_attn_mma_share_qkv.cu:L860).

PyTorch is used for device memory and streams only; every computation happens in libb200k.so.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _loader as L

_lib = L.lib

_DTYPE_ENUM = {
    torch.float32: L.F32,
    torch.float16: L.F16,
    torch.bfloat16: L.BF16,
    torch.int8: L.I8,
    torch.int32: L.I32,
}
if hasattr(torch, "float8_e4m3fn"):
    _DTYPE_ENUM[torch.float8_e4m3fn] = L.FP8_E4M3
    _DTYPE_ENUM[torch.float8_e5m2] = L.FP8_E5M2

_TH_NAME = {
    torch.float16: "torch::kHalf",
    torch.float32: "torch::kFloat32",
    torch.int32: "torch::kInt32",
    torch.bfloat16: "torch::kBFloat16",
    torch.int8: "torch::kInt8",
}


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t: torch.Tensor) -> int:
    """The current CUDA stream of the tensor's device as a raw handle.  The private torch hook skips building a
    torch.cuda.Stream object (the wrapper's per-call cost matters for a 17 us GEMM); the public API is the fallback."""
    if _raw_stream is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_dtype(t: torch.Tensor, dtype: torch.dtype) -> None:
    if t.dtype != dtype:
        # same message as the reference's CHECK_TORCH_TENSOR_DTYPE
        raise RuntimeError("values must be " + _TH_NAME.get(dtype, str(dtype)))


def _check_cuda_contig(*ts: torch.Tensor) -> None:
    dev = None
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("b200k: tensors must live on a CUDA device (there is no CPU path)")
        if not t.is_contiguous():
            raise RuntimeError("b200k: tensors must be contiguous")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("b200k: tensors must be on the same device")


class _DeviceGuard:
    """The reference launches on whatever device is current; we launch on the tensors' device."""

    def __init__(self, t: torch.Tensor):
        self.dev = t.device
        self.ctx = None

    def __enter__(self):
        if self.dev.index != torch.cuda.current_device():   # the common case switches nothing
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            return self.ctx.__exit__(*a)
        return False


# ------------------------------------------------------------------------------------------------ HGEMM
def hgemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, tn: bool = False, variant: int = L.HGEMM_AUTO) -> None:
    """c[M,N] = a[M,K] @ B (in place).  ``b`` has logical shape [K,N]; for ``tn`` its storage is B^T row-major
    (the reference's ``as_col_major``, kernels/hgemm/tools/utils.py:L135-140), i.e. it arrives as a [K,N]-shaped
    view whose memory is [N,K] contiguous."""
    _check_dtype(a, torch.float16)
    _check_dtype(b, torch.float16)
    _check_dtype(c, torch.float16)
    M, K = a.size(0), a.size(1)
    N = b.size(1)
    if b.size(0) != K or c.size(0) != M or c.size(1) != N:
        raise RuntimeError("Tensor size mismatch!")
    if tn:
        # Two spellings of "storage is B^T [N,K] row-major": the reference's as_col_major() returns a CONTIGUOUS
        # [K,N]-shaped tensor holding B^T's elements (tools/utils.py:L135-140); a strided view b = Bt.t() is the other.
        _check_cuda_contig(a, b if b.is_contiguous() else b.t(), c)
    else:
        _check_cuda_contig(a, b, c)
    with _DeviceGuard(a):
        L.check(_lib.b200k_hgemm_f16(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, 1 if tn else 0, variant,
                                     _stream(a)))


def gemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, tn: bool = False, variant: int = L.HGEMM_AUTO,
         a_km: bool = False) -> None:
    """c = a @ B for fp16, bf16 (fp32 accumulation) or fp32 operands (TF32 tensor-core product, fp32 accumulation and
    output); same layouts as :func:`hgemm`.  ``a_km``: ``a`` has logical shape [M,K] but its storage is A^T [K,M] row-major
    (pass ``At.t()`` of a contiguous ``At``) - the BLAS "NT"/"TT" cases, f16 / bf16."""
    if a_km:
        if a.dtype not in (torch.float16, torch.bfloat16):
            raise RuntimeError("values must be torch::kHalf or torch::kBFloat16")
        _check_dtype(b, a.dtype)
        _check_dtype(c, a.dtype)
        M, K, N = a.size(0), a.size(1), b.size(1)
        if b.size(0) != K or c.size(0) != M or c.size(1) != N:
            raise RuntimeError("Tensor size mismatch!")
        _check_cuda_contig(a.t(), (b if b.is_contiguous() else b.t()) if tn else b, c)
        with _DeviceGuard(a):
            L.check(_lib.b200k_gemm_ex(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, 1, 1 if tn else 0,
                                       _DTYPE_ENUM[a.dtype], variant, _stream(a)))
        return
    if a.dtype not in (torch.float16, torch.bfloat16, torch.float32):
        raise RuntimeError("values must be torch::kHalf, torch::kBFloat16 or torch::kFloat32")
    _check_dtype(b, a.dtype)
    _check_dtype(c, a.dtype)
    M, K = a.size(0), a.size(1)
    N = b.size(1)
    if b.size(0) != K or c.size(0) != M or c.size(1) != N:
        raise RuntimeError("Tensor size mismatch!")
    if tn:
        _check_cuda_contig(a, b if b.is_contiguous() else b.t(), c)
    else:
        _check_cuda_contig(a, b, c)
    with _DeviceGuard(a):
        L.check(_lib.b200k_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, 1 if tn else 0, _DTYPE_ENUM[a.dtype],
                                variant, _stream(a)))


# ------------------------------------------------------------------------------------------------ attention
FA2_HEADDIMS = (32, 64, 96, 128)


def _check_qkvo(q, k, v, o, v_is_dn=False, dtype=torch.float16):
    for t in (q, k, v, o):
        _check_dtype(t, dtype)
    if q.dim() != 4:
        raise RuntimeError("Tensor size mismatch!")
    B, H, N, D = q.shape
    vshape = (B, H, D, N) if v_is_dn else (B, H, N, D)
    if tuple(k.shape) != (B, H, N, D) or tuple(v.shape) != vshape or tuple(o.shape) != (B, H, N, D):
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(q, k, v, o)
    return B, H, N, D


def fa2_fwd(q, k, v, o, scale: Optional[float] = None, v_is_dn: bool = False, variant: int = 0, causal: bool = False,
            seqlens_k: Optional[torch.Tensor] = None) -> None:
    """FA-2 forward, [B,H,N,D] fp16 (the reference's layout and dtype) or bf16.  ``causal`` and ``seqlens_k`` (int32 [B] on
    the device: valid keys per batch) are the caller-facing options of SURVEY 8(f)-4; the reference has neither."""
    dt = q.dtype if q.dtype == torch.bfloat16 else torch.float16
    B, H, N, D = _check_qkvo(q, k, v, o, v_is_dn, dt)
    if D not in FA2_HEADDIMS:
        raise RuntimeError("headdim not support!")
    sl = 0
    if seqlens_k is not None:
        _check_dtype(seqlens_k, torch.int32)
        _check_cuda_contig(seqlens_k)
        if seqlens_k.numel() != B:
            raise RuntimeError("Tensor size mismatch!")
        sl = seqlens_k.data_ptr()
    with _DeviceGuard(q):
        if dt == torch.float16 and not causal and seqlens_k is None:
            L.check(_lib.b200k_fa2_fwd_f16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D,
                                           float(scale) if scale else 0.0, 1 if v_is_dn else 0, variant, _stream(q)))
        else:
            L.check(_lib.b200k_fa2_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D,
                                       float(scale) if scale else 0.0, 1 if v_is_dn else 0, _DTYPE_ENUM[dt],
                                       1 if causal else 0, sl, variant, _stream(q)))


def ffpa_fwd(q, k, v, o, scale: Optional[float] = None, variant: int = 0) -> None:
    B, H, N, D = _check_qkvo(q, k, v, o)
    with _DeviceGuard(q):
        rc = _lib.b200k_ffpa_fwd_f16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D,
                                     float(scale) if scale else 0.0, variant, _stream(q))
    if rc == L.EHEADDIM:
        raise RuntimeError("headdim not support!")
    L.check(rc)


# ------------------------------------------------------------------------------------------------ support kernels
_ws_cache: dict = {}


def _workspace(dev: torch.device) -> torch.Tensor:
    """Reduction workspace (partials + ticket), one per (device, stream): two streams running reductions concurrently
    on one device must not share partials.  The library zeroes the ticket itself on every call."""
    key = (dev.type, dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(int(_lib.b200k_reduce_workspace_bytes()), dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def elementwise_add(a, b, c) -> None:
    if a.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(b, a.dtype)
    _check_dtype(c, a.dtype)
    if a.numel() != b.numel() or a.numel() != c.numel():
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(a, b, c)
    with _DeviceGuard(a):
        L.check(_lib.b200k_elementwise_add(a.data_ptr(), b.data_ptr(), c.data_ptr(), a.numel(), _DTYPE_ENUM[a.dtype],
                                           _stream(a)))


def block_all_reduce_sum(x: torch.Tensor, acc_f16: bool = False) -> torch.Tensor:
    """Returns a new 1-element tensor (f32, or i32 for int8 input) like the reference bindings
    (kernels/reduce/block_all_reduce.cu:L734-760)."""
    if x.dtype not in _DTYPE_ENUM or x.dtype == torch.int32:
        raise RuntimeError("values must be a float/half/bfloat16/fp8/int8 tensor")
    _check_cuda_contig(x)
    out = torch.empty(1, dtype=torch.int32 if x.dtype == torch.int8 else torch.float32, device=x.device)
    with _DeviceGuard(x):
        L.check(_lib.b200k_block_all_reduce_sum(x.data_ptr(), out.data_ptr(), x.numel(), _DTYPE_ENUM[x.dtype],
                                                1 if acc_f16 else 0, _workspace(x.device).data_ptr(), _stream(x)))
    return out


SOFTMAX_ALL, SOFTMAX_PER_TOKEN, SOFTMAX_SAFE, SOFTMAX_ONLINE = 0, 1, 2, 3


def softmax(x, y, mode: int) -> None:
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32")
    _check_dtype(y, x.dtype)
    if x.shape != y.shape:
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    if mode == SOFTMAX_ALL:
        # The total is global in this mode, so the row shape is free: the reference calls softmax_f32 with a flat
        # tensor (softmax.py:L63); fold it into rows of the largest power of two <= 4096 dividing numel so that the
        # normalisation pass runs on the whole grid instead of one CTA.
        n = x.numel()
        if x.dim() >= 2 and x.size(-1) <= 16384:
            S, H = n // x.size(-1), x.size(-1)
        else:
            H = 4096
            while H > 1 and n % H:
                H //= 2
            S, H = (n // H, H) if H >= 32 else (1, n)
    else:
        S, H = x.numel() // x.size(-1), x.size(-1)
    with _DeviceGuard(x):
        L.check(_lib.b200k_softmax(x.data_ptr(), y.data_ptr(), S, H, _DTYPE_ENUM[x.dtype], mode,
                                   _workspace(x.device).data_ptr(), _stream(x)))


def rms_norm(x, y, g: float, eps: float = 1e-5, acc_f16: bool = False, eps_inside_k: bool = False) -> None:
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32")
    _check_dtype(y, x.dtype)
    if x.shape != y.shape or x.dim() != 2:
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    with _DeviceGuard(x):
        L.check(_lib.b200k_rms_norm(x.data_ptr(), y.data_ptr(), x.size(0), x.size(1), float(g), float(eps),
                                    _DTYPE_ENUM[x.dtype], 1 if acc_f16 else 0, 1 if eps_inside_k else 0, _stream(x)))


def rope_f32(x, out, ref_quirk: bool = True) -> None:
    _check_dtype(x, torch.float32)
    _check_dtype(out, torch.float32)
    if x.shape != out.shape or x.dim() != 2:
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, out)
    with _DeviceGuard(x):
        L.check(_lib.b200k_rope_f32(x.data_ptr(), out.data_ptr(), x.size(0), x.size(1), 1 if ref_quirk else 0,
                                    _stream(x)))


def histogram_i32(a: torch.Tensor, nbins: Optional[int] = None) -> torch.Tensor:
    """Returns int32 counts of length max(a)+1, like the reference (kernels/histogram/histogram.cu:L50-68), which
    also reads max(a) back to the host to size its output."""
    _check_dtype(a, torch.int32)
    _check_cuda_contig(a)
    with _DeviceGuard(a):
        if nbins is None:
            mx = torch.empty(1, dtype=torch.int32, device=a.device)
            L.check(_lib.b200k_max_i32(a.data_ptr(), a.numel(), mx.data_ptr(), _stream(a)))
            nbins = int(mx.item()) + 1
        y = torch.empty(max(nbins, 1), dtype=torch.int32, device=a.device)
        L.check(_lib.b200k_histogram_i32(a.data_ptr(), a.numel(), y.data_ptr(), y.numel(), _stream(a)))
    return y


def embedding(idx, weight, out) -> None:
    _check_dtype(idx, torch.int32)
    if weight.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(out, weight.dtype)
    if weight.dim() != 2 or out.numel() != idx.numel() * weight.size(1):
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(idx, weight, out)
    with _DeviceGuard(idx):
        L.check(_lib.b200k_embedding(idx.data_ptr(), weight.data_ptr(), out.data_ptr(), idx.numel(), weight.size(0),
                                     weight.size(1), _DTYPE_ENUM[weight.dtype], _stream(idx)))


# ------------------------------------------------------------------------------------------------ support kernels, set 2
ACT_OPS = {"relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID, "gelu": L.ACT_GELU, "swish": L.ACT_SWISH, "elu": L.ACT_ELU,
           "hardswish": L.ACT_HARDSWISH, "hardshrink": L.ACT_HARDSHRINK}


def activation(x, y, op: str, ref_clamp: bool = True) -> None:
    """y = op(x) elementwise; f32 or f16.  ``ref_clamp`` keeps the reference's input clamp for sigmoid / gelu
    (kernels/sigmoid/sigmoid.cu:L19-22, kernels/gelu/gelu.cu:L19-22)."""
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(y, x.dtype)
    if x.numel() != y.numel():
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    with _DeviceGuard(x):
        L.check(_lib.b200k_activation(x.data_ptr(), y.data_ptr(), x.numel(), _DTYPE_ENUM[x.dtype], ACT_OPS[op],
                                      1 if ref_clamp else 0, _stream(x)))


def layer_norm(x, y, g: float, b: float, eps: float = 1e-5, eps_inside_k: bool = True) -> None:
    """Row-wise layer norm with scalar scale / bias, x[N,K] f32 or f16.  ``eps_inside_k`` = the reference's
    rsqrt(sum/(K + eps)) (kernels/layer-norm/layer_norm.cu:L69)."""
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(y, x.dtype)
    if x.shape != y.shape or x.dim() != 2:
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    with _DeviceGuard(x):
        L.check(_lib.b200k_layer_norm(x.data_ptr(), y.data_ptr(), x.size(0), x.size(1), float(g), float(b), float(eps),
                                      _DTYPE_ENUM[x.dtype], 1 if eps_inside_k else 0, _stream(x)))


def dot_prod(a, b) -> torch.Tensor:
    """Returns a new 1-element f32 tensor like the reference bindings (kernels/dot-product/dot_product.cu:L233-283)."""
    if a.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(b, a.dtype)
    if a.numel() != b.numel():
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(a, b)
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    with _DeviceGuard(a):
        L.check(_lib.b200k_dot_prod(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _DTYPE_ENUM[a.dtype],
                                    _workspace(a.device).data_ptr(), _stream(a)))
    return out


def mat_transpose(x, y) -> None:
    """y[N,M] = x[M,N]^T, f32 (kernels/mat-transpose/mat_transpose.cu:L296-339)."""
    _check_dtype(x, torch.float32)
    _check_dtype(y, torch.float32)
    if x.dim() != 2 or y.numel() != x.numel():
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    with _DeviceGuard(x):
        L.check(_lib.b200k_mat_transpose_f32(x.data_ptr(), y.data_ptr(), x.size(0), x.size(1), _stream(x)))


def transpose_16bit_batched(x: torch.Tensor, y: torch.Tensor) -> None:
    """y[..., N, M] = x[..., M, N]^T for f16 / bf16 tensors (leading dims are the batch); exact."""
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("values must be torch::kHalf or torch::kBFloat16")
    _check_dtype(y, x.dtype)
    if x.dim() < 2 or y.numel() != x.numel():
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(x, y)
    M, N = x.size(-2), x.size(-1)
    with _DeviceGuard(x):
        L.check(_lib.b200k_transpose_u16_batched(x.data_ptr(), y.data_ptr(), x.numel() // (M * N), M, N, _stream(x)))


def gemv(a, x, y) -> None:
    """y[M,1] = a[M,K] @ x[K,1], f32 or f16 with f32 accumulation (kernels/sgemv/sgemv.cu:L126-195, hgemv.cu:L130-199)."""
    if a.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("values must be torch::kFloat32 or torch::kHalf")
    _check_dtype(x, a.dtype)
    _check_dtype(y, a.dtype)
    if a.dim() != 2 or x.numel() != a.size(1) or y.numel() != a.size(0):
        raise RuntimeError("Tensor size mismatch!")
    _check_cuda_contig(a, x, y)
    with _DeviceGuard(a):
        L.check(_lib.b200k_gemv(a.data_ptr(), x.data_ptr(), y.data_ptr(), a.size(0), a.size(1), _DTYPE_ENUM[a.dtype],
                                _stream(a)))


def default_scale(D: int) -> float:
    return 1.0 / math.sqrt(D)
