"""Drop-ins for the reference's small JIT modules, one namespace object per `load(name=...)`:
elementwise_lib (kernels/elementwise/elementwise.cu:L150-168), reduce_lib (kernels/reduce/block_all_reduce.cu:L792-813),
softmax_lib (kernels/softmax/softmax.cu:L866-884), rms_norm_lib (kernels/rms-norm/rms_norm.cu:L802-813),
rope_lib (kernels/rope/rope.cu:L115-119), hist_lib (kernels/histogram/histogram.cu:L74-77),
embedding_lib (kernels/embedding/embedding.cu:L121-129), and the second set: relu / sigmoid / gelu /
swish / elu / hardswish / hardshrink, layer_norm_lib, dot_product_lib, mat_transpose_lib, sgemv_lib, hgemv_lib.
Same names, same positional signatures.

The packing suffixes (x2/x4/x8/_pack) select nothing here: one 128-bit vectorised kernel serves every variant.
Where a reference variant has *different arithmetic* the difference is kept: f16-accumulating reductions
(`*_f16` acc suffix), the f16 RMS-norm kernels' `rsqrt(sum/(K+eps))`, and RoPE's integer-division frequency.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops as _ops


def _named(name, fn):
    fn.__name__ = fn.__qualname__ = name
    return fn


# ---- elementwise
def _add(name):
    return _named(name, lambda a, b, c: _ops.elementwise_add(a, b, c))


elementwise_lib = SimpleNamespace(**{n: _add(n) for n in (
    "elementwise_add_f32", "elementwise_add_f32x4", "elementwise_add_f16", "elementwise_add_f16x2",
    "elementwise_add_f16x8", "elementwise_add_f16x8_pack")})

# ---- reduce: block_all_reduce_sum_<pack>_<acc>(x) -> 1-element tensor
_REDUCE = {
    "f32_f32": (torch.float32, False), "f32x4_f32": (torch.float32, False),
    "f16_f16": (torch.float16, True), "f16_f32": (torch.float16, False),
    "f16x2_f16": (torch.float16, True), "f16x2_f32": (torch.float16, False),
    "f16x8_pack_f16": (torch.float16, True), "f16x8_pack_f32": (torch.float16, False),
    "bf16_bf16": (torch.bfloat16, True), "bf16_f32": (torch.bfloat16, False),
    "bf16x2_bf16": (torch.bfloat16, True), "bf16x2_f32": (torch.bfloat16, False),
    "bf16x8_pack_bf16": (torch.bfloat16, True), "bf16x8_pack_f32": (torch.bfloat16, False),
    "fp8_e4m3_f16": (getattr(torch, "float8_e4m3fn", None), True),
    "fp8_e4m3x16_pack_f16": (getattr(torch, "float8_e4m3fn", None), True),
    "fp8_e5m2_f16": (getattr(torch, "float8_e5m2", None), True),
    "fp8_e5m2x16_pack_f16": (getattr(torch, "float8_e5m2", None), True),
    "i8_i32": (torch.int8, False), "i8x16_pack_i32": (torch.int8, False),
}


def _reduce(suffix, dtype, acc16):
    def fn(x):
        if x.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        return _ops.block_all_reduce_sum(x, acc_f16=acc16)

    return _named("block_all_reduce_sum_" + suffix, fn)


reduce_lib = SimpleNamespace(**{"block_all_reduce_sum_" + k: _reduce(k, d, a) for k, (d, a) in _REDUCE.items()})


# ---- softmax
def _softmax(name, dtype, mode):
    def fn(x, y):
        if x.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        _ops.softmax(x, y, mode)

    return _named(name, fn)


softmax_lib = SimpleNamespace(
    softmax_f32=_softmax("softmax_f32", torch.float32, _ops.SOFTMAX_ALL),
    softmax_f32x4=_softmax("softmax_f32x4", torch.float32, _ops.SOFTMAX_ALL),
    softmax_f32_per_token=_softmax("softmax_f32_per_token", torch.float32, _ops.SOFTMAX_PER_TOKEN),
    softmax_f32x4_per_token=_softmax("softmax_f32x4_per_token", torch.float32, _ops.SOFTMAX_PER_TOKEN),
    safe_softmax_f32_per_token=_softmax("safe_softmax_f32_per_token", torch.float32, _ops.SOFTMAX_SAFE),
    safe_softmax_f32x4_per_token=_softmax("safe_softmax_f32x4_per_token", torch.float32, _ops.SOFTMAX_SAFE),
    safe_softmax_f16_f32_per_token=_softmax("safe_softmax_f16_f32_per_token", torch.float16, _ops.SOFTMAX_SAFE),
    safe_softmax_f16x2_f32_per_token=_softmax("safe_softmax_f16x2_f32_per_token", torch.float16, _ops.SOFTMAX_SAFE),
    safe_softmax_f16x8_pack_f32_per_token=_softmax("safe_softmax_f16x8_pack_f32_per_token", torch.float16, _ops.SOFTMAX_SAFE),
    online_safe_softmax_f32_per_token=_softmax("online_safe_softmax_f32_per_token", torch.float32, _ops.SOFTMAX_ONLINE),
    online_safe_softmax_f32x4_pack_per_token=_softmax("online_safe_softmax_f32x4_pack_per_token", torch.float32, _ops.SOFTMAX_ONLINE),
)


# ---- rms norm: rms_norm_<pack>[_<acc>](x, y, g)
def _rms(name, dtype, acc16, eps_inside_k):
    def fn(x, y, g):
        if x.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        _ops.rms_norm(x, y, g, 1e-5, acc_f16=acc16, eps_inside_k=eps_inside_k)

    return _named(name, fn)


rms_norm_lib = SimpleNamespace(
    rms_norm_f32=_rms("rms_norm_f32", torch.float32, False, False),
    rms_norm_f32x4=_rms("rms_norm_f32x4", torch.float32, False, False),
    # every f16-input kernel of the reference computes rsqrt(sum/(K+eps)) (rms_norm.cu:L164,L184,L224,L264,L290,L320,L352)
    rms_norm_f16_f16=_rms("rms_norm_f16_f16", torch.float16, True, True),
    rms_norm_f16x2_f16=_rms("rms_norm_f16x2_f16", torch.float16, True, True),
    rms_norm_f16x8_f16=_rms("rms_norm_f16x8_f16", torch.float16, True, True),
    rms_norm_f16x8_f32=_rms("rms_norm_f16x8_f32", torch.float16, False, True),
    rms_norm_f16_f32=_rms("rms_norm_f16_f32", torch.float16, False, True),
    rms_norm_f16x8_pack_f16=_rms("rms_norm_f16x8_pack_f16", torch.float16, True, True),
    rms_norm_f16x8_pack_f32=_rms("rms_norm_f16x8_pack_f32", torch.float16, False, True),
)

# ---- rope: the reference kernels' behaviour (integer-division frequency), see SURVEY.md §8 a9
rope_lib = SimpleNamespace(**{n: _named(n, lambda x, out: _ops.rope_f32(x, out, ref_quirk=True))
                              for n in ("rope_f32", "rope_f32_v2", "rope_f32x4_pack")})

# ---- histogram: histogram_i32(a) -> counts tensor
hist_lib = SimpleNamespace(**{n: _named(n, lambda a: _ops.histogram_i32(a)) for n in ("histogram_i32", "histogram_i32x4")})


# ---- embedding: embedding_<pack>(idx, weight, out)
def _emb(name, dtype):
    def fn(idx, weight, out):
        if weight.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        _ops.embedding(idx, weight, out)

    return _named(name, fn)


embedding_lib = SimpleNamespace(
    embedding_f32=_emb("embedding_f32", torch.float32), embedding_f32x4=_emb("embedding_f32x4", torch.float32),
    embedding_f32x4_pack=_emb("embedding_f32x4_pack", torch.float32), embedding_f16=_emb("embedding_f16", torch.float16),
    embedding_f16x8=_emb("embedding_f16x8", torch.float16), embedding_f16x8_pack=_emb("embedding_f16x8_pack", torch.float16),
)



# ---- activations: <op>_<pack>(x, y); six entry points per family (e.g. kernels/relu/relu.cu:L151-158)
def _act_lib(op):
    def make(name, dtype):
        def fn(x, y):
            if x.dtype != dtype:
                raise RuntimeError("values must be %s" % dtype)
            _ops.activation(x, y, op, ref_clamp=True)

        return _named(name, fn)

    names = {op + "_f32": torch.float32, op + "_f32x4": torch.float32, op + "_f16": torch.float16,
             op + "_f16x2": torch.float16, op + "_f16x8": torch.float16, op + "_f16x8_pack": torch.float16}
    return SimpleNamespace(**{n: make(n, d) for n, d in names.items()})


relu_lib, sigmoid_lib, gelu_lib, swish_lib = _act_lib("relu"), _act_lib("sigmoid"), _act_lib("gelu"), _act_lib("swish")
elu_lib, hardswish_lib, hardshrink_lib = _act_lib("elu"), _act_lib("hardswish"), _act_lib("hardshrink")


# ---- layer norm: layer_norm_<pack>[_<acc>](x, y, g, b)  (kernels/layer-norm/layer_norm.cu:L732-812)
def _ln(name, dtype):
    def fn(x, y, g, b):
        if x.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        _ops.layer_norm(x, y, g, b, 1e-5, eps_inside_k=True)

    return _named(name, fn)


layer_norm_lib = SimpleNamespace(**{n: _ln(n, d) for n, d in {
    "layer_norm_f32": torch.float32, "layer_norm_f32x4": torch.float32, "layer_norm_f16_f16": torch.float16,
    "layer_norm_f16_f32": torch.float16, "layer_norm_f16x2_f16": torch.float16, "layer_norm_f16x8_f16": torch.float16,
    "layer_norm_f16x8_pack_f16": torch.float16, "layer_norm_f16x8_pack_f32": torch.float16}.items()})


# ---- dot product: dot_prod_<pack>_f32(a, b) -> 1-element f32 tensor  (kernels/dot-product/dot_product.cu:L269-283)
def _dot(name, dtype):
    def fn(a, b):
        if a.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        return _ops.dot_prod(a, b)

    return _named(name, fn)


dot_product_lib = SimpleNamespace(**{n: _dot(n, d) for n, d in {
    "dot_prod_f32_f32": torch.float32, "dot_prod_f32x4_f32": torch.float32, "dot_prod_f16_f32": torch.float16,
    "dot_prod_f16x2_f32": torch.float16, "dot_prod_f16x8_pack_f32": torch.float16}.items()})

# ---- transpose: mat_transpose_<variant>(x, y)  (kernels/mat-transpose/mat_transpose.cu:L296-359)
mat_transpose_lib = SimpleNamespace(**{n: _named(n, lambda x, y: _ops.mat_transpose(x, y)) for n in (
    "mat_transpose_f32_col2row", "mat_transpose_f32_row2col", "mat_transpose_f32x4_col2row", "mat_transpose_f32x4_row2col",
    "mat_transpose_f32_col2row2d", "mat_transpose_f32_row2col2d", "mat_transpose_f32x4_col2row2d",
    "mat_transpose_f32x4_row2col2d", "mat_transpose_f32_diagonal2d", "mat_transpose_f32x4_shared_col2row2d",
    "mat_transpose_f32x4_shared_row2col2d", "mat_transpose_f32x4_shared_bcf_col2row2d",
    "mat_transpose_f32x4_shared_bcf_row2col2d")})


# ---- gemv: sgemv_k*(a, x, y) / hgemv_k*(a, x, y) with the reference's K checks (sgemv.cu:L120-195)
def _gemv(name, dtype, k_multiple=None, k_equal=None):
    def fn(a, x, y):
        if a.dtype != dtype:
            raise RuntimeError("values must be %s" % dtype)
        K = a.size(1)
        if k_multiple and K % k_multiple:
            raise RuntimeError("K must be multiple of %d" % k_multiple)
        if k_equal and K != k_equal:
            raise RuntimeError("K must be %d" % k_equal)
        _ops.gemv(a, x, y)

    return _named(name, fn)


sgemv_lib = SimpleNamespace(sgemv_k32_f32=_gemv("sgemv_k32_f32", torch.float32, k_multiple=32),
                            sgemv_k128_f32x4=_gemv("sgemv_k128_f32x4", torch.float32, k_multiple=128),
                            sgemv_k16_f32=_gemv("sgemv_k16_f32", torch.float32, k_equal=16))
hgemv_lib = SimpleNamespace(hgemv_k32_f16=_gemv("hgemv_k32_f16", torch.float16, k_multiple=32),
                            hgemv_k128_f16x4=_gemv("hgemv_k128_f16x4", torch.float16, k_multiple=128),
                            hgemv_k16_f16=_gemv("hgemv_k16_f16", torch.float16, k_equal=16))

# name passed to torch.utils.cpp_extension.load(name=...) by each reference script -> namespace
BY_LOAD_NAME = {
    "elementwise_lib": elementwise_lib, "block_all_reduce_lib": reduce_lib, "softmax_lib": softmax_lib,
    "rms_norm_lib": rms_norm_lib, "rope_lib": rope_lib, "hist_lib": hist_lib, "embedding_lib": embedding_lib,
    "relu_lib": relu_lib, "sigmoid_lib": sigmoid_lib, "gelu_lib": gelu_lib, "swish_lib": swish_lib, "elu_lib": elu_lib,
    "hardswish_lib": hardswish_lib, "hardshrink_lib": hardshrink_lib, "layer_norm_lib": layer_norm_lib,
    "dot_product_lib": dot_product_lib, "mat_transpose_lib": mat_transpose_lib, "sgemv_lib": sgemv_lib,
    "hgemv_lib": hgemv_lib,
}
# rope.py and embedding.py call load(name="rope") / load(name="embedding") (kernels/rope/rope.py:L13-15,
# kernels/embedding/embedding.py:L11-13)
LOAD_NAME_ALIASES = {"rope": "rope_lib", "embedding": "embedding_lib"}
