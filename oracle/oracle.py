"""TEST INFRASTRUCTURE — CPU restatement of the reference algorithms (the parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module,
and only as the checker.  The product (cuda-learn-notes_b200/) never imports it; it has no CPU path at all.

Every function restates what a reference kernel computes, in fp32 (or exact integer) arithmetic on the CPU with
numpy / torch, citing the reference file:line it follows.  Floating-point kernels keep a torch fp32 reference
(the task allows this for floating-point paths); byte/integer kernels are exact.

PINNING STATUS
  * attention: pinned by the reference's own known-answer fixtures — all-ones Q/K/V -> O == 1 and the `--range-k`
    fixture (kernels/flash-attn/flash_attn_mma.py:L23-26, L353-369) — see tests/test_oracle.py, tests/golden/.
  * histogram: pinned by the reference fixture list(range(10))*1000 -> 1000 per bin (kernels/histogram/histogram.py:L22-31).
  * HGEMM and the other support kernels: the reference holds NO golden vectors or numeric checks for them
    (hgemm.py prints two elements) => unpinned by reference FIXTURES.  They are pinned instead
    on the GPU box against OUTPUTS OF THE REFERENCE ITSELF: its kernels are built unmodified from /root/reference into
    oracle/_ref/ by oracle/build_ref.py — `hgemm flash ffpa` (libref_hgemm.so, ref_flash_attn_lib.so, pyffpa_cuda.so;
    tests/test_gpu_vs_reference.py) and `support` (one ref_<op>_lib.so per bandwidth-kernel TU, 19 of them;
    tests/test_gpu_vs_reference_support.py runs every exported name of every TU against the product and, for the
    quirk modes, against this file's restatement).  The reference is CUDA-only and cannot run in the GPU-less build
    container, so these checks live in the `-m gpu` suite.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ HGEMM
def hgemm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """C = A @ B with exact (fp32) accumulation, rounded once to fp16.
    What every reference HGEMM entry point approximates (kernels/hgemm/pybind/hgemm.cc:L58-107); the bench script's
    own comparator is torch.matmul (kernels/hgemm/hgemm.py:L349)."""
    return (a.float().cpu() @ b.float().cpu()).half()


def hgemm_f16acc_k16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Reference-faithful variant: the accumulator is rounded to fp16 after every k=16 slice, as
    mma.sync.m16n8k16.f16.f16.f16.f16 does (kernels/hgemm/mma/basic/hgemm_mma_stage.cu:L51, main loop L733-876).
    Inside one k16 slice the tensor core sums exactly; small shapes only (python loop over K/16)."""
    a32, b32 = a.float().cpu(), b.float().cpu()
    M, K = a32.shape
    acc = torch.zeros(M, b32.shape[1], dtype=torch.float16)
    for k0 in range(0, K, 16):
        acc = (acc.float() + a32[:, k0:k0 + 16] @ b32[k0:k0 + 16]).half()
    return acc


# ------------------------------------------------------------------------------------------------ attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float | None = None, causal: bool = False,
              seqlens=None) -> torch.Tensor:
    """O = softmax(Q K^T * scale) V in fp32, rounded once to the input dtype: `unfused_standard_attn`
    (kernels/flash-attn/flash_attn_mma.py:L384-388); scale = 1/sqrt(D) as hard-coded by both references
    (flash_attn_mma_share_qkv.cu:L95, ffpa-attn-mma/csrc/cuffpa/launch_templates.cuh:L320).
    The reference has no masks; `causal` (row r sees keys <= r) and `seqlens` (per-batch count of valid keys, a
    key-padding mask) restate the textbook definitions for the SURVEY 8(f)-4 options of b200k_fa2_fwd."""
    q32, k32, v32 = q.float().cpu(), k.float().cpu(), v.float().cpu()
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    s = (q32 @ k32.transpose(-1, -2)) * scale
    N = q.shape[-2]
    if causal:
        keep = torch.ones(N, N, dtype=torch.bool).tril()
        s = s.masked_fill(~keep, float("-inf"))
    if seqlens is not None:
        sl = torch.as_tensor(seqlens).cpu().long().view(-1, 1, 1, 1)
        s = s.masked_fill(torch.arange(N).view(1, 1, 1, N) >= sl, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v32).to(q.dtype if q.dtype in (torch.float16, torch.bfloat16) else torch.float16)


def attention_tiled(q, k, v, Bc: int = 64, pv_acc_f16: bool = False, o_store_f16: bool = False,
                    scale: float | None = None) -> torch.Tensor:
    """Tile-faithful restatement of the reference's online softmax (FA-2): per KV tile of Bc keys
       m_new = max(m_old, rowmax(S*scale));  P = exp(S*scale - m_new) rounded to fp16 for the PV MMA;
       O = exp(m_old - m_new) * O + P V;  l = exp(m_old - m_new) * l + rowsum(P_f32);  finally O / l.
    Follows flash_attn_mma_share_qkv.cu:L403-484 (softmax), L566-675 (rescale + final 1/l) and
    ffpa-attn-mma/include/cuffpa/prefill.cuh:L273-533.  `pv_acc_f16` rounds each tile's P·V to fp16 (the f16-acc
    kernels), `o_store_f16` keeps the running O in fp16 between tiles (FFPA for D > 64,
    launch_templates.cuh:L72-80).  Small shapes only."""
    q32, k32, v32 = q.float().cpu(), k.float().cpu(), v.float().cpu()
    N, D = q32.shape[-2], q32.shape[-1]
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    lead = q32.shape[:-2]
    m = torch.full(lead + (N, 1), -float("inf"))
    l = torch.zeros(lead + (N, 1))
    o = torch.zeros(lead + (N, D))
    for j0 in range(0, N, Bc):
        s = (q32 @ k32[..., j0:j0 + Bc, :].transpose(-1, -2)) * scale
        m_new = torch.maximum(m, s.max(dim=-1, keepdim=True).values)
        p = torch.exp(s - m_new)
        alpha = torch.exp(m - m_new)
        pv = p.half().float() @ v32[..., j0:j0 + Bc, :]
        if pv_acc_f16:
            pv = pv.half().float()
        o = alpha * o + pv
        if o_store_f16:
            o = o.half().float()
        l = alpha * l + p.sum(dim=-1, keepdim=True)
        m = m_new
    return (o / l).half()


def make_range_k(B, H, N, D):
    """The reference's `--range-k` fixture: K[:, :, i, :] = (i + 1) / N (flash_attn_mma.py:L362-366)."""
    k = torch.ones(B, H, N, D, dtype=torch.half)
    for i in range(N):
        k[:, :, i, :] = (i + 1) / N
    return k


# ------------------------------------------------------------------------------------------------ support kernels
def elementwise_add(a, b):
    """c = a + b in the tensor's own dtype (IEEE round-to-nearest) — kernels/elementwise/elementwise.cu:L24-168
    (__hadd / __hadd2 for half).  Bit-exact: fp32 holds the exact sum of two halves' roundings innocuously."""
    return (a.cpu().float() + b.cpu().float()).to(a.dtype) if a.dtype != torch.float32 else a.cpu() + b.cpu()


def reduce_sum(x) -> float | int:
    """sum(x): exact int for int8 (block_all_reduce.cu:L640-686), float64 reference sum otherwise
    (the reference's atomicAdd order is non-deterministic; tests use a tolerance — block_all_reduce.cu:L42-62)."""
    if x.dtype == torch.int8:
        return int(x.cpu().to(torch.int64).sum().item())
    return float(x.cpu().float().double().sum().item())


def softmax_per_token(x):
    """Row softmax — safe/online variants all equal softmax(x, dim=-1) (kernels/softmax/softmax.cu:L150-391,
    script oracle kernels/softmax/softmax.py `torch.softmax`)."""
    return torch.softmax(x.cpu().float(), dim=-1)


def softmax_all(x):
    """softmax_f32 / softmax_f32x4: exp(x_i) / sum over the WHOLE tensor, no max subtraction (softmax.cu:L102-146)."""
    e = torch.exp(x.cpu().double())
    return (e / e.sum()).float()


def rms_norm(x, g: float = 1.0, eps: float = 1e-5, eps_inside_k: bool = False):
    """y = x * rsqrt(mean(x^2) + eps) * g (rms_norm.cu:L53-100, script oracle rms_norm.py:L25-29).
    eps_inside_k=True restates the f16-input kernels' rsqrt(sum / (K + eps)) (rms_norm.cu:L164,L184,L224,L264)."""
    xf = x.cpu().float()
    K = xf.shape[-1]
    ss = xf.pow(2).sum(-1, keepdim=True)
    denom = ss / (K + eps) if eps_inside_k else ss / K + eps
    return xf * torch.rsqrt(denom) * g


def rope(x, ref_quirk: bool, theta: float = 10000.0):
    """Interleaved-pair rotary embedding on [seq_len, hidden] f32.
    ref_quirk=False: the script's naive_rope (kernels/rope/rope.py:L71-91): angle = pos * theta^(-2i/hidden).
    ref_quirk=True : what rope_f32 / rope_f32_v2 / rope_f32x4_pack actually compute (rope.cu:L20-69): the exponent
    `token_idx / (N*2)` is an INTEGER division = 0, so every pair rotates by angle = pos."""
    xf = x.cpu().float()
    M, Hd = xf.shape
    pos = torch.arange(M, dtype=torch.float32)[:, None]
    if ref_quirk:
        ang = pos.expand(M, Hd // 2)
    else:
        inv = 1.0 / (theta ** (torch.arange(0, Hd, 2, dtype=torch.float32) / Hd))
        ang = pos * inv[None, :]
    c, s = torch.cos(ang.double()), torch.sin(ang.double())
    x1, x2 = xf[:, 0::2].double(), xf[:, 1::2].double()
    out = torch.empty_like(xf)
    out[:, 0::2] = (x1 * c - x2 * s).float()
    out[:, 1::2] = (x1 * s + x2 * c).float()
    return out


def histogram(a) -> np.ndarray:
    """Counts of each value, length max(a)+1 (kernels/histogram/histogram.cu:L18-72) — exact."""
    an = a.cpu().numpy().astype(np.int64)
    return np.bincount(an, minlength=int(an.max()) + 1).astype(np.int32)


def embedding(idx, weight):
    """out[i,:] = weight[idx[i],:] (kernels/embedding/embedding.cu:L16-119) — exact copy."""
    return weight.cpu()[idx.cpu().long()]


# ------------------------------------------------------------------------------------------------ support kernels, set 2
#   All in float64 on the CPU; f16 inputs are taken as the exact values they hold.
_CLAMP_F32 = 88.3762626647949          # MAX_EXP_F32 / -MIN_EXP_F32 (sigmoid.cu:L19-20, gelu.cu:L19-20)
_CLAMP_F16 = (-9.703125, 11.09375)     # MIN_EXP_F16 / MAX_EXP_F16 as rounded to half (sigmoid.cu:L21-22, gelu.cu:L21-22)


def activation(x, op: str, ref_clamp: bool = True):
    """The seven activation families.  relu.cu:L21-24, sigmoid.cu:L27-35, gelu.cu:L38-52 (tanh approximation: the
    default GELU_OPS / HALF_GELU_OPS), swish.cu:L20-22, elu.cu:L41-43 (alpha = 1), hardswish.cu:L36-44,
    hardshrink.cu:L33-39 (lambda = 0.5).  ref_clamp restates the input clamp the sigmoid and gelu kernels apply first
    (per input dtype: the f32 kernels clamp to +-88.376, the f16 kernels to [-9.703, 11.094])."""
    xf = x.detach().cpu().double()
    if ref_clamp and op in ("sigmoid", "gelu"):
        if x.dtype == torch.float16:
            xf = xf.clamp(_CLAMP_F16[0], _CLAMP_F16[1])
        else:
            xf = xf.clamp(-_CLAMP_F32, _CLAMP_F32)
    if op == "relu":
        return xf.clamp_min(0.0)
    if op == "sigmoid":
        return 1.0 / (1.0 + torch.exp(-xf))
    if op == "gelu":
        return 0.5 * xf * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (xf + 0.044715 * xf ** 3)))
    if op == "swish":
        return xf / (1.0 + torch.exp(-xf))
    if op == "elu":
        return torch.where(xf > 0, xf, torch.exp(xf) - 1.0)
    if op == "hardswish":
        return torch.where(xf >= 3.0, xf, torch.where(xf <= -3.0, torch.zeros_like(xf), xf * (xf + 3.0) / 6.0))
    if op == "hardshrink":
        return torch.where((xf > 0.5) | (xf < -0.5), xf, torch.zeros_like(xf))
    raise ValueError(op)


def layer_norm(x, g: float = 1.0, b: float = 0.0, eps: float = 1e-5, eps_inside_k: bool = True):
    """y = (x - mean) * rsqrt(v) * g + b per row (layer_norm.cu:L54-73).  eps_inside_k=True restates the reference's
    v = sum((x-mean)^2) / (K + eps) (L69); False is the textbook sum/K + eps (script oracle layer_norm.py: F.layer_norm)."""
    xf = x.detach().cpu().double()
    K = xf.shape[-1]
    d = xf - xf.mean(-1, keepdim=True)
    ss = d.pow(2).sum(-1, keepdim=True)
    v = ss / (K + eps) if eps_inside_k else ss / K + eps
    return d * torch.rsqrt(v) * g + b


def dot_prod(a, b) -> float:
    """sum(a * b) (dot_product.cu:L35-53), exact products, float64 accumulation."""
    return float((a.detach().cpu().double().flatten() * b.detach().cpu().double().flatten()).sum())


def mat_transpose(x):
    """y[N,M] = x[M,N]^T (mat_transpose.cu:L29-37) - a pure permutation, bit-exact."""
    return x.detach().cpu().t().contiguous()


def gemv(a, x):
    """y = a @ x (sgemv.cu:L32-52, hgemv.cu:L34-52) in float64."""
    return a.detach().cpu().double() @ x.detach().cpu().double().reshape(-1, 1)


def gemm_tf32_bound(a, b):
    """C = A @ B in float64 together with a rigorous elementwise bound on what a TF32 tensor-core product may return:
    the tensor core ignores the low 13 mantissa bits of each fp32 operand (relative error < 2^-10 per operand, so
    < 2^-9 + 2^-20 per product) and accumulates in fp32 (sgemm_wmma_tf32_stage.cu:L25-60 rounds with
    wmma::__float_to_tf32 instead, which is tighter).  Returns (exact, bound)."""
    ad, bd = a.detach().cpu().double(), b.detach().cpu().double()
    exact = ad @ bd
    mag = ad.abs() @ bd.abs()
    return exact, mag * (2.0 ** -9 + 2.0 ** -18)
