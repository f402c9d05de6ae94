"""Batch-sharded attention forward across the GPUs of one box (BASELINE config #5).

Attention is embarrassingly parallel over batch x heads: the reference's grid is literally (N/Br, B*H)
(kernels/flash-attn/mma/basic/flash_attn_mma_share_qkv.cu:L777-778) and it has no multi-GPU code at all.  Here:
one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), rank r owns batches [r*B/G, (r+1)*B/G) — a
contiguous slice of the outermost dim, so a zero-copy view — and runs the single-GPU tcgen05 kernel on it.  The only
collective is the input distribution named by the north star: ONE broadcast of the packed Q|K|V buffer from rank 0.
Outputs stay sharded (all_gather only on request, for parity checks).  No cross-GPU reduction exists in this path.

The compute callback is injectable so the host logic (partitioning, collective, layout) is testable with gloo on CPU;
the default is the CUDA kernel and there is no CPU fallback in the product path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous batch range of `rank`; B need not divide evenly (first B % world ranks get one more)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _default_attn(q, k, v, o):
    from . import ops

    if q.size(-1) <= 128:
        ops.fa2_fwd(q, k, v, o)
    else:
        ops.ffpa_fwd(q, k, v, o)


def broadcast_qkv(qkv: Optional[torch.Tensor], shape, device, src: int = 0, group=None) -> torch.Tensor:
    """One broadcast of the packed [3, B, H, N, D] fp16 buffer from `src` to every rank."""
    B, H, N, D = shape
    if dist.get_rank(group) == src:
        assert qkv is not None and tuple(qkv.shape) == (3, B, H, N, D)
        buf = qkv.contiguous()
    else:
        buf = torch.empty(3, B, H, N, D, dtype=torch.float16, device=device)
    dist.broadcast(buf, src=src, group=group)
    return buf


def sharded_attention_fwd(qkv_full: torch.Tensor, attn_fn: Callable = _default_attn, group=None,
                          out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """Run attention on this rank's batch slice of a replicated packed buffer [3,B,H,N,D].
    Returns (o_shard [b_local,H,N,D], (lo, hi))."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = qkv_full.size(1)
    lo, hi = shard_bounds(B, world, rank)
    q, k, v = qkv_full[0, lo:hi], qkv_full[1, lo:hi], qkv_full[2, lo:hi]  # contiguous views (outermost-dim slices)
    if out is None:
        out = torch.empty_like(q)
    if hi > lo:
        attn_fn(q, k, v, out)
    return out, (lo, hi)


def gather_output(o_shard: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """all_gather of the sharded outputs — for parity checks only; not part of the timed path."""
    world = dist.get_world_size(group)
    _, H, N, D = o_shard.shape
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, H, N, D, dtype=o_shard.dtype, device=o_shard.device)
    pad[: o_shard.size(0)] = o_shard
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
