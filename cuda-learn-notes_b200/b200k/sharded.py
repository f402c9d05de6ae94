"""Batch-sharded attention forward across the GPUs of one box (BASELINE config #5).

Attention is embarrassingly parallel over batch x heads: the reference's grid is literally (N/Br, B*H)
(kernels/flash-attn/mma/basic/flash_attn_mma_share_qkv.cu:L777-778) and it has no multi-GPU code at all.  Here:
one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), rank r owns batches [r*B/G, (r+1)*B/G) — a
contiguous slice of the outermost dim — and runs the single-GPU tcgen05 kernel on it.  Outputs stay sharded
(gather only on request, for parity checks).  No cross-GPU reduction exists in this path.

The only communication is the input distribution from the rank that holds Q, K, V.  Three ways, same result:

``broadcast``  ONE broadcast of the packed [3,B,H,N,D] buffer (the north star's wording; every rank receives G times
               the bytes it needs, the shard is a zero-copy view of the replicated buffer).
``scatter``    rank r receives only its own [3,b_r,H,N,D] slices (1/G of the bytes per receiver; the source still
               sends (G-1)/G of the buffer through its own NVLink ports, which is the floor for any scheme whose
               inputs start on one GPU).
``pipelined``  the scatter cut into per-batch chunks, posted chunk-major on NCCL's own stream before any compute is
               launched: chunk c of every rank leaves the source before chunk c+1 of any rank, and each rank launches
               the attention kernel of chunk c as soon as that chunk has landed, so all of the compute except the
               last chunk's runs under the transfer.

The compute callback is injectable so the host logic (partitioning, collective, chunk order) is testable with gloo
on CPU; the default is the CUDA kernel and there is no CPU fallback in the product path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

MODES = ("broadcast", "scatter", "pipelined")


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


@dataclass
class Shard:
    """What one rank holds after the input distribution: its batch slice of Q, K, V ([b,H,N,D] each) and, per chunk
    of that slice, the outstanding communication handles that must complete before the chunk may be read."""
    q: torch.Tensor
    k: torch.Tensor
    v: torch.Tensor
    span: Tuple[int, int]
    chunks: List[Tuple[int, int, list]] = field(default_factory=list)  # (lo, hi, works) relative to the slice
    keep: Optional[torch.Tensor] = None   # keeps the backing buffer alive (views)
    src_works: list = field(default_factory=list)  # on the source: the send handles (wait before reusing qkv)

    def wait_all(self):
        for _, _, works in self.chunks:
            for w in works:
                w.wait()
        for w in self.src_works:
            w.wait()


def _send_recv_chunked(qkv, shape, device, chunk_batches, src, group) -> Shard:
    """Chunk-major scatter with point-to-point ops: for c = 0, 1, ...: the source sends batches
    [lo_r + c*cb, lo_r + (c+1)*cb) of Q, K and V to every other rank r, as ONE grouped NCCL call per chunk index."""
    B, H, N, D = shape
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = [shard_bounds(B, world, r) for r in range(world)]
    lo, hi = spans[rank]
    nb = hi - lo
    cb = max(1, int(chunk_batches)) if chunk_batches else max(1, max(h - l for l, h in spans))
    nchunks = max((h - l + cb - 1) // cb for l, h in spans)
    if rank == src:
        assert qkv is not None and tuple(qkv.shape) == (3, B, H, N, D) and qkv.is_contiguous()
        sh = Shard(qkv[0, lo:hi], qkv[1, lo:hi], qkv[2, lo:hi], (lo, hi), keep=qkv)
    else:
        buf = torch.empty(3, nb, H, N, D, dtype=torch.float16, device=device)
        sh = Shard(buf[0], buf[1], buf[2], (lo, hi), keep=buf)
    for c in range(nchunks):
        ops = []
        if rank == src:
            for r in range(world):
                if r == src:
                    continue
                rl, rh = spans[r]
                a, b = rl + c * cb, min(rl + (c + 1) * cb, rh)
                if a < b:
                    for t in range(3):
                        ops.append(dist.P2POp(dist.isend, qkv[t, a:b], r, group))
        else:
            a, b = c * cb, min((c + 1) * cb, nb)
            if a < b:
                for t in range(3):
                    ops.append(dist.P2POp(dist.irecv, sh.keep[t, a:b], src, group))
        works = dist.batch_isend_irecv(ops) if ops else []
        a, b = c * cb, min((c + 1) * cb, nb)
        if rank == src:
            sh.src_works += works
            if a < b:
                sh.chunks.append((a, b, []))        # the source's own slice is already in place
        elif a < b:
            sh.chunks.append((a, b, works))
    return sh


def distribute_qkv(qkv: Optional[torch.Tensor], shape, device, mode: str = "broadcast", src: int = 0, group=None,
                   chunk_batches: Optional[int] = 1) -> Shard:
    """Move the inputs from `src` (which passes the packed [3,B,H,N,D] fp16 tensor; others pass None) to the ranks.
    Returns immediately with communication possibly still in flight (see :class:`Shard`)."""
    if mode not in MODES:
        raise ValueError("mode must be one of %s" % (MODES,))
    B, H, N, D = shape
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if mode == "broadcast":
        buf = broadcast_qkv(qkv, shape, device, src, group)
        lo, hi = shard_bounds(B, world, rank)
        return Shard(buf[0, lo:hi], buf[1, lo:hi], buf[2, lo:hi], (lo, hi), [(0, hi - lo, [])] if hi > lo else [], keep=buf)
    return _send_recv_chunked(qkv, shape, device, None if mode == "scatter" else chunk_batches, src, group)


def attention_on_shard(sh: Shard, attn_fn: Callable = _default_attn, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run attention chunk by chunk on this rank's slice; each chunk waits (on the current stream) only for its own
    transfer.  Returns o_shard [b_local,H,N,D]."""
    if out is None:
        out = torch.empty_like(sh.q)
    for a, b, works in sh.chunks:
        for w in works:
            w.wait()          # NCCL: the current stream waits for the transfer; the host does not block
        attn_fn(sh.q[a:b], sh.k[a:b], sh.v[a:b], out[a:b])
    return out


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


def sharded_attention(qkv: Optional[torch.Tensor], shape, device, mode: str = "broadcast", src: int = 0, group=None,
                      attn_fn: Callable = _default_attn, out: Optional[torch.Tensor] = None,
                      chunk_batches: Optional[int] = 1) -> Tuple[torch.Tensor, Shard]:
    """Distribution + compute in one call: what a user of config #5 runs.  Returns (o_shard, shard)."""
    sh = distribute_qkv(qkv, shape, device, mode, src, group, chunk_batches)
    o = attention_on_shard(sh, attn_fn, out)
    for w in sh.src_works:
        w.wait()
    return o, sh


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


def shards_equal_to(o_shard: torch.Tensor, o_full_on_dst: Optional[torch.Tensor], B: int, dst: int = 0, group=None) -> bool:
    """Parity check without replicating the output: every rank sends its shard to `dst`, which compares it bit for bit
    with rows [lo,hi) of its own single-GPU result.  Returns the verdict on every rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ok = torch.ones(1, dtype=torch.int32, device=o_shard.device)
    if rank == dst:
        lo, hi = shard_bounds(B, world, dst)
        good = bool(torch.equal(o_shard, o_full_on_dst[lo:hi]))
        for r in range(world):
            if r == dst:
                continue
            lo, hi = shard_bounds(B, world, r)
            if hi > lo:
                tmp = torch.empty_like(o_full_on_dst[lo:hi])
                dist.recv(tmp, src=r, group=group)
                good = good and bool(torch.equal(tmp, o_full_on_dst[lo:hi]))
                del tmp
        ok.fill_(1 if good else 0)
    elif o_shard.size(0) > 0:
        dist.send(o_shard.contiguous(), dst=dst, group=group)
    dist.broadcast(ok, src=dst, group=group)
    return bool(ok.item())
