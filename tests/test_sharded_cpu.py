"""CPU, world_size 2 (gloo): host logic of the batch-sharded attention path — partitioning, the single broadcast,
zero-copy shard views, output gather.  The compute callback is injected (the oracle) because there is no GPU here;
the product default is the CUDA kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_batch():
    from b200k.sharded import shard_bounds

    for B in (1, 2, 7, 32):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q, mode="legacy"):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "cuda-learn-notes_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200k import sharded
    from oracle import oracle

    H, N, D = 2, 64, 32
    torch.manual_seed(123)
    full = torch.randn(3, B, H, N, D).half()  # identical on every rank (same seed) = what rank 0 would broadcast
    buf = sharded.broadcast_qkv(full if rank == 0 else None, (B, H, N, D), torch.device("cpu"))
    assert torch.equal(buf, full)
    calls = []

    def attn(qq, kk, vv, oo):
        assert qq.is_contiguous() and qq.data_ptr() == buf[0, lo_hi[0]:].data_ptr()  # zero-copy view of the shard
        calls.append(qq.shape[0])
        oo.copy_(oracle.attention(qq, kk, vv))

    lo_hi = sharded.shard_bounds(B, world, rank)
    ref = oracle.attention(full[0], full[1], full[2])
    if mode == "legacy":
        o, span = sharded.sharded_attention_fwd(buf, attn_fn=attn)
        assert span == lo_hi and o.shape[0] == lo_hi[1] - lo_hi[0]
        gathered = sharded.gather_output(o, B)
        ok = torch.equal(gathered, ref) and (calls == [o.shape[0]] or o.shape[0] == 0)
    else:
        # the three input distributions: only rank 0 passes data; every rank must end up with exactly its slice
        def attn2(qq, kk, vv, oo):
            calls.append(qq.shape[0])
            oo.copy_(oracle.attention(qq, kk, vv))

        o, sh = sharded.sharded_attention(full.clone() if rank == 0 else None, (B, H, N, D), torch.device("cpu"),
                                          mode=mode, attn_fn=attn2, chunk_batches=1)
        lo, hi = lo_hi
        ok = sh.span == lo_hi and torch.equal(sh.q, full[0, lo:hi]) and torch.equal(sh.k, full[1, lo:hi]) \
            and torch.equal(sh.v, full[2, lo:hi])
        if mode == "pipelined":
            ok = ok and calls == [1] * (hi - lo)           # one launch per one-batch chunk, in order
        else:
            ok = ok and calls == ([hi - lo] if hi > lo else [])
        if mode != "broadcast" and rank != 0:
            ok = ok and sh.keep.shape[1] == hi - lo        # a receiver holds 1/G of the bytes, not the whole buffer
        ok = ok and torch.equal(sharded.gather_output(o, B), ref)
        ok = ok and sharded.shards_equal_to(o, ref if rank == 0 else None, B)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,mode", [(4, "legacy"), (3, "legacy"), (4, "broadcast"), (5, "scatter"), (5, "pipelined"),
                                    (1, "pipelined")])
def test_sharded_attention_world2_gloo(B, mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
