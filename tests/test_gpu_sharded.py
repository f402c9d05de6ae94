"""GPU, multi-process (NCCL): the batch-sharded attention path of BASELINE config #5 against the
single-GPU kernel on the same inputs.  Every rank runs the tcgen05 kernel on its batch slice; the shards are sent to
rank 0 and compared BIT FOR BIT with rank 0's own one-call result (same kernel, same rows => identical bits), for the
three input distributions (one broadcast / scatter / chunked pipelined scatter).  Needs >= 2 GPUs (gpurun --gpus 2);
skipped on a one-GPU box.  The last test runs config #5 WHOLE (B=32: 2^31 elements per tensor, 16 GiB) as one call on
one GPU and checks that batches on either side of the 2^31-byte / 2^32-byte offsets equal their per-batch results."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "cuda-learn-notes_b200")]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from b200k import ops, sharded

    B, H, N, D = shape
    verdict = {}
    try:
        qkv = None
        o_full = None
        if rank == 0:
            torch.manual_seed(77)
            qkv = torch.randn(3, B, H, N, D, dtype=torch.half, device=dev)
            o_full = torch.full((B, H, N, D), float("nan"), dtype=torch.half, device=dev)
            (ops.fa2_fwd if D <= 128 else ops.ffpa_fwd)(qkv[0], qkv[1], qkv[2], o_full)   # the 1-GPU run
            assert torch.isfinite(o_full).all()
        for mode in sharded.MODES:
            o, sh = sharded.sharded_attention(qkv, shape, dev, mode=mode, chunk_batches=1)
            torch.cuda.synchronize()
            lo, hi = sharded.shard_bounds(B, world, rank)
            assert sh.span == (lo, hi) and o.shape[0] == hi - lo
            if mode != "broadcast" and rank != 0:
                assert sh.keep.shape[1] == hi - lo          # received only its own slice
            verdict[mode] = sharded.shards_equal_to(o, o_full, B)
            g = sharded.gather_output(o, B)
            if rank == 0:
                verdict[mode] = verdict[mode] and bool(torch.equal(g, o_full))
        q.put((rank, verdict))
    except Exception as e:  # noqa
        q.put((rank, {"error": repr(e)[:300]}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(4, 8, 1024, 128), (5, 4, 777, 64), (3, 2, 512, 256), (1, 4, 512, 128)])
def test_sharded_equals_single_gpu_bit_for_bit(shape):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r] == {"broadcast": True, "scatter": True, "pipelined": True}, (r, res[r])


def test_config5_whole_in_one_call_64bit_addressing():
    """(32,64,8192,128): 2^31 elements = 4 GiB per tensor; the reference indexes with 32-bit ints."""
    from b200k import ops

    free, _ = torch.cuda.mem_get_info()
    if free < 24 * 2 ** 30:
        pytest.skip("needs 24 GiB of free device memory")
    B, H, N, D = 32, 64, 8192, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.empty(3, B, H, N, D, dtype=torch.half, device="cuda")
    for t in range(3):
        for b in range(B):
            qkv[t, b].normal_(generator=g)
    o = torch.full((B, H, N, D), float("nan"), dtype=torch.half, device="cuda")
    ops.fa2_fwd(qkv[0], qkv[1], qkv[2], o)
    torch.cuda.synchronize()
    for b in (0, 7, 8, 15, 16, 31):            # byte offsets 0, ~2^30*.875, 2^30, ..., 2^31, ~2^32
        ob = torch.empty(1, H, N, D, dtype=torch.half, device="cuda")
        ops.fa2_fwd(qkv[0, b:b + 1], qkv[1, b:b + 1], qkv[2, b:b + 1], ob)
        assert torch.equal(ob[0], o[b]), b
    assert torch.isfinite(o[::5]).all()
