"""CPU: the C-ABI library loads without a GPU, exports every symbol include/b200k.h declares, and validates
arguments before touching the device.  No compute is attempted here."""
import ctypes

import pytest

from b200k import _loader as L


def test_header_symbols_are_all_exported_and_bound():
    declared = L.declared_symbols()
    assert len(declared) >= 15
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libb200k.so does not export %s" % name
    assert sorted(L._SIGS) == declared, "ctypes signature table and include/b200k.h disagree"


def test_abi_version_and_workspace():
    assert L.lib.b200k_abi_version() == 1
    assert L.lib.b200k_reduce_workspace_bytes() >= 2048 * 4


def test_argument_validation_happens_before_cuda():
    lib = L.lib
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.b200k_hgemm_f16(None, one, one, 8, 8, 8, 0, 0, None) == L.EARG
    assert lib.b200k_hgemm_f16(one, one, one, 8, 8, 7, 0, 0, None) == L.ESHAPE
    assert b"multiples of 8" in lib.b200k_last_error()
    assert lib.b200k_hgemm_f16(one, one, one, 0, 8, 8, 0, 0, None) == L.ESHAPE
    assert lib.b200k_fa2_fwd_f16(None, one, one, one, 1, 1, 8, 64, 0.0, 0, 0, None) == L.EARG
    assert lib.b200k_fa2_fwd_f16(one, one, one, one, 1, 0, 8, 64, 0.0, 0, 0, None) == L.ESHAPE
    assert lib.b200k_ffpa_fwd_f16(one, one, one, one, 1, 1, 8, 200, 0.0, 0, None) == L.EHEADDIM
    assert b"headdim not support" in lib.b200k_last_error()
    assert lib.b200k_elementwise_add(one, one, one, 4, 99, None) in (L.EDTYPE, L.ECUDA, L.EARCH)
    assert lib.b200k_rope_f32(one, one, 4, 7, 1, None) == L.ESHAPE
    assert lib.b200k_softmax(one, one, 4, 8, L.F32, 9, None, None) == L.EARG
    assert lib.b200k_embedding(one, one, one, 4, 4, 4, L.I8, None) == L.EDTYPE
    # generic GEMM entry: fp32 rows must be 16-byte multiples too (4 elements), unknown dtypes are refused
    assert lib.b200k_gemm(one, one, one, 8, 8, 6, 0, L.F32, 0, None) == L.ESHAPE
    assert b"multiples of 4" in lib.b200k_last_error()
    assert lib.b200k_gemm(one, one, one, 8, 8, 4, 0, L.BF16, 0, None) == L.ESHAPE
    assert lib.b200k_gemm(one, one, one, 8, 8, 8, 0, L.I8, 0, None) == L.EDTYPE
    assert lib.b200k_gemm(None, one, one, 8, 8, 8, 1, L.F32, 0, None) == L.EARG
    # second set of support kernels
    assert lib.b200k_layer_norm(one, one, 0, 8, 1.0, 0.0, 1e-5, L.F32, 1, None) == L.ESHAPE
    assert lib.b200k_mat_transpose_f32(one, None, 4, 4, None) == L.EARG
    assert lib.b200k_gemv(one, one, one, 4, 0, L.F32, None) == L.ESHAPE
    assert lib.b200k_dot_prod(one, one, None, 4, L.F32, one, None) == L.EARG
    # round-2 entry points
    assert lib.b200k_fa2_fwd(one, one, one, one, 1, 1, 8, 64, 0.0, 0, L.F32, 0, None, 0, None) == L.EDTYPE
    assert lib.b200k_fa2_fwd(one, one, one, one, 1, 1, 8, 64, 0.0, 1, L.BF16, 0, None, 0, None) == L.EARG   # bf16 + [B,H,D,N] V
    assert lib.b200k_fa2_fwd(one, one, one, one, 1, 1, 8, 48, 0.0, 0, L.F16, 1, None, 0, None) == L.EHEADDIM
    assert lib.b200k_gemm_ex(one, one, one, 12, 8, 8, 1, 0, L.F16, 0, None) == L.ESHAPE     # A^T storage: M % 8
    assert lib.b200k_gemm_ex(one, one, one, 8, 8, 8, 1, 0, L.F32, 0, None) == L.EDTYPE
    assert lib.b200k_transpose_u16_batched(one, None, 1, 4, 4, None) == L.EARG
    assert lib.b200k_transpose_u16_batched(one, one, 0, 4, 4, None) == L.ESHAPE


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    one = ctypes.c_void_p(16)
    rc = L.lib.b200k_hgemm_f16(one, one, one, 128, 128, 128, 0, 0, None)
    assert rc in (L.ECUDA, L.EARCH) and L.last_error()
    with pytest.raises(L.B200KError):
        L.device_info()



def _schedule(num_tiles, num_kb, clusters, tune=0):
    import numpy as np

    lib = L.lib
    n = lib.b200k_debug_hgemm_schedule(num_tiles, num_kb, clusters, tune, None, 0)
    assert n > 0
    rows = np.zeros((n, 7), dtype=np.int32)
    assert lib.b200k_debug_hgemm_schedule(num_tiles, num_kb, clusters, tune, rows.ctypes.data, n) == n
    return rows


@pytest.mark.parametrize("clusters", [74, 66, 8, 3])
def test_stream_k_schedule_covers_every_k_block_once_and_cannot_deadlock(clusters):
    """Host replay of the GEMM kernel's work-item schedule (the plan_stream_k() / get_work() the launcher and the kernel
    share) over many tile counts and K depths: (1) every k-block of every tile is computed exactly once; (2) a tile cut
    along K has exactly one finisher, which holds k-block 0, and its writers are the clusters right after it up to
    last_writer, each with one contiguous range, together the rest of the tile; (3) a writer item is always the first item
    of its cluster and waits for nobody, so the finisher's wait can always be satisfied; (4) stream-K items come before a
    cluster's data-parallel tiles and a cluster has at most two of them; (5) work per cluster differs by at most one
    k-block in the stream-K round."""
    import numpy as np

    for num_kb in (8, 9, 16, 32, 64, 128, 257):
        for num_tiles in sorted({1, 2, clusters - 1, clusters, clusters + 1, 2 * clusters - 1, 2 * clusters + 1,
                                 3 * clusters + clusters // 2, 256, 1024, 1000}):
            if num_tiles < 1:
                continue
            rows = _schedule(num_tiles, num_kb, clusters)
            cover = np.zeros((num_tiles, num_kb), dtype=np.int32)
            for c, i, t, k0, k1, kind, lw in rows:
                assert 0 <= t < num_tiles and 0 <= k0 < k1 <= num_kb
                cover[t, k0:k1] += 1
                assert kind == (1 if k0 > 0 else (2 if k1 < num_kb else 0))
            assert (cover == 1).all(), (num_tiles, num_kb)
            by_cluster = {}
            for r in rows:
                by_cluster.setdefault(int(r[0]), []).append(r)
            sk_load = []
            for c, items in by_cluster.items():
                assert [int(r[1]) for r in items] == list(range(len(items)))
                kinds = [int(r[5]) for r in items]
                n_partial = sum(1 for k in kinds if k != 0)
                assert n_partial <= 2
                for pos, r in enumerate(items):
                    if r[5] == 1:
                        assert pos == 0                      # a writer never sits behind another item
                    if r[5] == 2:
                        assert pos <= 1 and all(int(x[5]) == 1 for x in items[:pos])
            sk = num_tiles > clusters and num_tiles % clusters != 0
            partial_tiles = {int(r[2]) for r in rows if r[5] != 0}
            if not sk:
                assert not partial_tiles
                continue
            rem = num_tiles % clusters
            assert partial_tiles <= set(range(rem))
            for t in partial_tiles:
                parts = sorted((r for r in rows if r[2] == t), key=lambda r: r[3])
                fin = parts[0]
                assert fin[5] == 2 and fin[3] == 0
                writers = parts[1:]
                assert [int(w[0]) for w in writers] == list(range(int(fin[0]) + 1, int(fin[6]) + 1))
                assert all(int(w[5]) == 1 for w in writers)
                assert [int(p[3]) for p in parts[1:]] == [int(p[4]) for p in parts[:-1]] and parts[-1][4] == num_kb
            for c in range(clusters):
                sk_load.append(sum(int(r[4] - r[3]) for r in by_cluster.get(c, []) if r[2] < rem))
            assert max(sk_load) - min(sk_load) <= 1 and sum(sk_load) == rem * num_kb


def test_stream_k_schedule_off_switch_and_short_k():
    """tune bit 20 switches the remainder round off, K shorter than 8 k-blocks never uses it: whole tiles only, strided."""
    for tune, num_kb in ((1 << 20, 64), (0, 4)):
        rows = _schedule(300, num_kb, 74, tune)
        assert (rows[:, 5] == 0).all() and len(rows) == 300
        assert sorted(rows[:, 2].tolist()) == list(range(300))
        assert all(int(r[2]) % 74 == int(r[0]) for r in rows)
