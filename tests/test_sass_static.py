"""Static checks on the machine code of the shipped library (cuobjdump -sass; no GPU needed).  They pin properties that
the measurements of DESIGN.md rest on and that a careless edit loses silently:
* the tensor-core paths are tcgen05 (UTCHMMA) fed by TMA (UTMALDG), with tensor-memory loads / stores, and there is no
  legacy mma.sync (HMMA) anywhere in the library;
* DESIGN 3.0's hand-shake rule: a `cta_group::2` kernel contains a GPU-scope memory fence (MEMBAR.ALL.GPU, what
  mbarrier.arrive.release.cluster and barrier.cluster.arrive.release compile to) only at its two cluster barriers - plus,
  in the FFPA O^T kernel, the rarely taken rescale acknowledgement and the epilogue's cluster barrier.  A release.cluster
  arrive on a per-tile path cost 20 % of that kernel and 11 % of the 2048^3 GEMM."""
import collections
import functools
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cuda-learn-notes_b200", "b200k", "libb200k.so")
KEYS = ("MEMBAR.ALL.GPU", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "STAS")


@functools.lru_cache(maxsize=1)
def _per_kernel():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if not os.path.exists(LIB):
        pytest.skip("libb200k.so not built")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, legacy, cur = collections.defaultdict(collections.Counter), 0, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur]["_"] += 0
            continue
        if cur is None:
            continue
        if re.search(r"\bHMMA\.", line):
            legacy += 1
        for k in KEYS:
            if k in line:
                counts[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True, check=True).stdout.splitlines()
    return {n: counts[m] for n, m in zip(names, counts)}, legacy


def _kernels(prefix):
    per, _ = _per_kernel()
    ks = {n: c for n, c in per.items() if ("b200k::" + prefix) in n}
    assert ks, prefix
    return ks


def test_tensor_core_paths_are_tcgen05_and_tma_no_legacy_mma():
    per, legacy = _per_kernel()
    assert legacy == 0
    for prefix in ("hgemm_tcgen05_kernel", "fa2_fwd_tcgen05_kernel", "ffpa_fwd_tcgen05_kernel", "ffpa2_fwd_tcgen05_kernel",
                   "ffpa3_fwd_tcgen05_kernel"):
        for name, c in _kernels(prefix).items():
            assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, name
    for name, c in _kernels("fa2_fwd_tcgen05_kernel").items():
        assert c["STTM"] > 0 and c["UTMASTG"] > 0, name      # P goes back to tensor memory, O leaves by TMA store


def test_no_gpu_scope_fence_on_per_tile_paths_of_the_cluster_kernels():
    pair_gemms = {n: c for n, c in _kernels("hgemm_tcgen05_kernel").items() if "GemmCfg<2," in n}
    assert len(pair_gemms) >= 12
    for name, c in pair_gemms.items():
        assert c["MEMBAR.ALL.GPU"] == 2, (name, c["MEMBAR.ALL.GPU"])          # cluster barrier at start and at end
    for name, c in _kernels("ffpa2_fwd_tcgen05_kernel").items():
        assert c["MEMBAR.ALL.GPU"] == 2, (name, c["MEMBAR.ALL.GPU"])
    for name, c in _kernels("ffpa3_fwd_tcgen05_kernel").items():
        # start + end barriers, the epilogue's barrier (arrive + wait sides are separate code paths for the two warp groups),
        # and the rescale acknowledgement (two inlined copies of apply_decision)
        assert c["MEMBAR.ALL.GPU"] <= 6, (name, c["MEMBAR.ALL.GPU"])
        assert c["STAS"] >= 4, name                                            # factors, flags and verdict travel by st.async
    # the single-CTA kernels have no cluster at all
    for prefix in ("fa2_fwd_tcgen05_kernel", "ffpa_fwd_tcgen05_kernel"):
        for name, c in _kernels(prefix).items():
            assert c["MEMBAR.ALL.GPU"] == 0, name
