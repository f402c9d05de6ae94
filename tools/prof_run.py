"""Tiny driver for ncu captures: runs one op a few times.  usage: prof_run.py hgemm M N K [variant] | fa2 B H N D | ffpa B H N D"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch  # noqa: E402
from b200k import ops  # noqa: E402

op = sys.argv[1]
torch.manual_seed(1)
if op == "hgemm":
    M, N, K = [int(x) for x in sys.argv[2:5]]
    variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    a = torch.randn(M, K, dtype=torch.half, device="cuda")
    b = torch.randn(K, N, dtype=torch.half, device="cuda")
    c = torch.empty(M, N, dtype=torch.half, device="cuda")
    for _ in range(4):
        ops.hgemm(a, b, c, variant=variant)
else:
    B, H, N, D = [int(x) for x in sys.argv[2:6]]
    q, k, v = [torch.randn(B, H, N, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    fn = ops.ffpa_fwd if op == "ffpa" else ops.fa2_fwd
    for _ in range(4):
        fn(q, k, v, o)
torch.cuda.synchronize()
print("done")
