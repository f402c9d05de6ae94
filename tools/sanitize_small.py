"""Small-shape run of every tensor-core kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops
torch.manual_seed(0)
a = torch.randn(300, 264, dtype=torch.half, device="cuda"); b = torch.randn(264, 520, dtype=torch.half, device="cuda")
c = torch.empty(300, 520, dtype=torch.half, device="cuda")
for v in (1, 2, 3):
    ops.hgemm(a, b, c, variant=v)
    ops.hgemm(a, b.t().contiguous().t(), c, tn=True, variant=v)
for D, fn in ((64, ops.fa2_fwd), (128, ops.fa2_fwd), (96, ops.fa2_fwd), (256, ops.ffpa_fwd), (576, ops.ffpa_fwd)):
    q, k, v = [torch.randn(1, 2, 333, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    fn(q, k, v, o)
x = torch.randn(77, 1000, device="cuda"); y = torch.empty_like(x)
ops.softmax(x, y, 2); ops.rms_norm(x, y, 1.0); ops.block_all_reduce_sum(x)
torch.cuda.synchronize()
print("sanitize run done")
