"""Driver for ncu captures of the support kernels: a few launches each at HBM-bound sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops
torch.manual_seed(0)
n = 64 * 1024 * 1024
a = torch.randn(n, device="cuda"); b = torch.randn(n, device="cuda"); c = torch.empty_like(a)
x = torch.randn(16384, 8192, dtype=torch.half, device="cuda"); y = torch.empty_like(x)
xf = torch.randn(16384, 4096, device="cuda"); yf = torch.empty_like(xf)
w = torch.randn(32768, 1024, dtype=torch.half, device="cuda"); idx = torch.randint(0, 32768, (65536,), dtype=torch.int32, device="cuda")
out = torch.empty(65536, 1024, dtype=torch.half, device="cuda")
hist_in = torch.randint(0, 256, (64 * 1024 * 1024,), dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.elementwise_add(a, b, c)
    ops.block_all_reduce_sum(a)
    ops.softmax(x, y, ops.SOFTMAX_SAFE)
    ops.rms_norm(x, y, 1.0)
    ops.softmax(xf, yf, ops.SOFTMAX_SAFE)
    ops.rope_f32(xf, yf, True)
    ops.embedding(idx, w, out)
    ops.histogram_i32(hist_in, nbins=256)
torch.cuda.synchronize()
print("done")
