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
# round-1 additions: bf16 / TF32 GEMM (incl. the fp32 MN-major operand and a ragged N), D = 512 pair kernel, second support set
for dt in (torch.bfloat16, torch.float32):
    A = torch.randn(300, 264, device="cuda").to(dt); Bm = torch.randn(264, 392, device="cuda").to(dt)
    Cc = torch.empty(300, 392, device="cuda").to(dt)
    ops.gemm(A, Bm, Cc)
    ops.gemm(A, Bm.t().contiguous().t(), Cc, tn=True)
q, k, v = [torch.randn(1, 1, 300, 512, dtype=torch.half, device="cuda") for _ in range(3)]
o = torch.empty_like(q)
ops.ffpa_fwd(q, k, v, o)
xh = torch.randn(33, 1000, dtype=torch.half, device="cuda"); yh = torch.empty_like(xh)
for op in ("relu", "sigmoid", "gelu", "swish", "elu", "hardswish", "hardshrink"):
    ops.activation(x, y, op); ops.activation(xh.flatten()[1:], yh.flatten()[1:], op)
ops.layer_norm(x, y, 1.0, 0.0); ops.layer_norm(xh, yh, 1.0, 0.0); ops.dot_prod(x, y)
t = torch.empty(1000, 77, device="cuda"); ops.mat_transpose(x, t)
gv = torch.randn(1000, 1, device="cuda"); gy = torch.empty(77, 1, device="cuda"); ops.gemv(x, gv, gy)
ops.rope_f32(x, y, True); ops.rope_f32(x, y, False)
torch.cuda.synchronize()
print("sanitize run done")
# round-2 additions: 512x256 pair tile, stream-K remainder round (forced on a small problem), A^T storage, the masked and
# bf16 attention builds, CTA-pair FFPA at D = 768 / 1024
a2 = torch.randn(1024, 520, dtype=torch.half, device="cuda"); b2 = torch.randn(520, 768, dtype=torch.half, device="cuda")
c2 = torch.empty(1024, 768, dtype=torch.half, device="cuda")
ops.hgemm(a2, b2, c2, variant=4)
ops.hgemm(a2, b2, c2, variant=2 | (1 << 22))      # stream-K forced (12 tiles, 74 clusters)
a3 = torch.randn(5120, 512, dtype=torch.half, device="cuda"); b3 = torch.randn(512, 4096, dtype=torch.half, device="cuda")
c3 = torch.empty(5120, 4096, dtype=torch.half, device="cuda")
ops.hgemm(a3, b3, c3, variant=2)                   # 320 tiles: 24 of them stream-K
ops.gemm(a2.t().contiguous().t(), b2, c2, a_km=True)
for D in (64, 128):
    q, k, v = [torch.randn(2, 2, 333, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    sl = torch.tensor([333, 100], dtype=torch.int32, device="cuda")
    ops.fa2_fwd(q, k, v, o, causal=True, seqlens_k=sl)
    qb, kb, vb = q.bfloat16(), k.bfloat16(), v.bfloat16()
    ob = torch.empty_like(qb)
    ops.fa2_fwd(qb, kb, vb, ob, causal=True)
for D in (768, 1024):
    q, k, v = [torch.randn(1, 1, 300, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    ops.ffpa_fwd(q, k, v, o)
torch.cuda.synchronize()
print("sanitize run (round 2 additions) done")
# final round-2 additions: 512x256 tile with fewer k-blocks than ring stages, batched 16-bit transpose
for K in (64, 192, 320):
    a4 = torch.randn(1024, K, dtype=torch.half, device="cuda"); b4 = torch.randn(K, 520, dtype=torch.half, device="cuda")
    c4 = torch.empty(1024, 520, dtype=torch.half, device="cuda")
    ops.hgemm(a4, b4, c4, variant=4)
xt = torch.randn(2, 3, 77, 130, dtype=torch.half, device="cuda"); yt = torch.empty(2, 3, 130, 77, dtype=torch.half, device="cuda")
ops.transpose_16bit_batched(xt, yt)
torch.cuda.synchronize()
print("sanitize run (final additions) done")
# the O^T FFPA kernel (default at D = 512, variant 0x200 at D = 256), incl. a ragged two-tile case
for D, N in ((512, 300), (256, 257), (512, 64)):
    q, k, v = [torch.randn(1, 2, N, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    ops.ffpa_fwd(q, k, v, o, variant=0x200)
torch.cuda.synchronize()
print("sanitize run (O^T FFPA) done")
