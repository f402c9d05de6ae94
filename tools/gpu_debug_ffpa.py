"""Diagnostic for the FFPA kernel: structured inputs, error broken down by row quadrant / 64-column chunk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops

def ref_attn(q, k, v):
    s = (q.float() @ k.float().transpose(-1, -2)) / (q.size(-1) ** 0.5)
    return torch.softmax(s, dim=-1) @ v.float()

def report(tag, o, ref):
    err = (o.float() - ref).abs()[0, 0]
    N, D = err.shape
    print("== %s  max %.4f mean %.5f" % (tag, err.max().item(), err.mean().item()))
    rows = [(r0, min(r0 + 32, N)) for r0 in range(0, min(N, 256), 32)]
    for r0, r1 in rows:
        line = " rows %3d-%3d:" % (r0, r1)
        for c0 in range(0, D, 64):
            line += " %.3f" % err[r0:r1, c0:c0 + 64].max().item()
        print(line)

torch.manual_seed(0)
for variant in (0, 4, 2, 6):
    for (N, D) in ((128, 256), (256, 256), (384, 256), (256, 512)):
        q = torch.randn(1, 1, N, D, dtype=torch.half, device="cuda")
        k = torch.randn(1, 1, N, D, dtype=torch.half, device="cuda")
        v = torch.randn(1, 1, N, D, dtype=torch.half, device="cuda")
        for name, (qq, kk, vv) in {"rand": (q, k, v), "v_ones": (q, k, torch.ones_like(v)), "q_zero": (torch.zeros_like(q), k, v)}.items():
            o = torch.full_like(q, float("nan"))
            ops.ffpa_fwd(qq, kk, vv, o, variant=variant)
            torch.cuda.synchronize()
            report("variant %d N=%d D=%d %s" % (variant, N, D, name), o, ref_attn(qq, kk, vv))
