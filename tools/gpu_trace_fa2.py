"""Cycle-level trace of one FA-2 CTA (pipeline analysis) + A/B timing of kernel variants.  Run on the B200 box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops, _loader as L

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = {}
for (B, H, N, D) in ((4, 48, 8192, 64), (4, 64, 8192, 128)):
    torch.manual_seed(1)
    q, k, v = [torch.randn(B, H, N, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.empty_like(q)
    fl = 4.0 * B * H * N * N * D
    for rep in range(3):
        if D == 64:
            variants = ((0, "default"), (0x1C000, "no poly"))
        else:
            variants = ((0, "default (shared S)"), (0x400, "P aliases S"), (0x1C000, "shared S, no poly"))
        if rep == 0 and D == 64:
            s_ = (q[:1, :2].float() @ k[:1, :2].float().transpose(-1, -2)) / D ** 0.5
            ref = torch.softmax(s_, -1) @ v[:1, :2].float()
            for variant, name in variants:
                o.zero_()
                ops.fa2_fwd(q, k, v, o, variant=variant)
                err = (o[:1, :2].float() - ref).abs()
                print("check %s: max %.2e mean %.2e allclose %s" % (name, err.max().item(), err.mean().item(),
                      torch.allclose(o[:1, :2].float(), ref, rtol=1e-2, atol=1e-3)), flush=True)
        for variant, name in variants:
            t = timeit(lambda: ops.fa2_fwd(q, k, v, o, variant=variant))
            print("D=%d %s: %.3f ms %.0f TFLOPS" % (D, name, t, fl / t * 1e-9), flush=True)
    for variant, name in ((0x100, "default"),) + (((0x500, "P aliases S"),) if D == 128 else ()):
        tr = torch.zeros(3 * 32 * 8, dtype=torch.int64, device="cuda")
        L.check(L.lib.b200k_debug_set_trace(tr.data_ptr()))
        ops.fa2_fwd(q, k, v, o, variant=variant)
        torch.cuda.synchronize()
        t = tr.cpu().view(3, 32, 8)
        t0 = int(t[t > 0].min())
        rel = (t - t0).clamp(min=-1)
        out["D%d_%s" % (D, name)] = rel.tolist()
        print("== trace D=%d %s (cycles since first event; rows j=4..9)" % (D, name))
        for role, rn in enumerate(("MMA (D<=96): kfull sfree0 S0 sfree1 S1 pfull0[0] pfull1[0] end | (D=128 shared S): - PV0(j) S0(j+1) PV1(j-1) S1(j) pfull0[0] pfull1[0]", "WG0: sfull ld max/turn st0 arrive0 - end", "WG1")):
            print(rn)
            for j in range(6, 10):
                print("   j=%d " % j + " ".join("%7d" % x for x in rel[role, j].tolist()))
    L.lib.b200k_debug_set_trace(None)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fa2_trace.json"), "w"))
