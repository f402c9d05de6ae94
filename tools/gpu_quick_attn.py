"""Quick A/B timing of attention variants: usage gpu_quick_attn.py B H N D kernel(fa2|ffpa) variant [variant...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops
B, H, N, D = [int(x) for x in sys.argv[1:5]]
kern = sys.argv[5]
variants = [int(x, 0) for x in sys.argv[6:]]
torch.manual_seed(1)
q, k, v = [torch.randn(B, H, N, D, dtype=torch.half, device="cuda") for _ in range(3)]
o = torch.empty_like(q)
fn = ops.ffpa_fwd if kern == "ffpa" else ops.fa2_fwd
s_ = (q[:1, :1].float() @ k[:1, :1].float().transpose(-1, -2)) / D ** 0.5
ref = torch.softmax(s_, -1) @ v[:1, :1].float()
fl = 4.0 * B * H * N * N * D
import statistics
res = {v: [] for v in variants}
for rep in range(int(os.environ.get('REPS', '2'))):
    for var in variants:
        o.zero_()
        fn(q, k, v, o, variant=var)
        torch.cuda.synchronize()
        ok = torch.allclose(o[:1, :1].float(), ref, rtol=1e-2, atol=1e-3)
        for _ in range(2): fn(q, k, v, o, variant=var)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): fn(q, k, v, o, variant=var)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 8
        res[var].append(fl / t * 1e-9)
        print("%s D=%d variant 0x%x: %.3f ms %.0f TFLOPS ok=%s" % (kern, D, var, t, fl / t * 1e-9, ok), flush=True)
for var in variants:
    print("MEDIAN %s D=%d variant 0x%x: %.0f TFLOPS (min %.0f max %.0f)" % (kern, D, var, statistics.median(res[var]), min(res[var]), max(res[var])))
