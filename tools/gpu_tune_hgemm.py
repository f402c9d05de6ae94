"""HGEMM A/B tuning on the B200 box: configurations are timed round-robin (the GPU's clock / power state drifts, so
only interleaved comparisons mean anything).  Output: gpurun_out/hgemm_tune.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops

def timeit(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(1)
a = torch.randn(n, n, dtype=torch.half, device="cuda")
b = torch.randn(n, n, dtype=torch.half, device="cuda")
c = torch.empty(n, n, dtype=torch.half, device="cuda")
cfgs = {"cublas": None}
for gm in (8, 16, 4):
    for pol in (0, 1, 2):
        cfgs["gm%d_pol%d" % (gm, pol)] = 2 | (gm << 8) | (pol << 16)
res = {k: [] for k in cfgs}
fl = 2.0 * n ** 3
for k, v in cfgs.items():  # warm-up
    fn = (lambda: torch.matmul(a, b, out=c)) if v is None else (lambda v=v: ops.hgemm(a, b, c, variant=v))
    timeit(fn, 3)
for r in range(rounds):
    for k, v in cfgs.items():
        fn = (lambda: torch.matmul(a, b, out=c)) if v is None else (lambda v=v: ops.hgemm(a, b, c, variant=v))
        res[k].append(fl / timeit(fn, 20) * 1e-9)
with open(os.path.join(ROOT, "gpurun_out", "hgemm_tune.jsonl"), "w") as f:
    for k, v in res.items():
        r = {"n": n, "cfg": k, "mean": sum(v) / len(v), "min": min(v), "max": max(v), "all": [round(x) for x in v]}
        print(json.dumps(r), flush=True)
        f.write(json.dumps(r) + "\n")
