"""HGEMM tuning sweep on the B200 box: GROUP_M x L2-policy (variant high bits), against cuBLAS in the same process.
Output: gpurun_out/hgemm_tune.jsonl"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
import torch
from b200k import ops

def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = open(os.path.join(ROOT, "gpurun_out", "hgemm_tune.jsonl"), "w")
sizes = [int(x) for x in sys.argv[1:]] or [8192, 4096, 16384]
for n in sizes:
    torch.manual_seed(1)
    a = torch.randn(n, n, dtype=torch.half, device="cuda")
    b = torch.randn(n, n, dtype=torch.half, device="cuda")
    c = torch.empty(n, n, dtype=torch.half, device="cuda")
    iters = 30 if n <= 8192 else 6
    fl = 2.0 * n ** 3
    for rep in range(2):
        t = timeit(lambda: torch.matmul(a, b, out=c), iters)
        r = {"n": n, "cfg": "cublas", "tflops": fl / t * 1e-9}
        print(json.dumps(r), flush=True); out.write(json.dumps(r) + "\n")
        for base in (2, 1):
            for gm in (0, 4, 16, 32):
                for pol in (0, 1, 2, 3):
                    if base == 1 and (gm not in (0, 16) or pol not in (0, 1)):
                        continue
                    v = base | (gm << 8) | (pol << 16)
                    t = timeit(lambda: ops.hgemm(a, b, c, variant=v), iters)
                    r = {"n": n, "cfg": "v%d gm%d pol%d" % (base, gm or 8, pol), "tflops": fl / t * 1e-9}
                    print(json.dumps(r), flush=True); out.write(json.dumps(r) + "\n")
out.close()
