"""O^T FFPA kernel (variant 0x200) bring-up: parity against the CPU oracle and the default kernel, rescale path, timing."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
sys.path.insert(0, ROOT)
import torch
from b200k import ops
from oracle import oracle

OUT = os.path.join(ROOT, "gpurun_out", "r2")
os.makedirs(OUT, exist_ok=True)
V3 = 0x200


def emit(rec, f):
    print(json.dumps(rec), flush=True)
    f.write(json.dumps(rec) + "\n")
    f.flush()


def timeit(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    with open(os.path.join(OUT, "ffpa3.jsonl"), "a") as f:
        # structured: V = 1 -> O = 1; one tile
        for shape in [(1, 1, 128, 512), (1, 1, 128, 256)]:
            q, k = [torch.randn(*shape, dtype=torch.half, device="cuda") for _ in range(2)]
            v = torch.ones(*shape, dtype=torch.half, device="cuda")
            o = torch.full(shape, float("nan"), dtype=torch.half, device="cuda")
            ops.ffpa_fwd(q, k, v, o, variant=V3)
            torch.cuda.synchronize()
            emit({"what": "ones", "shape": list(shape), "ok": bool(torch.equal(o, v)), "min": float(o.float().min()), "max": float(o.float().max()),
                  "nan": int(torch.isnan(o).sum())}, f)
        for shape in [(1, 1, 128, 512), (1, 2, 256, 512), (1, 2, 300, 512), (1, 1, 64, 512), (1, 1, 1, 512), (1, 1, 257, 512), (1, 1, 512, 512), (2, 3, 1000, 256), (1, 2, 129, 256),
                      (1, 4, 2048, 512)]:
            torch.manual_seed(shape[2] + shape[3])
            q, k, v = [torch.randn(*shape, dtype=torch.half, device="cuda") for _ in range(3)]
            o = torch.full(shape, float("nan"), dtype=torch.half, device="cuda")
            ops.ffpa_fwd(q, k, v, o, variant=V3)
            o2 = torch.empty_like(o)
            ops.ffpa_fwd(q, k, v, o2)
            torch.cuda.synchronize()
            want = oracle.attention(q, k, v).float()
            err = float((o.cpu().float() - want).abs().max())
            emit({"what": "parity", "shape": list(shape), "ok": bool(torch.allclose(o.cpu().float(), want, rtol=1e-2, atol=1e-3)), "max_err": err,
                  "max_diff_vs_default": float((o.float() - o2.float()).abs().max()), "nan": int(torch.isnan(o).sum())}, f)
        # the lazy-rescale path: scores grow along the key axis by more than 2^8 several times
        torch.manual_seed(9)
        B, H, N, D = 1, 2, 1024, 512
        q = torch.randn(B, H, N, D, dtype=torch.half, device="cuda")
        k = torch.randn(B, H, N, D, dtype=torch.half, device="cuda") * 0.1
        ramp = torch.linspace(0, 60, N, device="cuda").view(1, 1, N, 1)
        k = (k.float() + ramp * q[:, :, :1].float() / (q[:, :, :1].float().pow(2).sum(-1, keepdim=True) ** 0.5) * 0.3).half()
        v = torch.randn(B, H, N, D, dtype=torch.half, device="cuda")
        o = torch.full((B, H, N, D), float("nan"), dtype=torch.half, device="cuda")
        ops.ffpa_fwd(q, k, v, o, variant=V3)
        torch.cuda.synchronize()
        want = oracle.attention(q, k, v).float()
        emit({"what": "rescale", "ok": bool(torch.allclose(o.cpu().float(), want, rtol=1e-2, atol=1e-3)), "max_err": float((o.cpu().float() - want).abs().max()),
              "nan": int(torch.isnan(o).sum())}, f)
        # timing at config #4 and a D = 256 case
        for shape in [(1, 32, 4096, 512), (1, 32, 8192, 256)]:
            Bq, Hq, Nq, Dq = shape
            q, k, v = [torch.randn(*shape, dtype=torch.half, device="cuda") for _ in range(3)]
            o = torch.empty_like(q)
            fl = 4.0 * Bq * Hq * Nq * Nq * Dq
            res = {"default": [], "otrans": []}
            fns = {"default": lambda: ops.ffpa_fwd(q, k, v, o), "otrans": lambda: ops.ffpa_fwd(q, k, v, o, variant=V3)}
            for n in fns:
                timeit(fns[n], 2)
            for r in range(5):
                for n in fns:
                    res[n].append(fl / timeit(fns[n], 5) * 1e-9)
            for n, vals in res.items():
                emit({"what": "time", "shape": list(shape), "variant": n, "median": sorted(vals)[2], "all": [round(x) for x in vals]}, f)


if __name__ == "__main__":
    main()
