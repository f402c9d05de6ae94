"""GPU bring-up check for the tcgen05 attention kernels (run on the B200 box through gpurun).

Each group of cases runs in its own subprocess with a timeout so a trap/hang does not take the others down.
Results: gpurun_out/attn_check.jsonl.   Usage: python tools/gpu_check_attn.py [--perf] [--ffpa]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
OUT = os.path.join(ROOT, "gpurun_out")

GROUPS = {
    "d64": [(1, 2, 256, 64), (1, 2, 128, 64), (2, 3, 1000, 64), (1, 1, 77, 64), (1, 4, 2048, 64)],
    "d128": [(1, 2, 256, 128), (2, 2, 1000, 128), (1, 2, 2048, 128)],
    "d32_96": [(1, 2, 512, 32), (1, 2, 333, 32), (1, 2, 512, 96), (1, 2, 333, 96)],
    "kat": [(1, 2, 512, 64)],
}
PERF = {
    "perf64": [(4, 48, 8192, 64)],
    "perf128": [(4, 64, 8192, 128)],
}
FFPA = {
    "ffpa": [(1, 2, 256, 256), (1, 2, 1000, 256), (1, 2, 512, 512), (1, 1, 384, 320), (1, 1, 256, 1024)],
}
FFPA_PERF = {"ffpa_perf": [(1, 32, 4096, 512), (1, 32, 4096, 256)]}


def ref_attn(q, k, v):
    import torch
    s = (q.float() @ k.float().transpose(-1, -2)) / (q.size(-1) ** 0.5)
    return torch.softmax(s, dim=-1) @ v.float()


def time_fn(torch, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def child(group):
    import torch
    from b200k import ops

    cases = {**GROUPS, **PERF, **FFPA, **FFPA_PERF}[group]
    perf = group.startswith("perf") or group.endswith("_perf")
    for (B, H, N, D) in cases:
        torch.manual_seed(1)
        if group == "kat":
            # the reference's deterministic fixture: all-ones Q/K/V -> O == 1 (flash_attn_mma.py:L353-369)
            q = torch.ones(B, H, N, D, dtype=torch.half, device="cuda")
            k, v = q.clone(), q.clone()
        else:
            q = torch.randn(B, H, N, D, dtype=torch.half, device="cuda")
            k = torch.randn(B, H, N, D, dtype=torch.half, device="cuda")
            v = torch.randn(B, H, N, D, dtype=torch.half, device="cuda")
        o = torch.full_like(q, float("nan"))
        fn = ops.ffpa_fwd if D > 128 else ops.fa2_fwd
        fn(q, k, v, o)
        torch.cuda.synchronize()
        res = {"group": group, "B": B, "H": H, "N": N, "D": D}
        if perf:
            # check one head only (the fp32 reference of the full problem is too big)
            ref = ref_attn(q[:1, :2], k[:1, :2], v[:1, :2])
            got = o[:1, :2].float()
        else:
            ref = ref_attn(q, k, v)
            got = o.float()
        err = (got - ref).abs()
        res["max_abs"] = err.max().item()
        res["mean_abs"] = err.mean().item()
        res["nonfinite"] = int((~torch.isfinite(o.float())).sum().item())
        res["ok"] = bool(torch.allclose(got, ref, rtol=1e-2, atol=1e-3)) and res["nonfinite"] == 0
        if perf:
            flops = 4.0 * B * H * N * N * D
            ms = time_fn(torch, lambda: fn(q, k, v, o))
            res["ms"] = ms
            res["tflops"] = flops / ms * 1e-9
            try:
                import torch.nn.functional as F
                from torch.nn.attention import SDPBackend, sdpa_kernel
                be = SDPBackend.FLASH_ATTENTION if D <= 256 else SDPBackend.EFFICIENT_ATTENTION
                with sdpa_kernel(be):
                    ms2 = time_fn(torch, lambda: F.scaled_dot_product_attention(q, k, v), iters=5, warm=2)
                res["sdpa_tflops"] = flops / ms2 * 1e-9
            except Exception as e:  # noqa
                res["sdpa_err"] = repr(e)[:200]
            if D <= 256:
                try:
                    from flash_attn import flash_attn_func
                    qq, kk, vv = [t.transpose(1, 2).contiguous() for t in (q, k, v)]
                    ms3 = time_fn(torch, lambda: flash_attn_func(qq, kk, vv), iters=5, warm=2)
                    res["fa2pkg_tflops"] = flops / ms3 * 1e-9
                except Exception as e:  # noqa
                    res["fa2pkg_err"] = repr(e)[:200]
        print("RESULT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        return
    os.makedirs(OUT, exist_ok=True)
    groups = dict(GROUPS)
    if "--perf" in sys.argv:
        groups.update(PERF)
    if "--ffpa" in sys.argv:
        groups.update(FFPA)
        if "--perf" in sys.argv:
            groups.update(FFPA_PERF)
    if "--only" in sys.argv:
        keep = sys.argv[sys.argv.index("--only") + 1].split(",")
        allg = {**GROUPS, **PERF, **FFPA, **FFPA_PERF}
        groups = {g: allg[g] for g in keep}
    nfail = 0
    with open(os.path.join(OUT, "attn_check.jsonl"), "a") as f:
        for g, cases in groups.items():
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--child", g], capture_output=True, text=True, timeout=300)
                out, err, rc = p.stdout, p.stderr, p.returncode
                timed_out = False
            except subprocess.TimeoutExpired as e:
                out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                err = e.stderr.decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
                rc, timed_out = -9, True
            results = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
            if len(results) < len(cases):
                results.append({"group": g, "ok": False, "rc": rc, "timeout": timed_out, "stderr": err[-1500:],
                                "stdout": "\n".join(l for l in out.splitlines() if not l.startswith("RESULT "))[-1500:]})
            for res in results:
                res["wall_s"] = round(time.time() - t0, 1)
                if not res.get("ok"):
                    nfail += 1
                f.write(json.dumps(res) + "\n")
                f.flush()
                print(json.dumps(res), flush=True)
    print("FAILURES: %d" % nfail)


if __name__ == "__main__":
    main()
