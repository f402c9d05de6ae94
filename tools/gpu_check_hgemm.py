"""GPU bring-up check for the tcgen05 HGEMM (run on the B200 box through gpurun).

Each (variant, layout, shape) case runs in its own subprocess with a timeout, so a trap or a hang in one
configuration does not take the others down.  Results go to gpurun_out/hgemm_check.jsonl.
Usage: python tools/gpu_check_hgemm.py [--perf]
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda-learn-notes_b200", "b200k", "libb200k.so")
OUT = os.path.join(ROOT, "gpurun_out")


def child(variant, tn, shapes, perf):
    import torch
    for (M, N, K) in shapes:
        child_one(torch, variant, tn, M, N, K, perf)


def child_one(torch, variant, tn, M, N, K, perf):

    lib = ctypes.CDLL(LIB)
    lib.b200k_last_error.restype = ctypes.c_char_p
    lib.b200k_hgemm_f16.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 3 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    torch.manual_seed(1)
    a = torch.randn(M, K, dtype=torch.half, device="cuda")
    b = torch.randn(K, N, dtype=torch.half, device="cuda")
    c = torch.full((M, N), float("nan"), dtype=torch.half, device="cuda")
    bb = b.t().contiguous() if tn else b
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        rc = lib.b200k_hgemm_f16(a.data_ptr(), bb.data_ptr(), c.data_ptr(), M, N, K, int(tn), variant, stream)
        if rc != 0:
            raise RuntimeError("rc=%d %s" % (rc, lib.b200k_last_error().decode()))

    run()
    torch.cuda.synchronize()
    ref = a.float() @ b.float()
    err = (c.float() - ref).abs()
    rel = err.max().item() / max(ref.abs().max().item(), 1e-6)
    bad = int((~torch.isfinite(c.float())).sum().item())
    # fp16 output rounding alone: relative 2^-11 per element
    tol_ok = bool(torch.allclose(c.float(), ref, rtol=2e-3, atol=2e-3 * (K ** 0.5)))
    res = {"variant": variant, "tn": tn, "M": M, "N": N, "K": K, "max_abs": err.max().item(), "rel_to_max": rel,
           "nonfinite": bad, "ok": tol_ok and bad == 0}
    if perf:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        iters = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms * 1e-9
        cc = torch.empty_like(c)
        for _ in range(3):
            torch.matmul(a, b, out=cc)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b, out=cc)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        res["torch_tflops"] = 2.0 * M * N * K / ms2 * 1e-9
    print("RESULT " + json.dumps(res), flush=True)


SHAPES = [(128, 256, 64), (256, 256, 256), (512, 768, 320), (300, 520, 264), (1024, 1024, 1024)]
PERF_SHAPES = [(2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192)]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        v, tn, perf = [int(x) for x in sys.argv[2:5]]
        child(v, tn, PERF_SHAPES if perf else SHAPES, perf)
        return
    perf = "--perf" in sys.argv
    os.makedirs(OUT, exist_ok=True)
    cases = [(v, tn, 0) for v in (1, 2, 3) for tn in (0, 1)]
    if perf:
        cases += [(v, tn, 1) for v in (1, 2, 3) for tn in (0, 1)]
    nfail = 0
    with open(os.path.join(OUT, "hgemm_check.jsonl"), "w") as f:
        for case in cases:
            t0 = time.time()
            results = []
            try:
                p = subprocess.run([sys.executable, __file__, "--child"] + [str(x) for x in case],
                                   capture_output=True, text=True, timeout=180)
                results = [json.loads(l[7:]) for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                expect = len(PERF_SHAPES if case[2] else SHAPES)
                if len(results) < expect:
                    results.append({"case": case, "ok": False, "rc": p.returncode,
                                    "stderr": p.stderr[-800:], "stdout": p.stdout[-800:]})
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                results = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
                results.append({"case": case, "ok": False, "timeout": True})
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
