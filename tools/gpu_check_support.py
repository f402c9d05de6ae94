"""GPU check + bandwidth measurement of the support kernels (run on the B200 box through gpurun).
Results: gpurun_out/support_check.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
OUT = os.path.join(ROOT, "gpurun_out")


def time_fn(torch, fn, iters=20, warm=3):
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


def main():
    import torch
    from b200k import ops

    os.makedirs(OUT, exist_ok=True)
    results = []

    def rec(name, ok, bytes_moved=None, fn=None, **kw):
        r = {"kernel": name, "ok": bool(ok), **kw}
        if fn is not None and bytes_moved:
            ms = time_fn(torch, fn)
            r["ms"] = ms
            r["GBps"] = bytes_moved / ms * 1e-6
        results.append(r)
        print(json.dumps(r), flush=True)

    torch.manual_seed(0)
    dev = "cuda"
    # ---- elementwise add (bit exact)
    for dt in (torch.float32, torch.float16):
        for n in (4096 * 4096, 1000003):
            a = torch.randn(n, dtype=dt, device=dev)
            b = torch.randn(n, dtype=dt, device=dev)
            c = torch.empty_like(a)
            ops.elementwise_add(a, b, c)
            rec("elementwise_add_%s_n%d" % (str(dt)[6:], n), torch.equal(c, a + b), 3 * n * a.element_size(),
                lambda: ops.elementwise_add(a, b, c))
    # ---- reduce
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        x = torch.randn(4096, 4096, dtype=dt, device=dev)
        y = ops.block_all_reduce_sum(x)
        ref = x.double().sum().item()
        rec("reduce_%s" % str(dt)[6:], abs(y.item() - ref) <= 1e-3 * (4096 * 4096) ** 0.5 + 1e-6 * abs(ref),
            x.numel() * x.element_size(), lambda: ops.block_all_reduce_sum(x), got=y.item(), ref=ref)
        y2 = ops.block_all_reduce_sum(x)
        rec("reduce_%s_deterministic" % str(dt)[6:], y.item() == y2.item())
    xi = torch.randint(-128, 128, (4096 * 4096 + 5,), dtype=torch.int8, device=dev)
    yi = ops.block_all_reduce_sum(xi)
    rec("reduce_i8", yi.item() == int(xi.long().sum().item()), xi.numel(), lambda: ops.block_all_reduce_sum(xi))
    if hasattr(torch, "float8_e4m3fn"):
        for dt in (torch.float8_e4m3fn, torch.float8_e5m2):
            x8 = (torch.randn(1024 * 1024, device=dev) * 0.5).to(dt)
            y8 = ops.block_all_reduce_sum(x8)
            ref8 = x8.float().double().sum().item()
            rec("reduce_%s" % str(dt)[6:], abs(y8.item() - ref8) < 2.0, x8.numel(), lambda: ops.block_all_reduce_sum(x8),
                got=y8.item(), ref=ref8)
    # ---- softmax
    for dt in (torch.float32, torch.float16):
        for (S, H) in ((4096, 256), (4096, 1024), (4096, 4096), (4096, 8192), (512, 16384), (333, 1000), (64, 77)):
            x = torch.randn(S, H, dtype=dt, device=dev)
            y = torch.empty_like(x)
            for mode in (1, 2, 3):
                ops.softmax(x, y, mode)
                ref = torch.softmax(x.float(), dim=-1)
                tol = 1e-5 if dt == torch.float32 else 1e-3
                ok = torch.allclose(y.float(), ref, rtol=1e-3, atol=tol)
                if mode == 2:
                    rec("softmax_%s_%dx%d" % (str(dt)[6:], S, H), ok, 2 * x.numel() * x.element_size(),
                        lambda: ops.softmax(x, y, 2))
                else:
                    rec("softmax_%s_%dx%d_mode%d" % (str(dt)[6:], S, H, mode), ok)
    x = torch.randn(4096, 1024, device=dev)
    y = torch.empty_like(x)
    ops.softmax(x, y, 0)
    ref = torch.softmax(x.flatten(), 0).view_as(x)
    rec("softmax_all_f32", torch.allclose(y, ref, rtol=1e-3, atol=1e-9))
    # ---- rms norm
    for dt in (torch.float32, torch.float16):
        for (N, K) in ((4096, 512), (4096, 1024), (4096, 4096), (4096, 8192), (100, 1000)):
            x = torch.randn(N, K, dtype=dt, device=dev)
            y = torch.empty_like(x)
            ops.rms_norm(x, y, 1.0)
            xf = x.float()
            ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)
            ok = torch.allclose(y.float(), ref, rtol=1e-3, atol=1e-5 if dt == torch.float32 else 2e-3)
            rec("rms_norm_%s_%dx%d" % (str(dt)[6:], N, K), ok, 2 * x.numel() * x.element_size(),
                lambda: ops.rms_norm(x, y, 1.0))
    # ---- rope
    for (M, N) in ((4096, 512), (8192, 1024), (100, 6)):
        x = torch.randn(M, N, device=dev)
        out = torch.empty_like(x)
        ops.rope_f32(x, out, ref_quirk=False)
        xc = torch.view_as_complex(x.reshape(M, N // 2, 2))
        freqs = 1.0 / (10000.0 ** (torch.arange(0, N, 2, device=dev).float() / N))
        ang = torch.outer(torch.arange(M, device=dev).float(), freqs)
        ref = torch.view_as_real(xc * torch.polar(torch.ones_like(ang), ang)).flatten(1)
        ok = torch.allclose(out, ref, rtol=1e-3, atol=2e-3)
        ops.rope_f32(x, out, ref_quirk=True)
        ang2 = torch.arange(M, device=dev).float()[:, None].expand(M, N // 2)
        ref2 = torch.view_as_real(xc * torch.polar(torch.ones_like(ang2), ang2.contiguous())).flatten(1)
        ok2 = torch.allclose(out, ref2, rtol=1e-3, atol=2e-3)
        rec("rope_f32_%dx%d" % (M, N), ok and ok2, 2 * x.numel() * 4, lambda: ops.rope_f32(x, out, True), textbook=ok, quirk=ok2)
    # ---- histogram (bit exact) incl. the reference's fixture: range(10)*1000 -> 1000 per bin
    a = torch.tensor(list(range(10)) * 1000, dtype=torch.int32, device=dev)
    h = ops.histogram_i32(a)
    rec("histogram_fixture", h.tolist() == [1000] * 10)
    a = torch.randint(0, 50000, (10_000_000,), dtype=torch.int32, device=dev)
    h = ops.histogram_i32(a)
    rec("histogram_10M_50000bins", torch.equal(h.long(), torch.bincount(a.long(), minlength=h.numel())), a.numel() * 4,
        lambda: ops.histogram_i32(a, nbins=50000))
    a = torch.randint(0, 256, (10_000_003,), dtype=torch.int32, device=dev)
    h = ops.histogram_i32(a)
    rec("histogram_10M_256bins", torch.equal(h.long(), torch.bincount(a.long(), minlength=h.numel())), a.numel() * 4,
        lambda: ops.histogram_i32(a, nbins=256))
    # ---- embedding (bit exact)
    for dt in (torch.float32, torch.float16):
        for (n, rows, emb) in ((4096, 1024, 1024), (2048, 1024, 512), (1000, 77, 100)):
            w = torch.randn(rows, emb, dtype=dt, device=dev)
            idx = torch.randint(0, rows, (n,), dtype=torch.int32, device=dev)
            out = torch.empty(n, emb, dtype=dt, device=dev)
            ops.embedding(idx, w, out)
            rec("embedding_%s_%dx%d" % (str(dt)[6:], n, emb), torch.equal(out, w[idx.long()]), 2 * out.numel() * out.element_size(),
                lambda: ops.embedding(idx, w, out))
    nfail = sum(1 for r in results if not r["ok"])
    with open(os.path.join(OUT, "support_check.jsonl"), "w") as f:
        for r in results:
            f.write(json.dumps(r) + "\n")
    print("FAILURES: %d / %d" % (nfail, len(results)))


if __name__ == "__main__":
    main()
