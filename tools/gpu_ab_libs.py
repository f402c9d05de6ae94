#!/usr/bin/env python
"""Round-robin A/B timing of two builds of libb200k.so on the same box.

The B200's clocks depend on its recent power/thermal history, so two numbers taken minutes apart are not comparable;
this tool alternates the libraries launch-set by launch-set and prints the per-round figures and the medians.

    python tools/gpu_ab_libs.py build_ab/libb200k_old.so cuda-learn-notes_b200/b200k/libb200k.so [--rounds 7]
"""
import argparse
import ctypes
import json
import statistics
import sys
from ctypes import c_float, c_int, c_int64, c_void_p

import torch


def load(path):
    lib = ctypes.CDLL(path)
    lib.b200k_hgemm_f16.restype = c_int
    lib.b200k_hgemm_f16.argtypes = [c_void_p] * 3 + [c_int64] * 3 + [c_int, c_int, c_void_p]
    lib.b200k_fa2_fwd_f16.restype = c_int
    lib.b200k_fa2_fwd_f16.argtypes = [c_void_p] * 4 + [c_int64] * 4 + [c_float, c_int, c_int, c_void_p]
    lib.b200k_ffpa_fwd_f16.restype = c_int
    lib.b200k_ffpa_fwd_f16.argtypes = [c_void_p] * 4 + [c_int64] * 4 + [c_float, c_int, c_void_p]
    return lib


def time_ms(fn, iters):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        fn(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    libs = [load(p) for p in args.libs]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)

    cases = []
    # hgemm 8192^3 and 4096^3 (variant auto)
    for n in (2048, 4096, 8192):
        a = torch.randn(n, n, device=dev, dtype=torch.float16)
        b = torch.randn(n, n, device=dev, dtype=torch.float16)
        c = torch.empty(n, n, device=dev, dtype=torch.float16)
        cases.append(("hgemm_%d" % n, 2.0 * n ** 3, 20,
                      lambda lib, st, a=a, b=b, c=c, n=n: lib.b200k_hgemm_f16(a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                                                              n, n, n, 0, 0, st)))
    for (B, H, N, D) in ((4, 48, 8192, 64), (4, 64, 8192, 128)):
        q, k, v = (torch.randn(B, H, N, D, device=dev, dtype=torch.float16) for _ in range(3))
        o = torch.empty_like(q)
        cases.append(("fa2_d%d" % D, 4.0 * B * H * N * N * D, 5,
                      lambda lib, st, q=q, k=k, v=v, o=o, s=(B, H, N, D): lib.b200k_fa2_fwd_f16(
                          q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), *s, 0.0, 0, 0, st)))
    # variant 0x400 at D = 512: the D-sliced CTA-pair kernel (the default there is the O^T kernel)
    for (B, H, N, D, var) in ((1, 32, 8192, 256, 0), (1, 32, 4096, 512, 0), (1, 32, 4096, 512, 0x400), (1, 32, 4096, 768, 0),
                              (1, 32, 4096, 1024, 0), (1, 32, 8192, 320, 0)):
        q, k, v = (torch.randn(B, H, N, D, device=dev, dtype=torch.float16) for _ in range(3))
        o = torch.empty_like(q)
        cases.append(("ffpa_d%d_v%x" % (D, var), 4.0 * B * H * N * N * D, 5,
                      lambda lib, st, q=q, k=k, v=v, o=o, s=(B, H, N, D), var=var: lib.b200k_ffpa_fwd_f16(
                          q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), *s, 0.0, var, st)))
    if args.only:
        cases = [c for c in cases if args.only in c[0]]

    for name, flops, iters, fn in cases:
        res = [[] for _ in libs]
        for r in range(args.rounds):
            for i, lib in enumerate(libs):
                def run(st, lib=lib):
                    rc = fn(lib, st)
                    if rc != 0:
                        raise RuntimeError("%s rc=%d" % (name, rc))
                ms = time_ms(run, iters)
                res[i].append(flops / ms / 1e9)
        out = {"case": name}
        for i, p in enumerate(args.libs):
            out["lib%d_median_tflops" % i] = round(statistics.median(res[i]), 1)
            out["lib%d_rounds" % i] = [round(x) for x in res[i]]
        print(json.dumps(out))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
