"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference kernels into oracle/_ref/ (git-ignored, travels to the GPU box).

The reference (DefTruth/CUDA-Learn-Notes) is CUDA; it compiles here (nvcc cross-compiles for sm_100a) but can only
run on the GPU box.  Sources are compiled from where they lie under /root/reference — nothing is copied.  We do
not run the reference's own build system (setup.py / JIT `load`); the flags below restate the ones it passes
(kernels/hgemm/tools/utils.py:L58-93, kernels/flash-attn/flash_attn_mma.py:L120-156, ffpa-attn-mma/env.py:L190-340)
with the arch switched to sm_100a — exactly what its JIT path would produce on a B200.

Products (all optional; every consumer checks for presence):
  oracle/_ref/libref_hgemm.so        C-ABI shim around hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem  (torch-free, ~10 s)
  oracle/_ref/ref_flash_attn_lib.so  torch extension: flash_attn_mma_stages_split_q_shared_qkv{,_acc_f32}  (~2-4 min)
  oracle/_ref/pyffpa_cuda.so         torch extension: the reference's own 3-file FFPA module                (long)
  oracle/_ref/ref_<op>_lib.so        torch extensions, one per bandwidth-kernel TU of SURVEY 8(a) a6-a12 and 8(f)-3
                                     (kernels/<op>/<op>.cu with its own PYBIND11_MODULE block; ~1-2 min each, 8 at a time)

usage: python oracle/build_ref.py [hgemm] [flash] [ffpa] [support]      (default: hgemm flash)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200K_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
          "-U__CUDA_NO_HALF2_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "--expt-relaxed-constexpr",
          "--expt-extended-lambda", "--use_fast_math", "-diag-suppress", "177", "-Xcompiler", "-fPIC"]


def run(cmd, log):
    t0 = time.time()
    with open(log, "w") as f:
        rc = subprocess.call(cmd, stdout=f, stderr=subprocess.STDOUT)
    print("[build_ref] rc=%d %.0fs  %s" % (rc, time.time() - t0, " ".join(cmd[:1] + cmd[-3:])), flush=True)
    return rc


def torch_flags(name):
    import torch
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    defs = ["-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["-L" + libdir, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
            "-Xlinker", "-rpath", "-Xlinker", libdir]
    return inc, defs, link


def build_hgemm():
    h = os.path.join(REF, "kernels", "hgemm")
    out = os.path.join(OUT, "libref_hgemm.so")
    cmd = ["nvcc", *ARCH, *COMMON, "-I", os.path.join(h, "utils"), "-I", os.path.join(h, "mma", "basic"),
           "-shared", os.path.join(HERE, "ref_shims", "ref_hgemm_shim.cu"), "-o", out, "-lcublas"]
    return run(cmd, os.path.join(OUT, "build_hgemm.log"))


def build_torch_ext(name, sources, extra_inc, extra_defs, jobs=2):
    inc, defs, link = torch_flags(name)
    objs, procs = [], []
    for src in sources:
        obj = os.path.join(OUT, name + "_" + os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = ["nvcc", *ARCH, *COMMON, *defs, *extra_defs, *inc]
        for i in extra_inc:
            cmd += ["-I", i]
        cmd += ["-x", "cu", "-c", src, "-o", obj]
        log = open(obj + ".log", "w")
        procs.append((subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), src, time.time()))
        while sum(1 for p, _, _ in procs if p.poll() is None) >= jobs:
            time.sleep(1)
    rc = 0
    for p, src, t0 in procs:
        r = p.wait()
        print("[build_ref] rc=%d %.0fs  %s" % (r, time.time() - t0, src), flush=True)
        rc |= r
    if rc:
        return rc
    out = os.path.join(OUT, name + ".so")
    return run(["nvcc", *ARCH, "-shared", *objs, "-o", out, *link], os.path.join(OUT, "link_" + name + ".log"))


def build_flash():
    fa = os.path.join(REF, "kernels", "flash-attn")
    srcs = [os.path.join(fa, "mma", "basic", "flash_attn_mma_share_qkv.cu"),
            os.path.join(fa, "mma", "basic", "flash_attn_mma_share_qkv_F32F16F16F32.cu"),
            os.path.join(HERE, "ref_shims", "ref_flash_attn_pybind.cc")]
    inc = [fa, os.path.join(fa, "utils"), os.path.join(fa, "mma"), os.path.join(fa, "mma", "basic")]
    return build_torch_ext("ref_flash_attn_lib", srcs, inc, [], jobs=3)


def build_ffpa():
    ff = os.path.join(REF, "ffpa-attn-mma")
    srcs = [os.path.join(ff, "csrc", "pybind", "ffpa_attn_api.cc"),
            os.path.join(ff, "csrc", "cuffpa", "ffpa_attn_F16F16F16_L1.cu"),
            os.path.join(ff, "csrc", "cuffpa", "ffpa_attn_F16F16F32_L1.cu")]
    # default ENABLE_FFPA_* environment of the reference (env.py:L12-106): all stages, prefetch, swizzle q/k/v
    defs = ["-DENABLE_FFPA_ALL_STAGES", "-DENABLE_FFPA_PREFETCH_QKV", "-DENABLE_FFPA_SMEM_SWIZZLE_Q",
            "-DENABLE_FFPA_SMEM_SWIZZLE_K", "-DENABLE_FFPA_SMEM_SWIZZLE_V"]
    inc = [os.path.join(ff, "include"), os.path.join(ff, "csrc", "cuffpa")]
    return build_torch_ext("pyffpa_cuda", srcs, inc, defs, jobs=3)


# load name -> reference TU (kernels/<dir>/<file>.cu); the names are the ones each <op>.py passes to
# torch.utils.cpp_extension.load(name=...), prefixed with ref_ so they never collide with the product's shims.
SUPPORT_TUS = {
    "elementwise": ("elementwise", "elementwise.cu"), "reduce": ("reduce", "block_all_reduce.cu"),
    "softmax": ("softmax", "softmax.cu"), "rms_norm": ("rms-norm", "rms_norm.cu"), "rope": ("rope", "rope.cu"),
    "histogram": ("histogram", "histogram.cu"), "embedding": ("embedding", "embedding.cu"),
    "relu": ("relu", "relu.cu"), "sigmoid": ("sigmoid", "sigmoid.cu"), "gelu": ("gelu", "gelu.cu"),
    "swish": ("swish", "swish.cu"), "elu": ("elu", "elu.cu"), "hardswish": ("hardswish", "hardswish.cu"),
    "hardshrink": ("hardshrink", "hardshrink.cu"), "layer_norm": ("layer-norm", "layer_norm.cu"),
    "dot_product": ("dot-product", "dot_product.cu"), "mat_transpose": ("mat-transpose", "mat_transpose.cu"),
    "sgemv": ("sgemv", "sgemv.cu"), "hgemv": ("hgemv", "hgemv.cu"),
}


def build_support(jobs=8):
    """One torch-extension .so per reference TU, flags as in every kernels/<op>/<op>.py (e.g. rope/rope.py:L13-27)."""
    procs, rc = [], 0
    for key, (d, f) in SUPPORT_TUS.items():
        name = "ref_%s_lib" % key
        inc, defs, link = torch_flags(name)
        src = os.path.join(REF, "kernels", d, f)
        out = os.path.join(OUT, name + ".so")
        cmd = ["nvcc", *ARCH, *COMMON, *defs, *inc, "-shared", "-x", "cu", src, "-o", out, *link]
        log = open(os.path.join(OUT, "build_%s.log" % name), "w")
        procs.append((subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), src, time.time()))
        while sum(1 for p, _, _ in procs if p.poll() is None) >= jobs:
            time.sleep(1)
    for p, src, t0 in procs:
        r = p.wait()
        print("[build_ref] rc=%d %.0fs  %s" % (r, time.time() - t0, src), flush=True)
        rc |= r
    return rc


def main():
    if not os.path.isdir(REF):
        print("[build_ref] %s not present (GPU box?) - using prebuilt oracle/_ref if any" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ["hgemm", "flash"]
    rc = 0
    if "hgemm" in what:
        rc |= build_hgemm()
    if "flash" in what:
        rc |= build_flash()
    if "ffpa" in what:
        rc |= build_ffpa()
    if "support" in what:
        rc |= build_support()
    return rc


if __name__ == "__main__":
    sys.exit(main())
