"""Generates tests/golden/* (run HERE, in the build container, where /root/reference is mounted).

  ref_exports.json   the export lists parsed out of the reference's pybind blocks — the names a drop-in must provide
  kat_*.npz          the reference's own known-answer fixtures (all-ones QKV, --range-k, histogram range(10)*1000)
  seeded_*.npz       small seeded inputs with the CPU oracle's outputs (hash-frozen restatement; see oracle/oracle.py)

The fixtures travel to the GPU box with the repo; nothing under tests/ reads /root/reference at run time.
"""
import json
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def mdefs(path):
    text = open(os.path.join(REF, path)).read()
    text = re.sub(r"//.*", "", text)
    names = re.findall(r"TORCH_BINDING_COMMON_EXTENSION\(\s*([A-Za-z0-9_]+)\s*\)", text)
    names += re.findall(r"m\.def\(\s*\"([A-Za-z0-9_]+)\"", text)
    return [n for n in dict.fromkeys(names) if n != "func"]


def main():
    os.makedirs(OUT, exist_ok=True)
    exports = {
        "toy_hgemm": mdefs("kernels/hgemm/pybind/hgemm.cc"),
        "flash_attn_lib": mdefs("kernels/flash-attn/pybind/flash_attn.cc"),
        "pyffpa_cuda": mdefs("ffpa-attn-mma/csrc/pybind/ffpa_attn_api.cc"),
        "elementwise_lib": mdefs("kernels/elementwise/elementwise.cu"),
        "block_all_reduce_lib": mdefs("kernels/reduce/block_all_reduce.cu"),
        "softmax_lib": mdefs("kernels/softmax/softmax.cu"),
        "rms_norm_lib": mdefs("kernels/rms-norm/rms_norm.cu"),
        "rope_lib": mdefs("kernels/rope/rope.cu"),
        "hist_lib": mdefs("kernels/histogram/histogram.cu"),
        "embedding_lib": mdefs("kernels/embedding/embedding.cu"),
        # second set
        "relu_lib": mdefs("kernels/relu/relu.cu"), "sigmoid_lib": mdefs("kernels/sigmoid/sigmoid.cu"),
        "gelu_lib": mdefs("kernels/gelu/gelu.cu"), "swish_lib": mdefs("kernels/swish/swish.cu"),
        "elu_lib": mdefs("kernels/elu/elu.cu"), "hardswish_lib": mdefs("kernels/hardswish/hardswish.cu"),
        "hardshrink_lib": mdefs("kernels/hardshrink/hardshrink.cu"),
        "layer_norm_lib": mdefs("kernels/layer-norm/layer_norm.cu"),
        "dot_product_lib": mdefs("kernels/dot-product/dot_product.cu"),
        "mat_transpose_lib": mdefs("kernels/mat-transpose/mat_transpose.cu"),
        "sgemv_lib": mdefs("kernels/sgemv/sgemv.cu"), "hgemv_lib": mdefs("kernels/hgemv/hgemv.cu"),
    }
    # macro-generated binding names the regex cannot see
    red = open(os.path.join(REF, "kernels/reduce/block_all_reduce.cu")).read()
    exports["block_all_reduce_lib"] = ["block_all_reduce_sum_%s_%s" % (a, b) for a, b in re.findall(
        r"^TORCH_BINDING_REDUCE\(\s*([a-z0-9_]+)\s*,\s*([a-z0-9_]+)\s*,", red, flags=re.M)]
    sm = open(os.path.join(REF, "kernels/softmax/softmax.cu")).read()
    exports["softmax_lib"] = sorted(set(exports["softmax_lib"]) | {"softmax_" + p for p in re.findall(
        r"^TORCH_BINDING_SOFTMAX\(\s*([a-z0-9_]+)\s*,", sm, flags=re.M)})
    hi = open(os.path.join(REF, "kernels/histogram/histogram.cu")).read()
    exports["hist_lib"] = ["histogram_" + p for p in re.findall(r"^TORCH_BINDING_HIST\(\s*([a-z0-9_]+)\s*,", hi, flags=re.M)]
    for k, v in exports.items():
        print(k, len(v))
    json.dump(exports, open(os.path.join(OUT, "ref_exports.json"), "w"), indent=1, sort_keys=True)

    # ---- known-answer fixtures the reference itself defines
    B, H, N, D = 1, 2, 256, 64
    ones = torch.ones(B, H, N, D, dtype=torch.half)
    o = oracle.attention(ones, ones, ones)
    np.savez_compressed(os.path.join(OUT, "kat_attention_all_ones.npz"), shape=np.array([B, H, N, D]), o=o.numpy())
    torch.manual_seed(20260922)
    q = torch.randn(B, H, N, D).half()
    v = torch.randn(B, H, N, D).half()
    k = oracle.make_range_k(B, H, N, D)
    o = oracle.attention(q, k, v)
    np.savez_compressed(os.path.join(OUT, "kat_attention_range_k.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(), o=o.numpy())
    a = np.array(list(range(10)) * 1000, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "kat_histogram.npz"), a=a, hist=np.full(10, 1000, dtype=np.int32))

    # ---- seeded cases with oracle outputs
    torch.manual_seed(1)
    A = torch.randn(96, 72).half()
    Bm = torch.randn(72, 80).half()
    np.savez_compressed(os.path.join(OUT, "seeded_hgemm.npz"), a=A.numpy(), b=Bm.numpy(), c=oracle.hgemm(A, Bm).numpy(),
                        c_f16acc=oracle.hgemm_f16acc_k16(A[:, :64], Bm[:64]).numpy())
    for D in (32, 64, 96, 128, 256, 320):
        torch.manual_seed(100 + D)
        q, k, v = [torch.randn(1, 2, 200, D).half() for _ in range(3)]
        np.savez_compressed(os.path.join(OUT, "seeded_attention_d%d.npz" % D), q=q.numpy(), k=k.numpy(), v=v.numpy(),
                            o=oracle.attention(q, k, v).numpy())
    torch.manual_seed(7)
    x = torch.randn(37, 200)
    np.savez_compressed(os.path.join(OUT, "seeded_rows.npz"), x=x.numpy(), softmax=oracle.softmax_per_token(x).numpy(),
                        softmax_all=oracle.softmax_all(x).numpy(), rms=oracle.rms_norm(x, 1.0).numpy(),
                        rms_eps_in_k=oracle.rms_norm(x, 1.0, eps_inside_k=True).numpy(),
                        rope=oracle.rope(x, False).numpy(), rope_quirk=oracle.rope(x, True).numpy(),
                        sum=np.array(oracle.reduce_sum(x)))


if __name__ == "__main__":
    main()
