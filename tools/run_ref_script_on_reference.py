"""TEST INFRASTRUCTURE: run an unmodified reference bench script on the REFERENCE's own kernels (oracle/_ref/ref_<op>_lib.so,
built by `oracle/build_ref.py support`), for the side-by-side logs next to the drop-in runs.
Same mechanics as b200k.run_ref_script (patch cpp_extension.load by name, chdir, runpy) — only the lookup differs.

    python tools/run_ref_script_on_reference.py baseline/_ref/kernels/rope/rope.py
"""
import importlib.util
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
# load(name=...) used by each script -> key of the reference build
KEY = {"elementwise_lib": "elementwise", "block_all_reduce_lib": "reduce", "softmax_lib": "softmax", "rms_norm_lib": "rms_norm",
       "rope": "rope", "rope_lib": "rope", "hist_lib": "histogram", "embedding": "embedding", "embedding_lib": "embedding",
       "relu_lib": "relu", "sigmoid_lib": "sigmoid", "gelu_lib": "gelu", "swish_lib": "swish", "elu_lib": "elu",
       "hardswish_lib": "hardswish", "hardshrink_lib": "hardshrink", "layer_norm_lib": "layer_norm",
       "dot_product_lib": "dot_product", "mat_transpose_lib": "mat_transpose", "sgemv_lib": "sgemv", "hgemv_lib": "hgemv"}


def main():
    script = os.path.abspath(sys.argv[1])
    import torch.utils.cpp_extension as ce

    def load(name, *a, **kw):
        mod_name = "ref_%s_lib" % KEY[name]
        path = os.path.join(REF_DIR, mod_name + ".so")
        print("[reference] cpp_extension.load(name=%r) -> %s (unmodified reference kernels, sm_100a build)" % (name, os.path.relpath(path, ROOT)))
        spec = importlib.util.spec_from_file_location(mod_name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ce.load = load
    os.chdir(os.path.dirname(script))
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
