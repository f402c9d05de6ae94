"""Run an UNMODIFIED reference bench script on the B200 kernels.

    python -m b200k.run_ref_script /path/to/CUDA-Learn-Notes/kernels/flash-attn/flash_attn_mma.py --B 4 --H 48 ...

`hgemm.py` and `test_ffpa_attn.py` first try `import toy_hgemm` / `import ffpa_attn` (kernels/hgemm/tools/utils.py:L116-121,
ffpa-attn-mma/env.py:L386-395) — having this directory on sys.path is enough for them.  `flash_attn_mma.py` and the small
`<op>.py` scripts always call `torch.utils.cpp_extension.load(name=..., sources=...)` (flash_attn_mma.py:L177-181); this
launcher replaces `load` by a lookup of our drop-in namespace for that `name`, chdirs to the script's directory (the
scripts use relative paths) and hands control to the script via runpy.  Unknown names fall through to the real `load`.
"""
from __future__ import annotations

import os
import runpy
import sys


def _namespaces():
    from . import flash_attn_lib, support_libs
    import toy_hgemm

    table = dict(support_libs.BY_LOAD_NAME)
    for alias, name in support_libs.LOAD_NAME_ALIASES.items():
        table[alias] = table[name]
    table["flash_attn_lib"] = flash_attn_lib
    table["hgemm_lib"] = toy_hgemm
    return table


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    import torch.utils.cpp_extension as ce

    real_load = ce.load
    table = _namespaces()

    def load(name, *a, **kw):
        if name in table:
            print("[b200k] cpp_extension.load(name=%r) -> B200 drop-in" % name)
            return table[name]
        return real_load(name, *a, **kw)

    ce.load = load
    os.chdir(os.path.dirname(script))
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
