"""b200k — B200-native (sm_100a) kernels behind the CUDA-Learn-Notes call signatures.

Layout:  _loader (ctypes binding of libb200k.so, the C ABI in include/b200k.h) -> ops (tensor-level wrappers)
-> drop-in shims: `toy_hgemm`, `ffpa_attn` (packages next to this one), `b200k.flash_attn_lib`,
`b200k.support_libs`; `b200k.sharded` holds the batch-sharded multi-GPU attention path.
"""
from . import _loader  # noqa: F401  (raises if libb200k.so is missing: no fallback)
from . import ops  # noqa: F401

__version__ = "0.1.0"
