"""Drop-in `ffpa_attn` package (ffpa-attn-mma/ffpa_attn/__init__.py:L1-20): same names, enums and signatures,
served by the tcgen05 FFPA kernel (b200k_ffpa_fwd_f16) instead of the `pyffpa_cuda` extension."""
from .interface import (  # noqa: F401
    faster_prefill_attn_func,
    ffpa,
    ffpa_acc_f16_L1,
    ffpa_acc_f32_L1,
    ffpa_mma_acc_f16_L1,
    ffpa_mma_acc_f32_L1,
    LevelType,
    MMAAccType,
)
from .version import __version__  # noqa: F401

L1 = LevelType.L1
L2 = LevelType.L2
L3 = LevelType.L3
FP32 = MMAAccType.FP32
FP16 = MMAAccType.FP16
