"""Mirror of ffpa-attn-mma/ffpa_attn/interface.py:L1-48 on top of the B200 kernel.

`ffpa_mma_acc_f16_L1` / `ffpa_mma_acc_f32_L1` are the raw entry points the reference binds from pyffpa_cuda
(csrc/pybind/ffpa_attn_api.cc:L8-17): `(Q, K, V, O, stages) -> None`.  Both accumulate in fp32 tensor memory here
(the reference's f16 variant accumulates QK^T and PV in half); `stages` is accepted and ignored.
"""
from enum import Enum
from functools import partial
from typing import Optional

import torch

from b200k import ops as _ops


class LevelType(Enum):
    L1 = 0
    L2 = 1
    L3 = 2


class MMAAccType(Enum):
    FP32 = 0
    FP16 = 1


def ffpa_mma_acc_f16_L1(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, O: torch.Tensor, stages: int) -> None:
    _ops.ffpa_fwd(Q, K, V, O)


def ffpa_mma_acc_f32_L1(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, O: torch.Tensor, stages: int) -> None:
    _ops.ffpa_fwd(Q, K, V, O)


_RAW_ENTRY = {MMAAccType.FP32: ffpa_mma_acc_f32_L1, MMAAccType.FP16: ffpa_mma_acc_f16_L1}


def faster_prefill_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    o: Optional[torch.Tensor] = None,
    num_stages: int = 2,
    level: LevelType = LevelType.L1,
    acc: MMAAccType = MMAAccType.FP32,
):
    """Same signature and defaults as the reference's public entry point (interface.py:L22-39): [B, H, N, D] tensors in,
    `o` returned (allocated zero-filled when the caller passes none), only level L1 exists."""
    if level != LevelType.L1:
        raise AssertionError("only support FFPA L1 level now.")
    out = o if isinstance(o, torch.Tensor) else torch.zeros_like(q)
    _RAW_ENTRY[MMAAccType(acc)](q, k, v, out, num_stages)
    return out


ffpa = faster_prefill_attn_func
ffpa_acc_f32_L1 = partial(faster_prefill_attn_func, level=LevelType.L1, acc=MMAAccType.FP32)
ffpa_acc_f16_L1 = partial(faster_prefill_attn_func, level=LevelType.L1, acc=MMAAccType.FP16)
