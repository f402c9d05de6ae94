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


def faster_prefill_attn_func(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    o: Optional[torch.Tensor] = None,
    num_stages: int = 2,
    level: LevelType = LevelType.L1,
    acc: MMAAccType = MMAAccType.FP32,
):
    # Q, K, V, O: [B, H, N, D] layout
    if not isinstance(o, torch.Tensor) or o is None:
        o = torch.zeros_like(q)
    assert level == LevelType.L1, "only support FFPA L1 level now."
    if acc == MMAAccType.FP32:
        ffpa_mma_acc_f32_L1(q, k, v, o, num_stages)
    else:
        ffpa_mma_acc_f16_L1(q, k, v, o, num_stages)
    return o


ffpa: callable = faster_prefill_attn_func
ffpa_acc_f32_L1 = partial(faster_prefill_attn_func, level=LevelType.L1, acc=MMAAccType.FP32)
ffpa_acc_f16_L1 = partial(faster_prefill_attn_func, level=LevelType.L1, acc=MMAAccType.FP16)
