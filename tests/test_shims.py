"""CPU: the drop-in modules expose exactly the names / arities of the reference's pybind blocks (frozen in
tests/golden/ref_exports.json by tools/gen_golden.py) and reproduce its error conventions."""
import inspect

import pytest
import torch


def test_toy_hgemm_exports(ref_exports):
    import toy_hgemm

    want = set(ref_exports["toy_hgemm"])
    assert len(want) == 38
    have = set(toy_hgemm.__all__)
    assert have == want, (want - have, have - want)
    for n in want:
        fn = getattr(toy_hgemm, n)
        npar = len(inspect.signature(fn).parameters)
        if n in ("init_cublas_handle", "destroy_cublas_handle"):
            assert npar == 0
        elif "stages" in n or n == "hgemm_mma_stages_block_swizzle_tn_cute":
            assert npar == 6, n  # (a, b, c, stages, swizzle, swizzle_stride): hgemm.cc:L36-55
        else:
            assert npar == 3, n


def test_flash_attn_lib_exports(ref_exports):
    from b200k import flash_attn_lib

    want = set(ref_exports["flash_attn_lib"])
    assert len(want) == 28
    assert set(flash_attn_lib.NAMES) == want
    for n in want:
        assert len(inspect.signature(getattr(flash_attn_lib, n)).parameters) == 5  # (Q, K, V, O, stages)


def test_ffpa_attn_package(ref_exports):
    import ffpa_attn

    for n in ref_exports["pyffpa_cuda"] + ["ffpa", "ffpa_acc_f32_L1", "ffpa_acc_f16_L1", "faster_prefill_attn_func",
                                           "LevelType", "MMAAccType", "L1", "L2", "L3", "FP32", "FP16", "__version__"]:
        assert hasattr(ffpa_attn, n), n
    sig = inspect.signature(ffpa_attn.ffpa)
    assert list(sig.parameters) == ["q", "k", "v", "o", "num_stages", "level", "acc"]  # interface.py:L22-30
    assert sig.parameters["num_stages"].default == 2
    with pytest.raises(AssertionError):
        ffpa_attn.ffpa(torch.zeros(1, 1, 8, 256, dtype=torch.half), None, None, level=ffpa_attn.L2)


def test_support_lib_exports(ref_exports):
    from b200k import support_libs

    for lib_name, ns in support_libs.BY_LOAD_NAME.items():
        want = set(ref_exports[lib_name])
        have = {k for k in vars(ns)}
        assert have == want, (lib_name, want - have, have - want)


def test_error_conventions_match_reference():
    import toy_hgemm
    from b200k import flash_attn_lib, ops

    a = torch.zeros(16, 16)  # float32, CPU
    with pytest.raises(RuntimeError, match="values must be torch::kHalf"):
        toy_hgemm.hgemm_naive_f16(a, a, a)
    h = torch.zeros(16, 16, dtype=torch.half)
    with pytest.raises(RuntimeError, match="Tensor size mismatch!"):
        toy_hgemm.hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem(h, torch.zeros(8, 16, dtype=torch.half), h, 2, True, 1)
    with pytest.raises(RuntimeError, match="CUDA device"):  # right dtype/shape, CPU tensor: no CPU path exists
        toy_hgemm.hgemm_cublas_tensor_op_nn(h, h, h)
    q = torch.zeros(1, 1, 128, 48, dtype=torch.half)
    with pytest.raises(RuntimeError, match="headdim not support!"):
        ops.fa2_fwd(q, q, q, q) if False else flash_attn_lib.flash_attn_mma_stages_split_q(torch.zeros(1, 1, 128, 256, dtype=torch.half), q, q, q, 1)
    with pytest.raises(RuntimeError, match="values must be"):
        ops.elementwise_add(torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int32))


def test_block_swizzle_args_are_accepted_and_ignored():
    # hgemm.py passes (stages, swizzle, swizzle_stride) positionally (hgemm.py:L115-135); shape errors must still surface
    import toy_hgemm

    h = torch.zeros(16, 16, dtype=torch.half)
    with pytest.raises(RuntimeError, match="CUDA device"):
        toy_hgemm.hgemm_mma_stages_block_swizzle_tn_cute(h, h, h, 3, True, 2048)
