"""GPU parity of the second set of support kernels against the CPU oracle, through the C ABI:
the seven activations, layer norm, dot product, fp32 transpose, GEMV.  Exact where the operation is a selection or a
permutation (relu, hardshrink, transpose), otherwise within the rounding of the output type (the tolerance is written
next to each check)."""
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
OPS = ["relu", "sigmoid", "gelu", "swish", "elu", "hardswish", "hardshrink"]


def _inputs(n, dtype, seed):
    torch.manual_seed(seed)
    x = torch.randn(n, device="cuda") * 4.0
    # values at the branch points and far outside the clamps
    special = torch.tensor([0.0, -0.0, 0.5, -0.5, 3.0, -3.0, 11.09375, 12.0, 20.0, -9.703125, -12.0, -30.0, 60.0, 100.0,
                            -100.0], device="cuda")
    x[: special.numel()] = special[: min(n, special.numel())]
    return x.to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("op", OPS)
def test_activations_vs_oracle(op, dtype):
    from b200k import ops

    for n, off in ((1 << 16, 0), (100003, 0), (4099, 1), (7, 0)):  # vector body, ragged tail, unaligned view, tiny
        base = _inputs(n + off, dtype, n)
        x = base[off:]
        for clamp in (True, False):
            y = torch.full_like(x, float("nan"))
            ops.activation(x, y, op, ref_clamp=clamp)
            want = oracle.activation(x, op, ref_clamp=clamp)
            if op in ("relu", "hardshrink"):
                assert torch.equal(y.cpu().double(), want), (op, dtype, n)
            elif dtype == torch.float32:
                # ex2.approx / rcp.approx: 2^-22 relative each; gelu multiplies by |x| <= 100
                assert torch.allclose(y.cpu().double(), want, rtol=2e-6, atol=1e-6), (op, n, clamp)
            else:
                # one rounding to fp16 (2^-11 relative) on top of the fp32 evaluation
                assert torch.allclose(y.cpu().double(), want, rtol=1e-3, atol=1e-6), (op, n, clamp)


def test_f16_gelu_saturates_like_the_reference_kernels():
    from b200k import ops

    x = torch.tensor([5.0, 11.0, 12.0, 1000.0], dtype=torch.half, device="cuda")
    y = torch.empty_like(x)
    ops.activation(x, y, "gelu", ref_clamp=True)
    assert y.cpu().tolist() == [5.0, 11.0, 11.09375, 11.09375]
    ops.activation(x, y, "gelu", ref_clamp=False)
    assert y.cpu().tolist() == [5.0, 11.0, 12.0, 1000.0]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(4096, 512), (64, 1024), (33, 8192), (7, 16384), (5, 100), (3, 8), (1, 32768)])
def test_layer_norm_vs_oracle(dtype, shape):
    from b200k import ops

    torch.manual_seed(shape[1])
    x = (torch.randn(*shape, device="cuda") * 2.0 + 0.7).to(dtype)
    for inside in (True, False):
        y = torch.full_like(x, float("nan"))
        ops.layer_norm(x, y, 1.25, -0.5, 1e-5, eps_inside_k=inside)
        want = oracle.layer_norm(x, 1.25, -0.5, 1e-5, eps_inside_k=inside)
        if dtype == torch.float32:
            assert torch.allclose(y.cpu().double(), want, rtol=1e-5, atol=2e-5), (shape, inside)
        else:
            assert torch.allclose(y.cpu().double(), want, rtol=1e-3, atol=2e-3), (shape, inside)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("n", [1, 31, 4096, 1 << 20, (1 << 22) + 5])
def test_dot_product_is_deterministic_and_accurate(dtype, n):
    from b200k import ops

    torch.manual_seed(n)
    a = torch.randn(n + 1, device="cuda").to(dtype)
    b = torch.randn(n + 1, device="cuda").to(dtype)
    for off in (0, 1):  # aligned and unaligned views
        aa, bb = a[off:off + n], b[off:off + n]
        got = ops.dot_prod(aa, bb)
        assert got.dtype == torch.float32 and got.numel() == 1
        scale = float((aa.double().abs() * bb.double().abs()).sum().cpu())
        # fp32 accumulation in a fixed tree of ~n/8 partial sums: error far below 1e-6 of the sum of magnitudes
        assert abs(float(got.cpu()) - oracle.dot_prod(aa, bb)) <= 1e-6 * scale + 1e-30
        assert torch.equal(got, ops.dot_prod(aa, bb))  # same bits every time (no atomics in the sum)


@pytest.mark.parametrize("shape", [(1024, 1024), (4096, 2048), (100, 37), (1, 513), (65, 64), (2048, 8)])
def test_transpose_is_exact(shape):
    from b200k import ops

    torch.manual_seed(shape[0])
    x = torch.randn(*shape, device="cuda")
    y = torch.full((shape[1], shape[0]), float("nan"), device="cuda")
    ops.mat_transpose(x, y)
    assert torch.equal(y.cpu(), oracle.mat_transpose(x))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(1024, 128), (1024, 16), (4096, 4096), (333, 100), (5, 8200), (1, 32)])
def test_gemv_vs_oracle(dtype, shape):
    from b200k import ops

    M, K = shape
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device="cuda").to(dtype)
    x = torch.randn(K, 1, device="cuda").to(dtype)
    y = torch.full((M, 1), float("nan"), device="cuda").to(dtype)
    ops.gemv(a, x, y)
    want = oracle.gemv(a, x)
    mag = (a.double().abs().cpu() @ x.double().abs().cpu())
    if dtype == torch.float32:
        assert ((y.cpu().double() - want).abs() <= 1e-6 * mag + 1e-30).all()
    else:  # fp32 accumulation, one rounding to fp16
        assert ((y.cpu().double() - want).abs() <= 1e-6 * mag + 1e-3 * want.abs() + 1e-7).all()


def test_shim_namespaces_route_to_the_kernels():
    """Each reference entry-point name of the twelve new libraries runs and matches the oracle on a small case."""
    from b200k import support_libs as S

    torch.manual_seed(5)
    x32 = torch.randn(64, 256, device="cuda")
    x16 = x32.half()
    for op in OPS:
        lib = getattr(S, op + "_lib")
        for name, fn in vars(lib).items():
            x = x16 if "f16" in name else x32
            y = torch.empty_like(x)
            fn(x, y)
            tol = dict(rtol=1e-3, atol=1e-6) if x.dtype == torch.half else dict(rtol=2e-6, atol=1e-6)
            assert torch.allclose(y.cpu().double(), oracle.activation(x, op), **tol), name
    for name, fn in vars(S.layer_norm_lib).items():
        x = x32 if name in ("layer_norm_f32", "layer_norm_f32x4") else x16
        y = torch.empty_like(x)
        fn(x, y, 1.0, 0.0)
        tol = dict(rtol=1e-3, atol=2e-3) if x.dtype == torch.half else dict(rtol=1e-5, atol=2e-5)
        assert torch.allclose(y.cpu().double(), oracle.layer_norm(x), **tol), name
    for name, fn in vars(S.dot_product_lib).items():
        x = x16 if "f16" in name else x32
        out = fn(x, x)
        assert abs(float(out.cpu()) - oracle.dot_prod(x, x)) <= 1e-5 * oracle.dot_prod(x, x), name
    for name, fn in vars(S.mat_transpose_lib).items():
        y = torch.empty(256, 64, device="cuda")
        fn(x32, y)
        assert torch.equal(y.cpu(), x32.cpu().t()), name
    for lib, dt in ((S.sgemv_lib, torch.float32), (S.hgemv_lib, torch.float16)):
        for name, fn in vars(lib).items():
            K = 16 if "k16" in name else 256
            a = torch.randn(128, K, device="cuda").to(dt)
            v = torch.randn(K, 1, device="cuda").to(dt)
            y = torch.empty(128, 1, device="cuda").to(dt)
            fn(a, v, y)
            assert torch.allclose(y.cpu().double(), oracle.gemv(a, v), rtol=2e-3, atol=2e-2 if dt == torch.half else 1e-4), name
    with pytest.raises(RuntimeError, match="K must be multiple of 128"):
        S.sgemv_lib.sgemv_k128_f32x4(torch.randn(4, 32, device="cuda"), torch.randn(32, 1, device="cuda"),
                                     torch.empty(4, 1, device="cuda"))


@pytest.mark.parametrize("shape", [(2, 3, 256, 1000), (1, 1, 64, 64), (1, 2, 77, 130), (4, 8, 256, 8192)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_transpose_16bit_batched_bit_exact(shape, dtype):
    from b200k import ops

    torch.manual_seed(sum(shape))
    x = torch.randn(*shape, device="cuda").to(dtype)
    y = torch.empty(*shape[:-2], shape[-1], shape[-2], dtype=dtype, device="cuda")
    ops.transpose_16bit_batched(x, y)
    assert torch.equal(y, x.transpose(-2, -1).contiguous())


def test_transposed_v_entry_point_large_headdim():
    """flash_attn_mma_stages_split_q_tiling_qk_swizzle_qkv with D = 256: V arrives as [B,H,D,N] (flash_attn.cc:L128-159)."""
    from b200k import flash_attn_lib
    from oracle import oracle

    torch.manual_seed(3)
    B, H, N, D = 1, 2, 512, 256
    q, k, v = [torch.randn(B, H, N, D, dtype=torch.half, device="cuda") for _ in range(3)]
    o = torch.zeros_like(q)
    flash_attn_lib.flash_attn_mma_stages_split_q_tiling_qk_swizzle_qkv(q, k, v.transpose(-2, -1).contiguous(), o, 1)
    assert torch.allclose(o.cpu().float(), oracle.attention(q, k, v).float(), rtol=1e-2, atol=1e-3)
