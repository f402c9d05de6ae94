"""GPU parity of the support kernels (through the C ABI and the drop-in namespaces) vs the CPU oracle.
Bit-exact for integer / copy / add kernels; tolerance written per assert for floating point."""
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 7, 4096, 1000003])
def test_elementwise_add_bit_exact(dtype, n):
    from b200k import ops

    torch.manual_seed(n)
    a = torch.randn(n, device="cuda").to(dtype)
    b = torch.randn(n, device="cuda").to(dtype)
    c = torch.empty_like(a)
    ops.elementwise_add(a, b, c)
    assert torch.equal(c.cpu(), oracle.elementwise_add(a, b))
    if n > 8:  # unaligned views take the scalar path
        ops.elementwise_add(a[1:], b[1:], c[1:])
        assert torch.equal(c[1:].cpu(), oracle.elementwise_add(a[1:], b[1:]))


def test_reduce_all_variants(ref_exports):
    from b200k import support_libs

    torch.manual_seed(0)
    x32 = torch.randn(1024, 1024, device="cuda")
    for name in ref_exports["block_all_reduce_lib"]:
        fn = getattr(support_libs.reduce_lib, name)
        if "fp8_e4m3" in name:
            x = (x32 * 0.5).to(torch.float8_e4m3fn)
        elif "fp8_e5m2" in name:
            x = (x32 * 0.5).to(torch.float8_e5m2)
        elif name.startswith("block_all_reduce_sum_i8"):
            x = torch.randint(-128, 128, (1024, 1024), dtype=torch.int8, device="cuda")
        elif "bf16" in name:
            x = x32.bfloat16()
        elif "f16" in name.split("sum_")[1].split("_")[0]:
            x = x32.half()
        else:
            x = x32
        y = fn(x)
        want = oracle.reduce_sum(x)
        if x.dtype == torch.int8:
            assert y.dtype == torch.int32 and int(y.item()) == want, name  # exact
        else:
            # half-precision pack sums (the *_f16 / *_bf16 acc variants) carry 2^-11 / 2^-8 relative error per pack
            tol = 2.0 if name.endswith("_f32") else 40.0
            assert abs(y.item() - want) < tol, (name, y.item(), want)
        assert fn(x).item() == y.item(), name + " not deterministic"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(4096, 256), (512, 1024), (300, 4096), (64, 8192), (16, 16384), (333, 1000), (5, 77)])
def test_softmax_and_rmsnorm_rows(dtype, shape):
    from b200k import ops

    torch.manual_seed(shape[1])
    x = torch.randn(*shape, device="cuda").to(dtype)
    y = torch.empty_like(x)
    atol = 1e-6 if dtype == torch.float32 else 1e-3
    want = oracle.softmax_per_token(x)
    for mode in (ops.SOFTMAX_PER_TOKEN, ops.SOFTMAX_SAFE, ops.SOFTMAX_ONLINE):
        ops.softmax(x, y, mode)
        assert torch.allclose(y.cpu().float(), want, rtol=1e-3, atol=atol), mode
    ops.rms_norm(x, y, 1.5)
    assert torch.allclose(y.cpu().float(), oracle.rms_norm(x, 1.5), rtol=2e-3, atol=atol * 10)
    if dtype == torch.float16:
        ops.rms_norm(x, y, 1.0, eps_inside_k=True)  # the reference's f16 kernels
        assert torch.allclose(y.cpu().float(), oracle.rms_norm(x, 1.0, eps_inside_k=True), rtol=2e-3, atol=2e-3)


def test_softmax_whole_tensor_mode():
    from b200k import support_libs

    torch.manual_seed(1)
    x = torch.randn(4096, 64, device="cuda")
    y = torch.empty_like(x)
    support_libs.softmax_lib.softmax_f32(x, y)
    assert torch.allclose(y.cpu(), oracle.softmax_all(x), rtol=1e-3, atol=1e-10)


@pytest.mark.parametrize("shape", [(4096, 512), (8192, 1024), (100, 6), (33, 10)])
def test_rope_textbook_and_reference_quirk(shape):
    from b200k import ops, support_libs

    torch.manual_seed(3)
    x = torch.randn(*shape, device="cuda")
    out = torch.empty_like(x)
    ops.rope_f32(x, out, ref_quirk=False)
    # angles reach seq_len radians and are formed in fp32 (as in the script's naive_rope): argument error up to
    # seq_len * 2^-23 rad, times |x| <= ~5
    atol = max(2e-3, shape[0] * 2.0 ** -23 * 6)
    assert torch.allclose(out.cpu(), oracle.rope(x, False), rtol=1e-3, atol=atol)
    for name in ("rope_f32", "rope_f32_v2", "rope_f32x4_pack"):  # drop-in names = what the reference kernels compute
        getattr(support_libs.rope_lib, name)(x, out)
        assert torch.allclose(out.cpu(), oracle.rope(x, True), rtol=1e-3, atol=atol), name


def test_histogram_bit_exact(golden):
    from b200k import ops, support_libs

    g = golden("kat_histogram.npz")
    a = torch.from_numpy(g["a"]).cuda()
    assert support_libs.hist_lib.histogram_i32(a).cpu().numpy().tolist() == g["hist"].tolist()
    for hi, n in ((256, 1_000_003), (50_000, 2_000_000), (3, 10)):
        torch.manual_seed(n)
        a = torch.randint(0, hi, (n,), dtype=torch.int32, device="cuda")
        assert ops.histogram_i32(a).cpu().numpy().tolist() == oracle.histogram(a).tolist()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(4096, 1024, 1024), (2048, 1024, 512), (1000, 77, 100), (3, 5, 7)])
def test_embedding_bit_exact(dtype, shape):
    from b200k import ops

    n, rows, emb = shape
    torch.manual_seed(n)
    w = torch.randn(rows, emb, device="cuda").to(dtype)
    idx = torch.randint(0, rows, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty(n, emb, dtype=dtype, device="cuda")
    ops.embedding(idx, w, out)
    assert torch.equal(out.cpu(), oracle.embedding(idx, w))
