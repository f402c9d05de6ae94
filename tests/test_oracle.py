"""CPU: the oracle (oracle/oracle.py) against the reference's own known-answer fixtures, closed forms and the frozen
golden vectors.  These pin the checker that the GPU parity tests rely on."""
import math

import numpy as np
import torch

from oracle import oracle


def test_all_ones_qkv_gives_ones(golden):
    # reference fixture --no-rand-qkv: all-ones Q/K/V -> O == 1 exactly (flash_attn_mma.py:L353-369)
    g = golden("kat_attention_all_ones.npz")
    B, H, N, D = [int(x) for x in g["shape"]]
    ones = torch.ones(B, H, N, D, dtype=torch.half)
    for fn in (oracle.attention, lambda q, k, v: oracle.attention_tiled(q, k, v, Bc=64),
               lambda q, k, v: oracle.attention_tiled(q, k, v, Bc=128, pv_acc_f16=True, o_store_f16=True)):
        o = fn(ones, ones, ones)
        assert torch.equal(o, torch.ones_like(o))
    assert np.array_equal(g["o"], np.ones((B, H, N, D), dtype=np.float16))


def test_range_k_fixture_closed_form(golden):
    # reference fixture --range-k: K[:, :, j, :] = (j+1)/N  =>  s_ij = c_i * (j+1)/N with c_i = sum_d(q_id)/sqrt(D)
    g = golden("kat_attention_range_k.npz")
    q, k, v = [torch.from_numpy(g[n]) for n in ("q", "k", "v")]
    B, H, N, D = q.shape
    assert torch.equal(k, oracle.make_range_k(B, H, N, D))
    c = q.double().sum(-1, keepdim=True) / math.sqrt(D)                       # [B,H,N,1]
    kj = k.double()[:, :, :, 0].unsqueeze(-2)                                 # [B,H,1,N] (fp16-rounded (j+1)/N)
    w = torch.softmax(c * kj, dim=-1)                                         # [B,H,N,N]
    o_closed = (w @ v.double()).half()
    o = oracle.attention(q, k, v)
    assert torch.allclose(o.float(), o_closed.float(), rtol=1e-3, atol=1e-3)
    assert np.array_equal(o.numpy(), g["o"])                                  # frozen vector


def test_tiled_restatement_matches_plain():
    torch.manual_seed(3)
    q, k, v = [torch.randn(1, 2, 192, 64).half() for _ in range(3)]
    ref = oracle.attention(q, k, v).float()
    for kw in (dict(Bc=64), dict(Bc=128), dict(Bc=64, pv_acc_f16=True), dict(Bc=128, o_store_f16=True)):
        got = oracle.attention_tiled(q, k, v, **kw).float()
        assert torch.allclose(got, ref, rtol=1e-2, atol=2e-3), kw


def test_histogram_fixture(golden):
    g = golden("kat_histogram.npz")  # histogram.py:L22-31: list(range(10))*1000 -> 1000 per bin
    assert g["a"].tolist() == list(range(10)) * 1000
    assert oracle.histogram(torch.from_numpy(g["a"])).tolist() == [1000] * 10 == g["hist"].tolist()


def test_hgemm_oracle_vs_float64(golden):
    g = golden("seeded_hgemm.npz")
    a, b = torch.from_numpy(g["a"]), torch.from_numpy(g["b"])
    exact = (a.double() @ b.double())
    c = oracle.hgemm(a, b)
    assert np.array_equal(c.numpy(), g["c"])
    # one fp16 rounding of the exact product: relative 2^-11
    assert torch.allclose(c.double(), exact, rtol=2 ** -10, atol=1e-4)
    # reference-faithful fp16-accumulate variant stays within its own error budget (K/16 roundings)
    c16 = oracle.hgemm_f16acc_k16(a[:, :64], b[:64])
    assert np.array_equal(c16.numpy(), g["c_f16acc"])
    ex64 = a[:, :64].double() @ b[:64].double()
    assert (c16.double() - ex64).abs().max() < 4 * 2 ** -10 * ex64.abs().max() + 1e-2


def test_frozen_attention_vectors(golden):
    for D in (32, 64, 96, 128, 256, 320):
        g = golden("seeded_attention_d%d.npz" % D)
        q, k, v = [torch.from_numpy(g[n]) for n in ("q", "k", "v")]
        assert np.array_equal(oracle.attention(q, k, v).numpy(), g["o"])


def test_row_kernels_oracle(golden):
    g = golden("seeded_rows.npz")
    x = torch.from_numpy(g["x"])
    sm = oracle.softmax_per_token(x)
    assert torch.allclose(sm.sum(-1), torch.ones(x.size(0)), atol=1e-5)
    assert np.allclose(sm.numpy(), g["softmax"], atol=1e-7)
    assert abs(float(oracle.softmax_all(x).double().sum()) - 1.0) < 1e-5
    r = oracle.rms_norm(x, 1.0)
    assert torch.allclose(r.pow(2).mean(-1), torch.ones(x.size(0)), atol=1e-3)
    # the reference's f16 kernels put eps inside K: rsqrt(sum/(K+eps)); difference is O(eps/K)
    r2 = oracle.rms_norm(x, 1.0, eps_inside_k=True)
    assert (r - r2).abs().max() < 1e-4 and not torch.equal(r, r2)
    assert abs(oracle.reduce_sum(x) - float(g["sum"])) < 1e-9


def test_rope_quirk_restates_reference_integer_division(golden):
    g = golden("seeded_rows.npz")
    x = torch.from_numpy(g["x"])
    M, Hd = x.shape
    N = Hd // 2
    # rope.cu:L26: exp_v = 1/powf(theta, token_idx / (N*2)) with INTEGER division => exponent 0 for every pair
    assert all((t // (N * 2)) == 0 for t in range(N))
    q = oracle.rope(x, True)
    t = oracle.rope(x, False)
    assert torch.allclose(q[0], x[0]) and torch.allclose(t[0], x[0])          # position 0: angle 0 in both
    assert torch.allclose(q[:, 0:2], t[:, 0:2], atol=1e-5)                    # pair 0 has frequency 1 in both
    assert not torch.allclose(q[5:], t[5:], atol=1e-3)                        # other pairs differ
    # rotation preserves each pair's norm
    assert torch.allclose(q[:, 0::2] ** 2 + q[:, 1::2] ** 2, x[:, 0::2] ** 2 + x[:, 1::2] ** 2, rtol=1e-4, atol=1e-5)


def test_elementwise_and_embedding_exact():
    torch.manual_seed(0)
    a, b = torch.randn(1000).half(), torch.randn(1000).half()
    c = oracle.elementwise_add(a, b)
    # correctly rounded half sum: compare against float64 sum rounded once
    assert torch.equal(c, (a.double() + b.double()).half())
    w = torch.randn(50, 8)
    idx = torch.randint(0, 50, (20,), dtype=torch.int32)
    assert torch.equal(oracle.embedding(idx, w), w[idx.long()])


def test_second_set_oracles_match_the_comparators_the_reference_scripts_print():
    """The reference's scripts for these kernels print their outputs next to a torch comparator (no assertion):
    torch.nn.GELU("tanh") (gelu.py:L62), torch.sigmoid (sigmoid.py:L70), x * sigmoid(x) (swish.py:L59),
    where(x > 0, x, exp(x) - 1) (elu.py:L50), F.hardswish / F.hardshrink(lambd=0.5) (hardswish.py:L51, hardshrink.py:L51),
    torch.dot, torch.matmul, x.t().  The oracle's plain mode (no reference clamp) has to agree with them."""
    import torch.nn.functional as F

    torch.manual_seed(11)
    x = (torch.randn(4096) * 4).double()
    pairs = {
        "relu": torch.relu(x), "sigmoid": torch.sigmoid(x), "gelu": F.gelu(x, approximate="tanh"), "swish": x * torch.sigmoid(x),
        "elu": torch.where(x > 0, x, torch.exp(x) - 1), "hardswish": F.hardswish(x), "hardshrink": F.hardshrink(x, lambd=0.5),
    }
    for op, want in pairs.items():
        assert torch.allclose(oracle.activation(x, op, ref_clamp=False), want, rtol=1e-12, atol=1e-14), op
    # the reference clamp only changes values outside the clamp range
    inside = x.abs() < 9.0
    for op in ("sigmoid", "gelu"):
        a, b = oracle.activation(x.half(), op, ref_clamp=True), oracle.activation(x.half(), op, ref_clamp=False)
        assert torch.equal(a[inside], b[inside])
    assert float(oracle.activation(torch.tensor([20.0]).half(), "gelu")[0]) == 11.09375
    assert float(oracle.activation(torch.tensor([20.0]), "gelu")[0]) == 20.0  # the f32 clamp is at 88.4

    m = torch.randn(37, 300).double()
    assert torch.allclose(oracle.layer_norm(m, 1.0, 0.0, 1e-5, eps_inside_k=False), F.layer_norm(m, (300,), eps=1e-5),
                          rtol=1e-10, atol=1e-12)
    # the reference's form (eps added to K) differs from the textbook one by a relative 1e-5/K, i.e. not at all in fp32
    assert torch.allclose(oracle.layer_norm(m, 1.5, -0.5, 1e-5, True), oracle.layer_norm(m, 1.5, -0.5, 0.0, False), rtol=1e-6)
    a, b = torch.randn(1000).double(), torch.randn(1000).double()
    assert abs(oracle.dot_prod(a, b) - float(torch.dot(a, b))) < 1e-10
    assert torch.equal(oracle.mat_transpose(m), m.t().contiguous())
    v = torch.randn(300, 1).double()
    assert torch.allclose(oracle.gemv(m, v), m @ v)
