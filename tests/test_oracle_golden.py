"""Pins the CPU oracle to the reference's own weight-free golden vectors.

Every expected value below is a known answer held by the reference's tests or docs
(cited per test); tolerances are the reference's own (src/common/utils.go:13-17:
THRESHOLD_F32 = 1e-3, THRESHOLD_EXACT = 0).
"""
import math

import numpy as np
import pytest

from oracle import oracle as O

THRESHOLD_F32 = 1e-3


def bf(x):
    return O.bf16_bits(np.asarray(x, np.float32))


def f(b):
    return O.bf16_to_f32(b)


# ---- src/dtype/bfloat16_test.go:8-66 -------------------------------------------------

@pytest.mark.parametrize("inp,exp", [(6.25, 6.25), (1.53, 1.5234375), (6.53, 6.5), (11.34, 11.3125), (586.25, 584.0)])
def test_bf16_truncation(inp, exp):
    b = O.lib().orc_f32_to_bf16(inp)
    assert O.lib().orc_bf16_to_f32(b) == np.float32(exp)
    assert f(bf([inp]))[0] == np.float32(exp)


# src/dtype/bfloat16_test.go:68-106 (little-endian storage)
@pytest.mark.parametrize("raw,bits,val", [
    (b"\xA5\x35", 0x35A5, 0.0000012293458), (b"\xF4\xB5", 0xB5F4, -0.00000181794167), (b"\x92\xB6", 0xB692, -0.000004351139)])
def test_bf16_little_endian(raw, bits, val):
    b = np.frombuffer(raw, dtype="<u2")
    assert int(b[0]) == bits
    assert f(b)[0] == np.float32(val)


# ---- src/ml/operations_test.go:782-829 ------------------------------------------------

def test_linear_f32_golden():
    w = np.array([[0.01, 0.02, 0.03], [0.04, 0.05, 0.06], [0.07, 0.08, 0.09], [0.10, 0.11, 0.12]], np.float32)
    x = np.array([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]], np.float32)
    exp = np.array([[0.014, 0.032, 0.05, 0.068], [0.032, 0.077, 0.122, 0.167]], np.float32)
    assert np.abs(O.linear_f32(x, w) - exp).max() <= THRESHOLD_F32


# ---- src/ml/operations_test.go:831-878 ------------------------------------------------

def test_linear_bf16_golden():
    w = bf([[0.01, 0.02, 0.03], [0.04, 0.05, 0.06], [0.07, 0.08, 0.09], [0.10, 0.11, 0.12]])
    x = bf([[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]])
    exp = np.array([[0.0138, 0.0317, 0.0495, 0.0673], [0.0317, 0.0761, 0.1210, 0.1660]], np.float32)
    got = f(O.linear_bf16(x, w))
    assert np.abs(got - exp).max() <= THRESHOLD_F32
    # (the literals are 4-decimal PyTorch prints, so 1e-3 is as tight as the reference pins this)
    assert np.all(np.abs(got - exp) < 1e-4)


# ---- src/ml/operations_test.go:880-946 ------------------------------------------------

def test_matmul_bf16_golden():
    a = bf([[[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]]] * 2)
    b = bf([[[0.01, 0.02, 0.03, 0.04], [0.05, 0.06, 0.07, 0.08], [0.09, 0.10, 0.11, 0.12]]] * 2)
    exp = np.array([[[3.7598e-02, 4.3457e-02, 4.9561e-02, 5.5420e-02], [8.2520e-02, 9.7168e-02, 1.1230e-01, 1.2695e-01]]] * 2,
                   np.float32)
    got = f(O.matmul_bf16(a, b))
    assert got.shape == (2, 2, 4)
    assert np.abs(got - exp).max() <= THRESHOLD_F32
    assert np.all(np.abs(got - exp) / exp < 1e-4)  # 5 significant digits printed


# ---- src/ml/operations_test.go:589-652 (Pow) ------------------------------------------

def test_pow2_golden():
    x = bf(np.arange(3, 8, dtype=np.float32))
    assert np.array_equal(O.pow2_bf16(x), np.array([9, 16, 25, 36, 49], np.float32))


# ---- src/ml/operations_test.go:654-780 (Mean over createTestInputTensor 1,2,3,...) -----

def test_mean_golden():
    t = np.arange(1, 61, dtype=np.float32).reshape(5, 4, 3)
    got = O.mean_f32(t)
    exp = t.mean(-1, keepdims=True)  # 2, 5, 8, ... exactly representable
    assert got.shape == (5, 4, 1)
    assert np.array_equal(got, exp)
    assert got[0, 0, 0] == 2.0 and got[4, 3, 0] == 59.0


# ---- docs/10-ROPE-ROTARY-POSITIONAL-EMBEDDINGS.md:276-288 (all 64 scaled inverse freqs) --

DOC_FREQS = [
    1.0000e+00, 8.1250e-01, 6.6016e-01, 5.3906e-01, 4.3945e-01, 3.5742e-01, 2.9102e-01, 2.3730e-01, 1.9336e-01, 1.5723e-01,
    1.2793e-01, 1.0449e-01, 8.4961e-02, 6.9336e-02, 5.6641e-02, 4.6143e-02, 3.7598e-02, 3.0518e-02, 2.4902e-02, 2.0264e-02,
    1.6479e-02, 1.3489e-02, 1.0986e-02, 8.9111e-03, 7.2632e-03, 5.9204e-03, 4.8218e-03, 3.9368e-03, 3.2043e-03, 2.1515e-03,
    1.3504e-03, 8.5068e-04, 5.1880e-04, 3.1090e-04, 1.7834e-04, 9.5367e-05, 7.7724e-05, 6.2943e-05, 5.1498e-05, 4.1962e-05,
    3.4094e-05, 2.7895e-05, 2.2650e-05, 1.8477e-05, 1.5080e-05, 1.2279e-05, 1.0014e-05, 8.1062e-06, 6.6459e-06, 5.3942e-06,
    4.4107e-06, 3.5912e-06, 2.9206e-06, 2.3842e-06, 1.9372e-06, 1.5795e-06, 1.2890e-06, 1.0431e-06, 8.5309e-07, 6.9663e-07,
    5.6624e-07, 4.6194e-07, 3.7625e-07, 3.0547e-07]


def test_rope_freqs_match_reference_doc():
    freqs, _ = O.rope_table(128, 8, 500000.0, True)
    got = f(freqs)
    for g, e in zip(got, DOC_FREQS):
        # the doc prints 5 significant digits of the bf16 value
        assert float("%.4e" % g) == pytest.approx(e, rel=0, abs=0), (g, e)


def test_rope_positions_are_bf16_quantised():
    # docs/10-ROPE...md:397,403: t[4094] and t[4095] both give 4.0800e+03 * freq
    _, cis = O.rope_table(128, 4096, 500000.0, True)
    assert np.array_equal(cis[4094], cis[4095])
    assert np.array_equal(cis[257], cis[256])          # bf16(257) == 256
    assert not np.array_equal(cis[255], cis[256])
    # freq[0] == 1 -> angle == bf16(pos): row 4094 col 0 is cos/sin(4080)
    assert cis[4094, 0, 0] == np.float32(math.cos(4080.0)) and cis[4094, 0, 1] == np.float32(math.sin(4080.0))
    assert np.array_equal(cis[0], np.stack([np.ones(64, np.float32), np.zeros(64, np.float32)], -1))


# ---- own known-answer tests for ops the reference never pins (SURVEY 8c last row) -------

def test_argmax_first_max_wins_and_nan():
    x = np.array([1, 5, 5, 2], np.float32)
    assert O.argmax_f32(x) == 1
    x = np.array([np.nan, -1, -1], np.float32)
    assert O.argmax_f32(x) == 1
    assert O.argmax_f32(np.array([np.nan], np.float32)) == -1


def test_softmax_f64_no_max_subtraction():
    x = np.array([[0.5, -1.25, 3.0]], np.float32)
    e = np.exp(x.astype(np.float64))
    exp = (e / e.sum()).astype(np.float32)
    assert np.array_equal(O.softmax_f32(x), exp)


def test_silu_table_known_values():
    tab = O.silu_table_bf16()
    assert tab[0] == 0                                   # silu(0) = 0
    one = int(bf([1.0])[0])
    assert tab[one] == bf([np.float32(1.0 / (1.0 + math.exp(-1.0)))])[0]
    ninf = int(bf([-np.inf])[0])
    assert np.isnan(f(tab[ninf:ninf + 1])[0])            # -inf/(1+inf) = NaN, like the Go table
    x = bf([0.25, -3.0, 7.5])
    got = O.silu_bf16(x)
    for b, g in zip(x, got):
        v = float(f(np.array([b]))[0])
        assert g == bf([np.float32(v / (1.0 + math.exp(-v)))])[0]


def test_rmsnorm_two_truncations():
    rng = np.random.default_rng(1)
    x = bf(rng.standard_normal((3, 64)))
    w = bf(1 + 0.1 * rng.standard_normal(64))
    xs = f(x)
    out = np.empty_like(x)
    for s in range(3):
        acc = np.float32(0)
        for j in range(64):
            acc = np.float32(acc + np.float32(xs[s, j] * xs[s, j]))
        ms = np.float32(acc / np.float32(64)) + np.float32(1e-5)
        r = np.float32(1.0 / math.sqrt(float(ms)))
        n1 = bf(xs[s] * r)
        out[s] = bf(f(n1) * f(w))
    assert np.array_equal(O.rmsnorm(x, w, 1e-5), out)


def test_rope_apply_uses_f64_intermediates():
    rng = np.random.default_rng(2)
    _, cis = O.rope_table(128, 64, 500000.0, True)
    x = bf(rng.standard_normal((2, 3, 128)))
    got = O.rope_apply(x, cis, 5)
    xs = f(x).astype(np.float64)
    a, b = xs[..., 0::2], xs[..., 1::2]
    c = cis[5:7, None, :, 0].astype(np.float64)
    d = cis[5:7, None, :, 1].astype(np.float64)
    re = (a * c - b * d).astype(np.float32)
    im = (a * d + b * c).astype(np.float32)
    exp = np.empty_like(x)
    exp[..., 0::2] = bf(re)
    exp[..., 1::2] = bf(im)
    assert np.array_equal(got, exp)


def test_linear_paths_agree_s1_vs_sN():
    rng = np.random.default_rng(3)
    x = bf(rng.standard_normal((5, 96)))
    w = bf(rng.standard_normal((37, 96)) * 0.1)
    full = O.linear_bf16(x, w)
    for s in range(5):
        assert np.array_equal(O.linear_bf16(x[s:s + 1], w), full[s:s + 1])
    # and against a literal sequential-f32 python loop
    xs, ws = f(x), f(w)
    for s, n in [(0, 0), (4, 36), (2, 17)]:
        acc = np.float32(0)
        for k in range(96):
            acc = np.float32(acc + np.float32(xs[s, k] * ws[n, k]))
        assert full[s, n] == bf([acc])[0]
