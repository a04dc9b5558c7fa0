"""CPU-only checks of the C-ABI boundary: liblnb.so loads, exports every symbol include/lnb.h
declares (and the ctypes table matches the header), reports errors the documented way, and its
host-side pieces (synthetic generator twin, shard arithmetic) agree with the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import lnb_b200 as L
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lnb.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lnb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = header_functions()
    assert len(names) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", L._capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (lnb_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in lnb.h but not exported: {missing}"
    unbound = [n for n in names if n not in L._capi.SIGNATURES]
    assert not unbound, f"declared in lnb.h but absent from the ctypes table: {unbound}"
    extra = [n for n in L._capi.SIGNATURES if n not in names]
    assert not extra, f"bound but not declared in lnb.h: {extra}"


def test_header_cites_reference_for_every_entry_point():
    src = open(HEADER).read()
    # every op-level / model-level declaration is preceded by a comment citing a reference file:line
    assert len(re.findall(r"[a-z_]+\.go:\d+", src)) >= 20


def test_no_gpu_means_loud_errors():
    lib = L._capi.lib
    if lib.lnb_device_count() > 0:
        pytest.skip("a GPU is present")
    args = L.synth.args_c(dict(L.synth.TINY))
    h = C.c_void_p()
    rc = lib.lnb_model_create(C.byref(args), 0, 0, 1, None, C.byref(h))
    assert rc == -2 and b"failed" in lib.lnb_last_error()           # LNB_ECUDA, never a CPU fallback
    x = np.zeros((1, 64), np.uint16)
    w = np.zeros((16, 64), np.uint16)
    out = np.zeros((1, 16), np.uint16)
    rc = lib.lnb_op_linear_bf16(L._capi.ptr(x, L._capi.u16p), L._capi.ptr(w, L._capi.u16p), L._capi.ptr(out, L._capi.u16p), 1, 64, 16, 0)
    assert rc < 0
    with pytest.raises(L._capi.LnbError):
        L.ml.LinearTransformation(L.ml.Tensor(x, L.ml.DT_BF16), L.ml.Tensor(w, L.ml.DT_BF16))


def test_argument_validation_needs_no_gpu():
    lib = L._capi.lib
    bad = dict(L.synth.TINY, n_heads=7)
    h = C.c_void_p()
    assert lib.lnb_model_create(C.byref(L.synth.args_c(bad)), 0, 0, 1, None, C.byref(h)) == -1
    assert b"n_kv_heads" in lib.lnb_last_error()
    assert lib.lnb_model_create(C.byref(L.synth.args_c(dict(L.synth.TINY))), 0, 2, 2, None, C.byref(h)) == -1
    assert lib.lnb_model_create(C.byref(L.synth.args_c(dict(L.synth.TINY))), 0, 0, 2, None, C.byref(h)) == -1
    assert b"unique id" in lib.lnb_last_error()
    assert lib.lnb_op_linear_bf16(None, None, None, 1, 1, 1, 0) == -1
    assert lib.lnb_forward(None, None, 1, 0, None, 0, None) == -1
    with pytest.raises(L._capi.LnbError, match="unknown tensor name"):
        L.synth.spec(dict(L.synth.TINY), "layers.0.attention.wz.weight")
    with pytest.raises(L._capi.LnbError, match="unknown tensor name"):
        L.synth.spec(dict(L.synth.TINY), "layers.99.attention.wq.weight")


def test_host_generator_twin_equals_oracle_generator():
    args = dict(L.synth.TINY)
    for name in ["tok_embeddings.weight", "norm.weight", "layers.1.attention.wo.weight", "layers.0.feed_forward.w2.weight",
                 "output.weight"]:
        shape = L.synth.tensor_shapes(args)[name]
        sc, off = L.synth.spec(args, name)
        a = L.synth.fill_host(args, name, 99)
        b = O.synth_fill(99, name, sc, off, int(np.prod(shape))).reshape(shape)
        assert np.array_equal(a, b), name
    # tensors of a million elements and more are filled by several threads (counter-based generator): same bits
    for n in (1_048_576, 5_000_001):
        out = np.empty(n, np.uint16)
        L._capi.check(L._capi.lib.lnb_synth_fill_host(77, b"layers.3.feed_forward.w1.weight", 0.027, 0.0, n, L._capi.ptr(out, L._capi.u16p)))
        assert np.array_equal(out, O.synth_fill(77, "layers.3.feed_forward.w1.weight", 0.027, 0.0, n))
    # published scales keep activations O(1): sqrt(3)/sqrt(fan_in), norms 1 +- 0.1
    sc, off = L.synth.spec(dict(L.synth.LLAMA31_8B), "layers.0.attention.wq.weight")
    assert abs(sc - 3 ** 0.5 / 64) < 1e-7 and off == 0.0
    sc, off = L.synth.spec(dict(L.synth.LLAMA31_8B), "layers.5.ffn_norm.weight")
    assert abs(sc - 0.1) < 1e-8 and off == 1.0


def test_tensor_inventory_is_the_reference_291():
    shapes = L.synth.tensor_shapes(dict(L.synth.LLAMA31_8B))
    assert len(shapes) == 291
    assert sum(int(np.prod(s)) for s in shapes.values()) == 8_030_261_248
    assert shapes["layers.31.feed_forward.w2.weight"] == (4096, 14336)
    assert L.model.ModelArgs().FFNDim == 14336 and L.model.ModelArgs().HeadDim == 128   # llamatransformer.go:569-577


@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_shard_windows_partition_every_tensor(tp):
    args = dict(L.synth.LLAMA31_8B, n_layers=1)
    for name, shape in L.synth.tensor_shapes(args).items():
        if len(shape) == 1:
            for r in range(tp):
                assert L.synth.shard_window(args, name, r, tp) == (0, 0, shape[0], 1)
            continue
        cover = np.zeros(shape, np.int8) if shape[0] * shape[1] < 1 << 22 else None
        rows_seen, cols_seen = 0, 0
        for r in range(tp):
            r0, c0, nr, nc = L.synth.shard_window(args, name, r, tp)
            assert r0 + nr <= shape[0] and c0 + nc <= shape[1]
            if name.startswith("tok_embeddings"):
                assert (r0, c0, nr, nc) == (0, 0, shape[0], shape[1])      # replicated
                continue
            rows_seen += nr if nc == shape[1] else 0
            cols_seen += nc if nr == shape[0] else 0
            if cover is not None:
                cover[r0:r0 + nr, c0:c0 + nc] += 1
        if not name.startswith("tok_embeddings"):
            if tp > 1:
                assert (rows_seen == shape[0]) != (cols_seen == shape[1]), name   # split along exactly one axis
            if cover is not None:
                assert (cover == 1).all(), name
    # attention heads stay whole and GQA-aligned: 32/tp q heads with their 8/tp kv heads
    r0, _, nr, _ = L.synth.shard_window(args, "layers.0.attention.wq.weight", tp - 1, tp)
    k0, _, nk, _ = L.synth.shard_window(args, "layers.0.attention.wk.weight", tp - 1, tp)
    assert nr % 128 == 0 and nk % 128 == 0 and (r0 // 128) // 4 == k0 // 128
    with pytest.raises(L._capi.LnbError):
        L.synth.shard_window(dict(L.synth.LLAMA31_8B), "output.weight", 0, 3)


def test_ml_tensor_host_ops_match_reference_semantics():
    ml = L.ml
    t = ml.Tensor.from_f32([1.53, 6.53, 11.34, 586.25])               # src/dtype/bfloat16_test.go:29-66
    assert list(t.to_f32_array()) == [1.5234375, 6.5, 11.3125, 584.0]
    x = ml.Tensor(np.arange(24, dtype=np.int32).reshape(4, 6), ml.DT_INT32)
    assert x.Slice([1], [3]).Size == [2, 6]                             # tensor.go:266-343
    assert x.Slice([3], [4]).Size == [1, 6]
    assert x.Slice([2, 1], [2, 4]).Size == [3] and list(x.Slice([2, 1], [2, 4]).RawData) == [13, 14, 15]
    with pytest.raises(ml.MlError):
        x.Slice([0], [5])
    with pytest.raises(ml.MlError):
        x.Slice([0, 0, 0], [1, 1, 1])
    f = ml.Tensor(np.array([[1.53, -2.0]], np.float32), ml.DT_F32)
    assert f.ToBFloat16().ToFloat32().RawData[0, 0] == np.float32(1.5234375)
    assert ml.Full([3], ml.DT_INT32, -1).RawData.tolist() == [-1, -1, -1]


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu():
    """the C++ host mirror (host/lnb_host.hpp) compiles against include/lnb.h and links liblnb.so"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    exe = os.path.join(ROOT, "host", "lnb_generate")
    assert os.path.exists(exe)
    if L._capi.lib.lnb_device_count() > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([exe, "16", "strict", "tiny"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "lnb:" in out.stderr          # no CPU fallback


def test_ml_builders_match_the_reference_literals_and_build_the_rope_table():
    """ml.ARange / Outer / Polar / Ones / TriangularUpper of the host mirror against tests/golden (the reference's own
    test literals), and model.precomputeFreqsCis composed from them against the oracle's table, bit for bit"""
    import math
    from tests.helpers import reference_vectors
    g = reference_vectors()
    ml = L.ml
    for c in g["arange_bf16"]:
        t = ml.ARange(*c["args"], ml.DT_BF16)
        assert t.Size == [len(c["expected"])] and np.array_equal(t.to_f32_array(), np.array(c["expected"], np.float32))
    with pytest.raises(ml.MlError, match="must be less than end value"):
        ml.ARange(5, 5, 1, ml.DT_BF16)
    o = g["outer_bf16"]
    got = ml.Outer(ml.ARange(*o["vec1_arange"], ml.DT_BF16), ml.ARange(*o["vec2_arange"], ml.DT_BF16))
    assert got.Size == [4, 3] and np.array_equal(got.to_f32_array(), np.array(o["expected"], np.float32))
    p = g["polar"]
    angle = np.array([[eval(e.replace("math.Pi", "math.pi"), {"math": math}) for e in p["angle_expr"]]], np.float32)
    got = ml.Polar(ml.Tensor(np.array([p["abs"]], np.float32), ml.DT_F32), ml.Tensor(angle, ml.DT_F32))
    assert got.DataType is ml.DT_COMPLEX and np.abs(got.RawData - np.array([[complex(a, b) for a, b in p["expected_re_im"]]])).max() <= 1e-3
    for c in g["triangular_upper"]:
        got = ml.TriangularUpper(ml.Full(c["size"], ml.DT_F32, np.float32(c["fill"])), c["diagonal"])
        assert np.array_equal(got.RawData, np.array(c["expected"], np.float32))
    assert ml.OnesLike(ml.Zeros([2, 3], ml.DT_BF16)).to_f32_array().tolist() == [[1.0] * 3] * 2
    for dim, end, theta, scaled in ((128, 512, 500000.0, True), (32, 128, 500000.0, True), (64, 64, 10000.0, False)):
        cis = L.model.precomputeFreqsCis(dim, end, theta, scaled)
        _, exp = O.rope_table(dim, end, theta, scaled)
        assert cis.Size == [end, dim // 2]
        assert np.array_equal(np.stack([cis.RawData.real, cis.RawData.imag], -1), exp)


def test_tensor_transpose_and_set_slice_match_the_reference_literals():
    """Tensor.Transpose / SetSlice / Slice of the host mirror on the reference's own cases (src/ml/tensor_test.go:171-479):
    the layout changes of attention and the KV-cache append are exactly these on the host side of the Go path"""
    from tests.helpers import reference_vectors
    g = reference_vectors()
    ml = L.ml
    for c in g["transpose"]:
        t = ml.Tensor(np.array(c["input"], np.float32), ml.DT_F32).Transpose(*c["dims"])
        exp = np.array(c["expected"], np.float32)
        assert t.Size == list(exp.shape) and np.array_equal(t.RawData, exp) and t.RawData.flags["C_CONTIGUOUS"]
    exp = np.array(g["set_slice"]["expected"], np.float32)
    inp = ml.Tensor.from_f32(np.arange(1, 21, dtype=np.float32).reshape(4, 5))          # createTestInputTensor([4, 5])
    a = ml.Zeros([10, 5], ml.DT_BF16)
    a.SetSlice([1], [5], inp)
    assert np.array_equal(a.Slice([1], [5]).to_f32_array(), exp) and not a.RawData[0].any() and not a.RawData[5:].any()
    b = ml.Zeros([20, 10, 5], ml.DT_BF16)
    b.SetSlice([19, 1], [19, 5], inp)
    assert np.array_equal(b.Slice([19, 1], [19, 5]).to_f32_array(), exp) and not b.RawData[:19].any()
    with pytest.raises(ml.MlError):
        a.SetSlice([1], [4], inp)
    with pytest.raises(ml.MlError):
        a.SetSlice([1], [5], ml.Tensor(np.zeros((4, 5), np.float32), ml.DT_F32))
