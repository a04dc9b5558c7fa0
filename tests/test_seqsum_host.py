"""CPU proof of the binade-scan evaluation of the reference's sequential fp32 sum (csrc/seqsum.cuh).

tests/native/seqsum_host.cpp models the control flow of rms_scale_scan_kernel (iterative) and rms_scale_seg_kernel
(one pass: predict the binades, fold runs, walk) on the SAME seq_term / seq_compose / seq_try_jump code the
kernels compile; here it is compared bit for bit with the one-accumulator loop (ml.Mean's order,
src/model/llamatransformer.go:641-656) on inputs chosen to break it: exact rounding ties, huge dynamic range,
zeros, subnormal sums, overflow to inf.  The GPU twin is tests/test_gpu_ops.py::test_rms_scale_*."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "native", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "seqsum_host.so")
    src = os.path.join(HERE, "native", "seqsum_host.cpp")
    hdr = os.path.join(ROOT, "llama-nuts-and-bolts_b200", "csrc", "seqsum.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.dirname(hdr),
                               src, "-o", so])
    l = C.CDLL(so)
    l.seqsum_reference.restype = C.c_float
    l.seqsum_reference.argtypes = [C.c_void_p, C.c_int]
    l.seqsum_scan_model.restype = C.c_float
    l.seqsum_scan_model.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    l.seqsum_seg_model.restype = C.c_float
    l.seqsum_seg_model.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return l


def both(lib, terms, nt, ch):
    terms = np.ascontiguousarray(terms, np.float32)
    assert terms.size == nt * ch
    it = C.c_int(0)
    a = np.float32(lib.seqsum_reference(terms.ctypes.data, terms.size))
    b = np.float32(lib.seqsum_scan_model(terms.ctypes.data, nt, ch, C.byref(it)))
    return a, b, it.value


def bf16_squares(x):
    xb = (np.asarray(x, np.float32).view(np.uint32) & 0xFFFF0000).view(np.float32)
    return (xb.astype(np.float64) ** 2).astype(np.float32)   # exact: 8-bit x 8-bit significands


SHAPES = [(512, 8), (32, 8), (64, 2), (512, 16), (32, 4)]


def adversarial_rows(rng, n, kind):
    if kind == 0:
        return rng.standard_normal(n)
    if kind == 1:
        return rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))                 # 17 decades of range
    if kind == 2:
        return np.where(rng.random(n) < 0.9, 0.0, rng.standard_normal(n))                # mostly zeros
    if kind == 3:
        return rng.integers(1, 256, n).astype(np.float32) * 2.0 ** rng.integers(-8, 8)   # exact ties everywhere
    if kind == 4:
        return np.full(n, rng.uniform(0.1, 3.0))
    if kind == 5:
        return rng.standard_normal(n) * np.linspace(1e-6, 1e3, n)
    return rng.standard_normal(n) * np.linspace(1e3, 1e-6, n)


def test_scan_equals_sequential_sum_on_bf16_squares(lib):
    rng = np.random.default_rng(1)
    for trial in range(700):
        nt, ch = SHAPES[trial % 5]
        sq = bf16_squares(adversarial_rows(rng, nt * ch, trial % 7))
        a, b, _ = both(lib, sq, nt, ch)
        assert a.view(np.uint32) == b.view(np.uint32), (trial, a, b)


def test_scan_equals_sequential_sum_on_general_terms(lib):
    rng = np.random.default_rng(2)
    for trial in range(300):
        nt, ch = SHAPES[trial % 3]
        n = nt * ch
        kind = trial % 5
        t = np.abs(rng.standard_normal(n)).astype(np.float32)        # full 24-bit significands
        if kind == 1:
            t = (t * 1e-42).astype(np.float32)                        # the sum stays subnormal
        elif kind == 2:
            t = np.zeros(n, np.float32)
        elif kind == 3:
            t[rng.integers(0, n)] = np.inf
        elif kind == 4:
            t = (t * 3e37).astype(np.float32)                         # overflows to +inf on the way
        a, b, _ = both(lib, t, nt, ch)
        assert a.view(np.uint32) == b.view(np.uint32), (trial, kind, a, b)


def test_scan_needs_few_round_trips_on_activation_like_rows(lib):
    """the point of the exercise: ~log2(chunks) trips through the scan loop, not one dependent add per element"""
    rng = np.random.default_rng(3)
    trips = []
    for _ in range(50):
        a, b, it = both(lib, bf16_squares(rng.standard_normal(4096)), 512, 8)
        assert a.view(np.uint32) == b.view(np.uint32)
        trips.append(it)
    assert max(trips) <= 24 and np.mean(trips) <= 16


def seg(lib, terms, nt, ch, sabotage=0):
    terms = np.ascontiguousarray(terms, np.float32)
    j, w = C.c_int(0), C.c_int(0)
    b = np.float32(lib.seqsum_seg_model(terms.ctypes.data, nt, ch, sabotage, C.byref(j), C.byref(w)))
    a = np.float32(lib.seqsum_reference(terms.ctypes.data, terms.size))
    return a, b, j.value, w.value


def test_one_pass_equals_sequential_sum_even_with_wrong_predictions(lib):
    """rms_scale_seg_kernel's model: exact with its own binade predictions AND with corrupted ones (sabotage):
    the walk's validity test, not the prediction, is what guarantees the bits"""
    rng = np.random.default_rng(4)
    for trial in range(900):
        nt, ch = SHAPES[trial % 5]
        n = nt * ch
        if trial % 3 == 2:
            t = np.abs(rng.standard_normal(n)).astype(np.float32)
            k = trial % 15
            if k == 2:
                t = (t * 1e-42).astype(np.float32)
            elif k == 5:
                t[:] = 0
            elif k == 8:
                t[rng.integers(0, n)] = np.inf
            elif k == 11:
                t = (t * 3e37).astype(np.float32)
        else:
            t = bf16_squares(adversarial_rows(rng, n, trial % 7))
        for sab in (0, trial + 1):
            a, b, _, _ = seg(lib, t, nt, ch, sab)
            assert a.view(np.uint32) == b.view(np.uint32), (trial, sab, a, b)


def test_one_pass_walk_is_short_on_activation_like_rows(lib):
    rng = np.random.default_rng(5)
    steps = []
    for _ in range(50):
        a, b, jumps, walked = seg(lib, bf16_squares(rng.standard_normal(4096) * 1.7), 512, 8)
        assert a.view(np.uint32) == b.view(np.uint32)
        steps.append(jumps + walked)
    assert max(steps) <= 40      # ~10 runs + ~12 real-FADD chunks instead of 4096 dependent adds
