"""Property tests (hypothesis) of the host-side rows: arbitrary Unicode text through the tokenizer, arbitrary tensors through the
.pth writer / reader, arbitrary non-negative rows through the exact-sum model.  Independent implementations are the oracle of
each property: the `regex` module for the split pattern, `tiktoken` for the merge, PyTorch for the file format, the plain
sequential loop for the sum."""
import ctypes as C
import os
import unicodedata

import numpy as np
import pytest
import regex
import tiktoken
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import lnb_b200 as L
from lnb_b200.torch_reader import TorchModelReader, TorchModelWriter
from lnb_b200.vocabulary import Load, SplitPieces
from tests.test_tokenizer_cpu import CORPUS, GO_PAT, train_bpe, write_model

# code points assigned in Unicode 15.0 (the tables of csrc/unicode_tables.hpp); newer assignments may be classified differently
# by the regex module's newer database
text15 = st.text(alphabet=st.characters(blacklist_categories=("Cs", "Cn")), max_size=200)
COMMON = dict(deadline=None, derandomize=True, database=None,          # same examples on every run: no flaky round-end surprises
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@pytest.fixture(scope="module")
def vocab(tmp_path_factory):
    ranks = train_bpe(CORPUS, 300)
    path = str(tmp_path_factory.mktemp("tokp") / "tokenizer.model")
    write_model(path, ranks)
    v = Load(path)
    yield v, tiktoken.Encoding("prop", pat_str=GO_PAT, mergeable_ranks=ranks, special_tokens={})
    v.close()


@settings(max_examples=400, **COMMON)
@given(text15)
def test_split_pieces_equal_the_regex_module_for_any_text(t):
    assert SplitPieces(t) == [m.group(0).encode() for m in regex.finditer(GO_PAT, t)]


@settings(max_examples=300, **COMMON)
@given(text15)
def test_tokenize_equals_tiktoken_and_round_trips_for_any_text(vocab, t):
    v, enc = vocab
    ids = v.TokenizeString(t)
    assert ids == enc.encode_ordinary(t)
    assert v.TokenBatchToBytes(ids) == t.encode()
    # (the streaming TokenToString is NOT lossless in general, by the reference's design: it releases one rune per call,
    #  so a token that completes a character AND starts with other text -- e.g. b" \xc2" + b"\x80" -- leaves bytes waiting;
    #  tokenize.go:222-231.  The byte-level cases are covered in test_tokenizer_cpu.py.)


@settings(max_examples=200, **COMMON)
@given(st.binary(max_size=120))
def test_tokenizer_accepts_arbitrary_bytes(vocab, b):
    v, _ = vocab
    pieces = SplitPieces(b)
    assert b"".join(pieces) == b and all(pieces)
    assert v.TokenBatchToBytes(v.TokenizeString(b)) == b


DTYPES = [(np.uint16, torch.bfloat16), (np.float16, torch.float16), (np.float32, torch.float32), (np.float64, torch.float64), (np.int8, torch.int8),
          (np.uint8, torch.uint8), (np.int16, torch.int16), (np.int32, torch.int32), (np.int64, torch.int64), (np.bool_, torch.bool)]
tensor_specs = st.lists(st.tuples(st.integers(0, 9), st.lists(st.integers(0, 5), max_size=4), st.integers(0, 2**31)), min_size=1, max_size=6)


@settings(max_examples=60, **COMMON)
@given(tensor_specs)
def test_pth_writer_reader_and_torch_agree_for_any_tensors(tmp_path, specs):
    p = str(tmp_path / "h.pth")
    arrs = {}
    w = TorchModelWriter(p)
    for i, (dt, shape, seed) in enumerate(specs):
        npd, _ = DTYPES[dt]
        rng = np.random.default_rng(seed)
        n = int(np.prod(shape)) if shape else 1
        raw = rng.integers(0, 256, size=n * np.dtype(npd).itemsize, dtype=np.uint8)
        a = (raw.view(npd) if npd is not np.bool_ else (raw & 1).astype(np.bool_)).reshape(shape)
        arrs[f"t{i}.weight"] = (a, dt)
        w.Add(f"t{i}.weight", a, dt)
    w.Finish()
    back = torch.load(p, weights_only=True)
    with TorchModelReader(p) as r:
        mine = r.Load()
        assert list(back) == list(arrs) == list(mine)
        for k, (a, dt) in arrs.items():
            assert back[k].dtype == DTYPES[dt][1] and tuple(back[k].shape) == a.shape
            tb = back[k].view(torch.uint16).numpy() if dt == 0 else back[k].numpy()
            assert tb.tobytes() == a.tobytes(), k                        # bit patterns, NaNs included
            assert mine[k].Size == a.shape and mine[k].RawData.tobytes() == a.tobytes(), k


@settings(max_examples=40, **COMMON)
@given(st.lists(st.tuples(st.sampled_from([torch.bfloat16, torch.float32, torch.int64, torch.uint8]), st.lists(st.integers(1, 4), min_size=1, max_size=3)),
                min_size=1, max_size=5), st.sampled_from([2, 3, 4, 5]))
def test_reader_matches_torch_save_for_any_state_dict(tmp_path, specs, proto):
    p = str(tmp_path / "s.pth")
    sd = {}
    for i, (dt, shape) in enumerate(specs):
        t = (torch.rand(shape) * 100).to(dt)
        sd[f"k{i}"] = t if i % 2 == 0 else t.flatten()[: max(1, t.numel() - 1)]        # views with shared / offset storage too
    torch.save(sd, p, pickle_protocol=proto)
    with TorchModelReader(p) as r:
        mine = r.Load()
        for k, t in sd.items():
            tb = t.contiguous().view(torch.uint16).numpy() if t.dtype == torch.bfloat16 else t.contiguous().numpy()
            assert mine[k].contiguous and mine[k].RawData.tobytes() == tb.tobytes(), k


# --- exact parallel evaluation of the sequential sum (csrc/seqsum.cuh host model) -----------------------------------------
from tests.test_seqsum_host import lib as seqsum_lib  # noqa: E402  (the module-scoped fixture that builds the host model)

rows = st.lists(st.floats(min_value=0.0, max_value=float(np.float32(3.0e38)), allow_nan=False, width=32), min_size=256, max_size=256)


@settings(max_examples=150, **COMMON)
@given(rows, st.integers(0, 2**31 - 1))
def test_one_pass_sum_is_exact_for_any_non_negative_row(seqsum_lib, vals, sabotage):
    t = np.ascontiguousarray(vals, np.float32)
    a = np.float32(seqsum_lib.seqsum_reference(t.ctypes.data, t.size))
    for sab in (0, sabotage):
        b = np.float32(seqsum_lib.seqsum_seg_model(t.ctypes.data, 32, 8, sab, None, None))
        assert a.view(np.uint32) == b.view(np.uint32)
    it = C.c_int(0)
    c = np.float32(seqsum_lib.seqsum_scan_model(t.ctypes.data, 32, 8, C.byref(it)))
    assert a.view(np.uint32) == c.view(np.uint32)
