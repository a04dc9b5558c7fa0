"""the native tiktoken / Llama-3 BPE tokenizer (csrc/tokenizer.cpp) against independent implementations.

The reference's bytePairMerge is "ported from Tiktoken Rust code" (src/inference/tokenize.go:110-112), so the
`tiktoken` package is the upstream second opinion for the merge; the `regex` module is the second opinion for the
split pattern of src/model/vocabulary.go:36.  Both are given the pattern with Go/RE2's ASCII-only `\\s`, which is
what the reference's regexp package implements.  No tokenizer.model exists offline, so the vocabulary is a
synthetic one (256 bytes in shuffled rank order + merges trained on a small multilingual corpus)."""
import base64
import collections
import random
import unicodedata

import numpy as np
import pytest
import regex
import tiktoken

import lnb_b200 as L
from lnb_b200.vocabulary import GenerationDecodingContext, Load, PromptPart, SplitPieces

# src/model/vocabulary.go:36 with \s spelled out the way RE2 defines it
GO_PAT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\t\n\f\r \p{L}\p{N}]+[\r\n]*"
          r"|[\t\n\f\r ]*[\r\n]+|[\t\n\f\r ]+")
CORPUS = ("Hello, world! It's a beautiful day, isn't it? We've been told they'll come; I'd say so.\n\n"
          "The quick brown fox jumps over 13 lazy dogs in 2024.  Tabs\tand\r\nnewlines   \n  spaces.\n"
          "naïve café Ünïcödé Ελληνικά кириллица 日本語のテキスト 中文 العربية हिन्दी 12345678 3.14159 x=y+z*2 {code} [brackets] <tags/>\n"
          "emoji 😀🇹🇷 and symbols ©®™ ±×÷ §¶ … — “quotes” ‘single’ don'T I'M you'RE 'ſ\n") * 3


def train_bpe(corpus, n_merges, seed=3):
    pieces = [m.group(0).encode() for m in regex.finditer(GO_PAT, corpus)]
    words = collections.Counter(tuple(bytes([b]) for b in p) for p in pieces)
    order = list(range(256))
    random.Random(seed).shuffle(order)
    ranks = {bytes([b]): i for i, b in enumerate(order)}
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), _ = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        ranks.setdefault(a + b, len(ranks))
        nw = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(w[i]); i += 1
            nw[tuple(out)] += c
        words = nw
    return ranks


def write_model(path, ranks):
    with open(path, "w") as f:
        for tok, r in sorted(ranks.items(), key=lambda kv: kv[1]):
            f.write(base64.b64encode(tok).decode() + " " + str(r) + "\n")


@pytest.fixture(scope="module")
def vocab(tmp_path_factory):
    ranks = train_bpe(CORPUS, 300)
    path = str(tmp_path_factory.mktemp("tok") / "tokenizer.model")
    write_model(path, ranks)
    v = Load(path)
    yield v, ranks
    v.close()


def sample_texts():
    texts = [CORPUS, "", " ", "a", "  leading spaces", "trailing   ", "\n\n\n", " \n \n x", "it's IT'S 'Re 'LL 'd 'VE 'M", "1234567 12 1",
             "!!!???...", "a\tb\fc", "x \r\n\r\n y", " nbsp ideographic　space", "'ſ ſ", "a  b", "١٢٣٤٥ Ⅻ ½"]
    rnd = random.Random(5)
    cps = list(range(32, 127)) + [9, 10, 13, 12] + list(range(0xA0, 0x250)) + list(range(0x370, 0x400)) + \
        list(range(0x4E00, 0x4E40)) + [0x1F600, 0x1F1F9, 0x200D, 0x2028, 0x3000, 0x660, 0x2160, 0xBC, 0x0301, 0xFE0F]
    alphabet = [chr(c) for c in cps if unicodedata.category(chr(c)) != "Cn"]
    for _ in range(400):
        texts.append("".join(rnd.choice(alphabet) if rnd.random() < 0.6 else rnd.choice(" \n'abctsrevmld19")
                             for _ in range(rnd.randrange(1, 60))))
    return texts


def test_vocabulary_layout_follows_tiktokenreader_go(vocab):
    v, ranks = vocab
    n = len(ranks)
    assert len(v) == n + 256                                          # reservedSpecialTokensCount (tiktokenreader.go:44)
    names = ["<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
             "<|finetune_right_pad_id|>", "<|step_id|>", "<|start_header_id|>", "<|end_header_id|>", "<|eom_id|>", "<|eot_id|>",
             "<|python_tag|>"] + [f"<|reserved_special_token_{2 + i}|>" for i in range(245)]
    for i, name in enumerate(names):
        assert v.TokenToId(name) == n + i and v.IdToToken(n + i) == name.encode()
    assert (v.BeginOfSentenceId, v.EndOfSentenceId, v.PadId, v.UnknownId) == (n, n + 1, -1, -1)
    assert v.StopTokenIds == (n + 8, n + 9)                           # <|eom_id|>, <|eot_id|> (:81)
    for tok, r in list(ranks.items())[::17]:
        assert v.TokenToId(tok) == r and v.IdToToken(r) == tok
    assert v.TokenToId(b"\xff\xfe not a token") == -1


def test_split_pattern_equals_the_regex_module(vocab):
    """Vocab::next_piece is a hand-written matcher for vocabulary.go:36: identical pieces to the `regex` module"""
    for t in sample_texts():
        assert SplitPieces(t) == [m.group(0).encode() for m in regex.finditer(GO_PAT, t)], repr(t)
    assert SplitPieces(b"ab\xff\xfe cd\xc3") == [b"ab", b"\xff\xfe", b" cd", b"\xc3"]   # stray bytes are "other" characters


def test_bpe_equals_tiktoken(vocab):
    v, ranks = vocab
    enc = tiktoken.Encoding("synthetic", pat_str=GO_PAT, mergeable_ranks=ranks, special_tokens={})
    for t in sample_texts():
        assert v.TokenizeString(t) == enc.encode_ordinary(t), repr(t)


def test_reference_pattern_quirks():

    def pieces(t):
        return [p.decode() for p in SplitPieces(t)]
    # Go's regexp has no lookahead, so the reference dropped `\s+(?!\S)` from Meta's pattern: a run of spaces before
    # a word stays whole instead of lending its last space to the word
    assert pieces("a  b") == ["a", "  ", "b"]
    assert pieces("a b") == ["a", " b"]
    # RE2's \s is ASCII only: U+3000 / U+00A0 are "other" characters, not white space
    assert pieces("x　y") == ["x", "　y"]
    # (?i) folds with Unicode simple folding: 's also matches LATIN SMALL LETTER LONG S
    assert pieces("it'ſ") == ["it", "'ſ"] and pieces("IT'S") == ["IT", "'S"]
    assert pieces("12345") == ["123", "45"] and pieces("x\r\n\r\n  y") == ["x", "\r\n\r\n", "  ", "y"]


def test_chat_template_follows_tokenize_go(vocab):
    v, ranks = vocab
    n = len(ranks)
    B_TXT, B_HDR, E_HDR, E_TURN = n, n + 6, n + 7, n + 9
    parts = [PromptPart("system", "You are a pirate."), PromptPart("user", ""), PromptPart("user", "Hello there!\nHow are you?")]
    got = v.Tokenize(parts)
    nn = v.TokenizeString("\n\n")
    exp = [B_TXT]
    for p in (parts[0], parts[2]):                                    # the empty part is skipped (tokenize.go:43-45)
        exp += [B_HDR] + v.TokenizeString(p.Header) + [E_HDR] + nn + v.TokenizeString(p.Content) + [E_TURN]
    exp += [B_HDR] + v.TokenizeString("assistant") + [E_HDR] + nn    # open assistant turn, no <|eot_id|> (:36-40,:74-78)
    assert got == exp
    assert v.Tokenize([]) == [B_TXT, B_HDR] + v.TokenizeString("assistant") + [E_HDR] + nn
    assert v.TokenizeBatch([parts, []]) == [got, v.Tokenize([])]
    text = v.TokenBatchToString(got)
    assert text == ("<|begin_of_text|><|start_header_id|>system<|end_header_id|>\n\nYou are a pirate.<|eot_id|>"
                    "<|start_header_id|>user<|end_header_id|>\n\nHello there!\nHow are you?<|eot_id|>"
                    "<|start_header_id|>assistant<|end_header_id|>\n\n")


def test_streaming_detokenizer_holds_back_incomplete_utf8(tmp_path):
    # a byte-only vocabulary: every multi-byte character arrives as separate byte tokens (the "byte fallback" case)
    ranks = {bytes([b]): b for b in range(256)}
    path = str(tmp_path / "bytes.model")
    write_model(path, ranks)
    v = Load(path)
    ids = v.TokenizeString("a🇹é")
    assert ids == list("a🇹é".encode())
    ctx = GenerationDecodingContext()
    outs = [v.TokenToString(i, ctx) for i in ids]
    assert outs == [("a", False), ("", True), ("", True), ("", True), ("🇹", False), ("", True), ("é", False)]
    assert not ctx.waitingBytes
    assert v.TokenBatchToString(ids) == "a🇹é"
    assert v.TokenBatchToString(ids[:3] + [-1] + ids[3:]) == "a"      # PadId ends the batch (tokenize.go:246-248)
    assert v.TokenBatchToBytes(ids[:3]) == "a🇹é".encode()[:3]
    with pytest.raises(L._capi.LnbError):
        v.TokenBatchToBytes([len(v)])
    v.close()


def test_pieces_outside_the_vocabulary_become_id_zero(tmp_path):
    """Go's map lookup yields 0 for a missing piece (tokenize.go:166-170); unreachable with the real vocabulary, which
    holds all 256 bytes, but part of the reference's behaviour"""
    ranks = {bytes([b]): i for i, b in enumerate(b"abc ")}
    path = str(tmp_path / "abc.model")
    write_model(path, ranks)
    v = Load(path)
    assert v.TokenizeString("ab zc") == [0, 1, 3, 0, 2]
    v.close()


def test_malformed_tokenizer_models_are_errors(tmp_path):
    def load(text):
        p = str(tmp_path / "m.model")
        open(p, "w").write(text)
        return Load(p)
    for bad in ("", "YQ==\n", "YQ== x\n", "!!!! 0\n", "YQ== 0\nYg== 0\n", "YQ== 5\n", "YQ= 0\n"):
        with pytest.raises(L._capi.LnbError):
            load(bad)
    with pytest.raises(L._capi.LnbError):
        Load(str(tmp_path / "missing.model"))
    v = load("YQ== 0\r\nYg== 1")                                       # CRLF and a missing final newline are fine
    assert len(v) == 2 + 256 and v.IdToToken(1) == b"b"
    v.close()


def test_cpp_host_tokenizer_equals_the_python_mirror(vocab, tmp_path):
    """host/lnb_host.hpp model::Tokenizer (C++ mirror) through `lnb_generate --tokenize`: same ids, lossless text"""
    import os
    import subprocess
    v, ranks = vocab
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "host"), "-s"])
    path = str(tmp_path / "tokenizer.model")
    write_model(path, ranks)
    text = "It's naïve: 12345 dogs 🇹🇷\n\nok?"
    out = subprocess.run([os.path.join(root, "host", "lnb_generate"), "--tokenize", path, text], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ids = [int(t) for t in out.stdout.splitlines()[0].split()[1:]]
    assert ids == v.Tokenize([PromptPart("user", text)])
    assert out.stdout.split("text: ", 1)[1].removesuffix("\n") == v.TokenBatchToString(ids)


def test_bpe_with_adversarial_rank_order_equals_tiktoken(tmp_path):
    """overlapping merges over a 3-letter alphabet in random rank order: the heap-based merge must pick the same pairs
    as tiktoken's scan (lowest rank, leftmost on ties), also on a 200 KB single piece"""
    rnd = random.Random(1)
    ranks = {bytes([b]): b for b in range(256)}
    while len(ranks) < 256 + 400:
        t = bytes(rnd.choice(b"abc") for _ in range(rnd.randrange(2, 7)))
        ranks.setdefault(t, len(ranks))
    path = str(tmp_path / "abc.model")
    write_model(path, ranks)
    v = Load(path)
    enc = tiktoken.Encoding("abc", pat_str=r"[abc]+|[^abc]+", mergeable_ranks=ranks, special_tokens={})
    for _ in range(300):
        t = "".join(rnd.choice("abc") for _ in range(rnd.randrange(1, 80)))
        assert v.TokenizeString(t) == enc.encode_ordinary(t), t
    big = "".join(rnd.choice("abc") for _ in range(200_000))
    ids = v.TokenizeString(big)
    assert v.TokenBatchToBytes(ids) == big.encode() and ids == enc.encode_ordinary(big)
    v.close()


def test_generate_string_streams_text_like_generate_string_internal(tmp_path):
    """InferenceEngine.GenerateStringFromOutputTokens (the reference's tests drive the streaming decoder with it,
    cmd/main_test.go:70-93): parts in order, waiting flags for split UTF-8, resend of an unfinished tail"""
    ranks = {bytes([b]): b for b in range(256)}
    ranks[b"ab"] = 256
    path = str(tmp_path / "bytes.model")
    write_model(path, ranks)
    v = Load(path)

    class FakeModel:            # the engine only needs model.Vocabulary here
        Vocabulary = v
        Transformer = None
    eng = L.inference.InferenceEngine(FakeModel, L.model.InferenceArgs(16))
    flag = "🇹".encode()        # F0 9F 87 B9 arrives as four byte tokens
    ids = [256] + list(flag) + [ord("!")]
    parts = list(eng.GenerateStringFromOutputTokens(ids))
    assert [p["DecodedString"] for p in parts] == ["ab", "", "", "", "🇹", "!"]
    assert [p["AddedToWaiting"] for p in parts] == [False, True, True, True, False, False]
    assert all(p["GenerationState"] == L.inference.GSInProgress and not p["IsResendOfWaiting"] for p in parts)
    # the stream ends inside a character: the held-back tokens are re-sent, the last one carries the final state
    stream = [(L.inference.GSInProgress, ord("x")), (L.inference.GSInProgress, flag[0]), (L.inference.GSFinishedByReachingSeqLen, flag[1])]
    parts = list(eng.GenerateStringGeneric(stream))
    assert [(p["AddedToWaiting"], p["IsResendOfWaiting"]) for p in parts] == [(False, False), (True, False), (True, False), (False, True), (False, True)]
    assert [p["GenerationState"] for p in parts] == [L.inference.GSInProgress] * 4 + [L.inference.GSFinishedByReachingSeqLen]
    assert [p["TokenId"] for p in parts[3:]] == [flag[0], flag[1]]
    # a finished state on a complete token is reported immediately
    parts = list(eng.GenerateStringGeneric([(L.inference.GSFinishedByReachingEOS, ord("z"))]))
    assert parts == [dict(DecodedString="z", TokenId=ord("z"), AddedToWaiting=False, IsResendOfWaiting=False,
                          GenerationState=L.inference.GSFinishedByReachingEOS)]
    with pytest.raises(L.ml.MlError):
        L.inference.InferenceEngine(type("M", (), {"Vocabulary": L.model.Vocabulary(), "Transformer": None}), L.model.InferenceArgs(16)).Tokenize([])
    v.close()


def test_cpp_streaming_detokenizer_equals_the_python_mirror(tmp_path):
    """model::Tokenizer::TokenToString (C++) through `lnb_generate --detok-stream`: same waiting flags and text as
    Vocabulary.TokenToString (Python) on a character split across byte tokens"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "host"), "-s"])
    path = str(tmp_path / "bytes.model")
    write_model(path, {bytes([b]): b for b in range(256)})
    v = Load(path)
    ids = list("a🇹é!".encode())
    out = subprocess.run([os.path.join(root, "host", "lnb_generate"), "--detok-stream", path] + [str(i) for i in ids],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ctx = GenerationDecodingContext()
    exp = []
    for i in ids:
        text, waiting = v.TokenToString(i, ctx)
        exp.append(f"{i} {'waiting' if waiting else 'text'} [{text}]")
    assert out.stdout.splitlines() == exp
    v.close()
