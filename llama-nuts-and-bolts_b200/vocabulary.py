"""Host mirror of src/tiktoken + model.Vocabulary + the tokenizer half of src/inference/tokenize.go over the native
tokenizer of liblnb.so (csrc/tokenizer.cpp).  Names follow the reference: `Load(path)` ~ tiktoken.Load +
model.NewVocabulary, `TokenizeString`, `Tokenize(promptParts)`, `TokenizeBatch`, `TokenBatchToString`, and a
streaming `TokenToString` with the reference's byte-fallback waiting (tokens that end inside a UTF-8 sequence are
held back until the sequence completes, tokenize.go:195-237).  The console's emoji alias annotation
(src/inference/emoji.go) is not reproduced: its data tables are third-party modules that are not in the reference tree."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import check, lib

B_TXT, B_HEADER, E_HEADER, E_TURN = "<|begin_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"


@dataclass
class PromptPart:  # inference.PromptPart (tokenize.go:21-25)
    Header: str
    Content: str


@dataclass
class GenerationDecodingContext:  # the byte-fallback half of generationDecodingContext
    waitingBytes: bytearray = field(default_factory=bytearray)


class Vocabulary:
    """model.Vocabulary (src/model/vocabulary.go:9-21) backed by lnb_vocab"""

    def __init__(self, h):
        self.h = h
        bos, eos, pad = C.c_int32(), C.c_int32(), C.c_int32()
        stop = (C.c_int32 * 2)()
        check(lib.lnb_vocab_special_ids(h, C.byref(bos), C.byref(eos), C.byref(pad), stop))
        self.BeginOfSentenceId, self.EndOfSentenceId, self.PadId, self.UnknownId = bos.value, eos.value, pad.value, -1
        self.StopTokenIds = (int(stop[0]), int(stop[1]))

    def __len__(self):
        return check(lib.lnb_vocab_size(self.h))

    def TokenToId(self, token: bytes | str) -> int:
        b = token.encode() if isinstance(token, str) else bytes(token)
        out = C.c_int32()
        check(lib.lnb_vocab_token_id(self.h, b, len(b), C.byref(out)))
        return out.value

    def IdToToken(self, tokenId: int) -> bytes:
        p, n = C.c_void_p(), C.c_int()
        check(lib.lnb_vocab_token_bytes(self.h, tokenId, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    # --- InferenceEngine.TokenizeString / Tokenize / TokenizeBatch (tokenize.go:27-107,175-193) -------------
    def TokenizeString(self, text: str | bytes) -> list[int]:
        b = text.encode("utf-8", "surrogateescape") if isinstance(text, str) else bytes(text)
        n = C.c_int(0)
        cap = len(b) + 8                                    # a piece never yields more tokens than it has bytes
        out = np.empty(cap, np.int32)
        check(lib.lnb_tokenize_string(self.h, b, len(b), _capi.ptr(out, _capi.i32p), cap, C.byref(n)))
        return out[:n.value].tolist()

    def Tokenize(self, promptParts: list[PromptPart]) -> list[int]:
        k = len(promptParts)
        hs = (C.c_char_p * max(1, k))(*[p.Header.encode() for p in promptParts])
        cs = (C.c_char_p * max(1, k))(*[p.Content.encode() for p in promptParts])
        cap = sum(len(p.Header.encode()) + len(p.Content.encode()) + 16 for p in promptParts) + 32
        out = np.empty(cap, np.int32)
        n = C.c_int(0)
        check(lib.lnb_tokenize_prompt(self.h, hs, cs, k, _capi.ptr(out, _capi.i32p), cap, C.byref(n)))
        return out[:n.value].tolist()

    def TokenizeBatch(self, prompts: list[list[PromptPart]]) -> list[list[int]]:
        return [self.Tokenize(p) for p in prompts]

    # --- detokenizer ----------------------------------------------------------------------------------------
    def TokenBatchToBytes(self, tokenIdBatch) -> bytes:
        ids = np.ascontiguousarray(tokenIdBatch, np.int32)
        n = C.c_int64(0)
        cap = 64 + 128 * len(ids)
        while True:
            buf = C.create_string_buffer(cap)
            rc = lib.lnb_detokenize(self.h, _capi.ptr(ids, _capi.i32p), len(ids), buf, cap, C.byref(n))
            if rc == 0:
                return buf.raw[:n.value]
            if n.value > cap:
                cap = n.value
                continue
            check(rc)

    def TokenToString(self, tokenId: int, decodingContext: GenerationDecodingContext) -> tuple[str, bool]:
        """(text ready to print, addedToWaiting): pieces that are not valid UTF-8 on their own are collected in
        waitingBytes and released rune by rune once they decode (tokenize.go:195-237)"""
        piece = self.IdToToken(tokenId)
        try:
            text = piece.decode("utf-8")
            if not decodingContext.waitingBytes:
                return text, False
        except UnicodeDecodeError:
            text = None
        if text is not None:                       # a complete piece arrives while bytes are waiting: they stay waiting
            return text, False
        decodingContext.waitingBytes += piece
        try:
            s = bytes(decodingContext.waitingBytes).decode("utf-8")
        except UnicodeDecodeError:
            return "", True
        first = s[0]                               # one rune per call, like utf8.DecodeRune (:225-227)
        del decodingContext.waitingBytes[:len(first.encode())]
        return first, False

    def TokenBatchToString(self, tokenIdBatch) -> str:
        ctx, out = GenerationDecodingContext(), []
        for t in tokenIdBatch:
            if t == self.PadId:
                break
            s, waiting = self.TokenToString(int(t), ctx)
            if not waiting:
                out.append(s)
        return "".join(out)

    def close(self):
        if self.h:
            lib.lnb_vocab_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def SplitPieces(text: str | bytes) -> list[bytes]:
    """Vocabulary.SplitRegexp.FindAllString(text, -1) (vocabulary.go:36): the pre-tokenization pieces"""
    b = text.encode("utf-8", "surrogateescape") if isinstance(text, str) else bytes(text)
    ends = np.empty(len(b) + 1, np.int64)
    n = C.c_int(0)
    check(lib.lnb_split_pieces(b, len(b), _capi.ptr(ends, _capi.i64p), len(ends), C.byref(n)))
    out, p = [], 0
    for e in ends[:n.value].tolist():
        out.append(b[p:e])
        p = e
    return out


def Load(vocabFilePath: str) -> Vocabulary:
    """tiktoken.Load + model.NewVocabulary (loader.go:84-96)"""
    h = C.c_void_p()
    check(lib.lnb_vocab_load(vocabFilePath.encode(), C.byref(h)))
    return Vocabulary(h)
