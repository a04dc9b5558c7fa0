"""Host mirror of the reference's generate loop, src/inference/inference.go:173-254
(`generateTokensInternal`): prefill the prompt at position 0, then one token per Forward,
greedy argmax of the last row, stop on EOS / SequenceLength.  The loop stays on the host like
the Go original; each iteration is one call through the C-ABI.
"""
from __future__ import annotations

import time

import numpy as np

from . import ml, model as model_mod

GSInProgress, GSFinishedByReachingEOS, GSFinishedByReachingSeqLen = 0, 1, 2


class InferenceEngine:
    def __init__(self, model: "model_mod.Model", inferenceArgs: "model_mod.InferenceArgs", logFn=None,
                 acc_mode: int = model_mod.LNB_ACC_FAST, max_rows: int = 8):
        self.model = model
        self.inferenceArgs = inferenceArgs
        self.logFn = logFn
        self.acc_mode = acc_mode
        self.max_rows = max_rows
        self.context_hook = None  # optional callable(InferenceContext), e.g. to enable the peer all-reduce

    def CreateInferenceContext(self) -> "model_mod.InferenceContext":  # inference.go:256-258
        ctx = model_mod.InferenceContext(self.model.Transformer, self.inferenceArgs, self.logFn,
                                         max_rows=self.max_rows, acc_mode=self.acc_mode)
        if self.context_hook is not None:
            self.context_hook(ctx)
        return ctx

    # --- tokenizer half of the engine (src/inference/tokenize.go), served by model.Vocabulary when it is a loaded one
    def _vocab(self):
        v = self.model.Vocabulary
        if not hasattr(v, "TokenizeString"):
            raise ml.MlError("the model was loaded without tokenizer.model (ids only)")
        return v

    def Tokenize(self, promptParts):                 # tokenize.go:27-95
        return self._vocab().Tokenize(promptParts)

    def TokenizeBatch(self, prompts):                # :97-107
        return self._vocab().TokenizeBatch(prompts)

    def TokenizeString(self, text, addBeginOfSentence: bool = False):   # :175-193 (the flag is unused there too)
        return self._vocab().TokenizeString(text)

    def TokenToString(self, tokenId, decodingContext):                  # :195-237
        return self._vocab().TokenToString(tokenId, decodingContext)

    def TokenBatchToString(self, tokenIdBatch):      # :239-258
        return self._vocab().TokenBatchToString(tokenIdBatch)

    # --- text streaming (generateStringInternal, inference.go:91-170) ------------------------------------------
    def GenerateString(self, promptTokens, **kw):
        """GenerateString (inference.go:56-58): GeneratedPart stream for the tokens the model generates"""
        return self.GenerateStringGeneric(self.GenerateTokens(promptTokens, **kw))

    def GenerateStringFromOutputTokens(self, outputTokens):
        """GenerateStringFromOutputTokens (inference.go:60-69): replays given ids through the detokenizer (the
        reference's own tests drive the streaming decoder this way)"""
        return self.GenerateStringGeneric((GSInProgress, int(t)) for t in outputTokens)

    def GenerateStringGeneric(self, tokenStream):
        """yields dict(DecodedString, TokenId, AddedToWaiting, IsResendOfWaiting, GenerationState) per token.  Tokens
        that end inside a UTF-8 sequence are reported with AddedToWaiting and, if the stream ends before they
        complete, re-sent at the end with their raw bytes (IsResendOfWaiting), the last one carrying the final
        state -- inference.go:104-170.  WaitingRunesExtraStr (emoji alias annotation) is not produced."""
        from .vocabulary import GenerationDecodingContext
        vocab = self._vocab()
        ctx = GenerationDecodingContext()
        waiting, last_state = [], GSInProgress
        for state, tok in tokenStream:
            text, added = vocab.TokenToString(tok, ctx)
            part = dict(DecodedString=text, TokenId=tok, AddedToWaiting=added, IsResendOfWaiting=False, GenerationState=GSInProgress)
            if state != GSInProgress and not waiting:
                part["GenerationState"] = state
            last_state = state
            if added:
                waiting.append(part)
            elif waiting:
                waiting = []
            yield part
        for i, w in enumerate(waiting):
            yield dict(DecodedString=vocab.IdToToken(w["TokenId"]).decode("utf-8", "replace"), TokenId=w["TokenId"], AddedToWaiting=False,
                       IsResendOfWaiting=True, GenerationState=last_state if i + 1 == len(waiting) else GSInProgress)

    def GenerateTokens(self, promptTokens, use_reference_api: bool = False, step_times: list | None = None,
                       prefill_chunk: int = 0):
        """Generator over (state, token id) exactly like generatedTokensCh.

        use_reference_api=True goes through the reference-shaped calls of every iteration
        (Transformer.Forward -> logits tensor -> Slice last row -> ml.Argmax); False uses the
        fused forward+argmax entry (same kernels, 4-byte read-back).

        prefill_chunk > 0 (EXTENSION): the prompt is fed in chunks of that many tokens (Forward at
        startPos > 0 with S > 1, lnb_session_set_chunked_prefill) instead of one call; prompts longer than the context's
        max_rows become possible, and the generated stream is the same."""
        infContext = self.CreateInferenceContext()
        try:
            promptLength = len(promptTokens)
            if prefill_chunk > 0:
                infContext.allow_chunked_prefill(True)
                if prefill_chunk > infContext.max_rows:
                    raise ml.MlError(f"prefill_chunk {prefill_chunk} exceeds the context's max_rows {infContext.max_rows}")
            if promptLength >= infContext.SequenceLength:  # :176-179
                raise ml.MlError(f"context SequenceLength {infContext.SequenceLength} must be higher than prompt "
                                 f"tokens length {promptLength}")
            if promptLength > infContext.max_rows and prefill_chunk <= 0:
                raise ml.MlError(f"prompt length {promptLength} exceeds the context's max_rows {infContext.max_rows}")
            pad = self.model.Vocabulary.PadId
            tokens = ml.Full([infContext.SequenceLength], ml.DT_INT32, pad)  # :181
            tokens.RawData[:promptLength] = np.asarray(promptTokens, np.int32)
            prevPos = 0
            if prefill_chunk > 0:      # all but the last chunk: Forward calls whose logits nobody needs
                while promptLength - prevPos > prefill_chunk:
                    self.model.Transformer.forward_argmax(infContext, tokens.RawData[prevPos:prevPos + prefill_chunk], prevPos)
                    prevPos += prefill_chunk
            for curPos in range(promptLength, infContext.SequenceLength):  # :194
                t0 = time.perf_counter()
                inputTokensSlice = tokens.Slice([prevPos], [curPos])
                if use_reference_api:
                    logits = self.model.Transformer.Forward(infContext, inputTokensSlice, prevPos)  # :202
                    logits = logits.Slice([logits.Size[0] - 1], [logits.Size[0]])                   # :207
                    nextTokenId = int(ml.Argmax(logits, len(logits.Size) - 1).Item())              # :211
                else:
                    nextTokenId, _ = self.model.Transformer.forward_argmax(infContext, inputTokensSlice.RawData, prevPos)
                existing = int(tokens.RawData[curPos])
                if existing != pad:  # :218-226 only replace token if prompt has already been generated
                    nextTokenId = existing
                tokens.RawData[curPos] = nextTokenId
                if step_times is not None:
                    step_times.append(time.perf_counter() - t0)
                eos = nextTokenId in self.model.Vocabulary.StopTokenIds
                prevPos = curPos
                if eos:
                    yield GSFinishedByReachingEOS, nextTokenId
                    break
                if curPos + 1 == infContext.SequenceLength:
                    yield GSFinishedByReachingSeqLen, nextTokenId
                    break
                yield GSInProgress, nextTokenId
        finally:
            infContext.close()

    def GenerateTokensBatch(self, prompts, step_times: list | None = None):
        """The consumer `InferenceEngine.TokenizeBatch` (src/inference/tokenize.go:97-107) never got in the
        reference: every prompt runs the loop of generateTokensInternal (:173-254) -- own
        InferenceContext / KV cache, own positions, own stop condition -- but all sequences advance together,
        one pass over the weights per step (lnb_forward_batch).  Yields (sequence index, state, token id);
        per sequence the stream is exactly what GenerateTokens(prompts[i]) yields."""
        n = len(prompts)
        if n < 1:
            raise ml.MlError("empty prompt batch")
        if n == 1:
            for state, tok in self.GenerateTokens(prompts[0]):
                yield 0, state, tok
            return
        infContext = model_mod.InferenceContext(self.model.Transformer, self.inferenceArgs, self.logFn,
                                                max_rows=max(self.max_rows, n), acc_mode=self.acc_mode, n_seq=n)
        if self.context_hook is not None:
            self.context_hook(infContext)
        try:
            seqLen = infContext.SequenceLength
            stop = self.model.Vocabulary.StopTokenIds
            for p in prompts:
                if len(p) >= seqLen:  # :176-179
                    raise ml.MlError(f"context SequenceLength {seqLen} must be higher than prompt tokens length {len(p)}")
                if len(p) < 1 or len(p) > infContext.max_rows:
                    raise ml.MlError(f"prompt length {len(p)} must be in [1, {infContext.max_rows}]")
            cur = [0] * n                                  # token fed next / position it is fed at
            pos = [0] * n
            done = [False] * n

            def emit(i, tok, curPos):
                """bookkeeping of one generated token at curPos (:218-248); returns the state to yield"""
                cur[i], pos[i] = tok, curPos
                if tok in stop:
                    done[i] = True
                    return GSFinishedByReachingEOS
                if curPos + 1 == seqLen:
                    done[i] = True
                    return GSFinishedByReachingSeqLen
                return GSInProgress

            t0 = time.perf_counter()
            for i, p in enumerate(prompts):                # prefill: the ordinary Forward on sequence i's cache
                infContext.set_active_sequence(i)
                nxt, _ = self.model.Transformer.forward_argmax(infContext, np.asarray(p, np.int32), 0)
                yield i, emit(i, int(nxt), len(p)), int(nxt)
            if step_times is not None:
                step_times.append(time.perf_counter() - t0)
            while not all(done):
                t0 = time.perf_counter()
                # finished sequences are stepped again at their last position (same token, same cache row: a no-op
                # for their state); the weights are read once per step regardless
                nxt, _ = infContext.forward_batch(cur, pos)
                if step_times is not None:
                    step_times.append(time.perf_counter() - t0)
                for i in range(n):
                    if not done[i]:
                        yield i, emit(i, int(nxt[i]), pos[i] + 1), int(nxt[i])
        finally:
            infContext.close()
