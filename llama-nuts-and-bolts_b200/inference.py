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

    def GenerateTokens(self, promptTokens, use_reference_api: bool = False, step_times: list | None = None):
        """Generator over (state, token id) exactly like generatedTokensCh.

        use_reference_api=True goes through the reference-shaped calls of every iteration
        (Transformer.Forward -> logits tensor -> Slice last row -> ml.Argmax); False uses the
        fused forward+argmax entry (same kernels, 4-byte read-back)."""
        infContext = self.CreateInferenceContext()
        try:
            promptLength = len(promptTokens)
            if promptLength >= infContext.SequenceLength:  # :176-179
                raise ml.MlError(f"context SequenceLength {infContext.SequenceLength} must be higher than prompt "
                                 f"tokens length {promptLength}")
            if promptLength > infContext.max_rows:
                raise ml.MlError(f"prompt length {promptLength} exceeds the context's max_rows {infContext.max_rows}")
            pad = self.model.Vocabulary.PadId
            tokens = ml.Full([infContext.SequenceLength], ml.DT_INT32, pad)  # :181
            tokens.RawData[:promptLength] = np.asarray(promptTokens, np.int32)
            prevPos = 0
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
