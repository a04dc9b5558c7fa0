"""Opt-in: the reference's WEIGHT-DEPENDENT goldens (src/model/llamatransformer_simulated_test.go), for
whoever has the real checkpoint.  Set LNB_MODEL_DIR to a Meta-Llama-3.1-8B-Instruct directory (params.json,
consolidated.00.pth, tokenizer.model -- what the reference calls models-original/...).  Skipped otherwise: the files
do not exist offline.  The expected values come from tests/golden/reference_vectors.json (extracted from the
reference's test source)."""
import os

import numpy as np
import pytest

from tests.helpers import reference_vectors

MODEL_DIR = os.environ.get("LNB_MODEL_DIR", "")
needs_model = pytest.mark.skipif(not (MODEL_DIR and os.path.exists(os.path.join(MODEL_DIR, "consolidated.00.pth"))),
                                 reason="LNB_MODEL_DIR does not point to a Meta-Llama-3.1-8B-Instruct directory")
GOLD = reference_vectors()["simulated_only_first_layer"]


@needs_model
def test_real_tokenizer_reproduces_the_reference_prompt_tokens():
    """Tokenize([user: "What is your name?"]) must give the 15 ids the reference's test hard-codes (:1369)"""
    import lnb_b200 as L
    from lnb_b200.vocabulary import Load, PromptPart
    v = Load(os.path.join(MODEL_DIR, "tokenizer.model"))
    assert len(v) == 128256
    assert v.Tokenize([PromptPart("user", "What is your name?")]) == GOLD["prompt_tokens"]
    assert v.TokenBatchToString(GOLD["prompt_tokens"]) == GOLD["prompt_text"]
    v.close()


@needs_model
@pytest.mark.gpu
def test_simulated_only_first_layer_tokens_and_logits():
    """TestSimulatedOnlyFirstLayer: layer 0 -> output norm -> LM head on the 15-token prompt, SequenceLength 20;
    the five generated ids must equal the reference's, the logits corners agree within its 0.3 tolerance"""
    import lnb_b200 as L
    m = L.model.LoadModel(MODEL_DIR)
    try:
        ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(GOLD["sequence_length"]), max_rows=16,
                                       acc_mode=L._capi.LNB_ACC_STRICT)
        ctx.set_layer_limit(1)
        prompt = np.array(GOLD["prompt_tokens"], np.int32)
        logits = m.Transformer.Forward(ctx, L.ml.Tensor(prompt, L.ml.DT_INT32), 0).RawData
        assert logits.shape == (15, 128256)
        for row, exp in zip(GOLD["logits_rows"], GOLD["logits_first3_last3"]):
            got = np.concatenate([logits[row, :3], logits[row, -3:]])
            assert np.abs(got - np.array(exp, np.float32)).max() <= 0.3, (row, got, exp)
        out, pos, tok = [], len(prompt), int(np.argmax(logits[-1]))
        out.append(tok)
        while pos + 1 < GOLD["sequence_length"]:
            tok, _ = m.Transformer.forward_argmax(ctx, np.array([tok], np.int32), pos)
            out.append(int(tok)); pos += 1
        assert out == GOLD["expected_output_tokens"]
        ctx.close()
    finally:
        m.Free()
