"""GPU parity of the model-level C-ABI (LlamaTransformer.Forward / generate loop) against the CPU
oracle on a structurally identical miniature of Llama-3.1 (all fusion paths, GQA, RoPE, KV cache).
STRICT mode must reproduce the oracle bit for bit; FAST within north_star's tolerance."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import bf16_ulp_diff, host_tensors, oracle_model

pytestmark = pytest.mark.gpu

SEED = 1234


@pytest.fixture(scope="module")
def L():
    import lnb_b200
    return lnb_b200


@pytest.fixture(scope="module")
def tiny(L):
    args = dict(L.synth.TINY)
    tensors = host_tensors(args, SEED)
    om = oracle_model(args, tensors)
    gm = L.model.LoadModelFromTensors(args, tensors)       # the reference's flow: host tensors -> upload
    yield args, tensors, om, gm
    gm.Free()
    om.close()


def test_tables_match_oracle(L, tiny):
    args, _, _, gm = tiny
    _, cis = O.rope_table(args["head_dim"], args["max_seq_len"] * 2, args["rope_theta"], bool(args["use_scaled_rope"]))
    assert np.array_equal(gm.Transformer.rope_table(), cis)
    assert np.array_equal(gm.Transformer.silu_table(), O.silu_table_bf16())


def test_device_generator_equals_host_generator(L, tiny):
    args, tensors, om, gm = tiny
    g2 = L.model.LoadSyntheticModel(args, seed=SEED)        # generated directly in HBM
    try:
        toks = np.array([3, 77, 1000, 5, 9], np.int32)
        c1 = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(16), acc_mode=L._capi.LNB_ACC_STRICT)
        c2 = L.model.InferenceContext(g2.Transformer, L.model.InferenceArgs(16), acc_mode=L._capi.LNB_ACC_STRICT)
        l1 = gm.Transformer.Forward(c1, L.ml.Tensor(toks, L.ml.DT_INT32), 0).RawData
        l2 = g2.Transformer.Forward(c2, L.ml.Tensor(toks, L.ml.DT_INT32), 0).RawData
        assert np.array_equal(l1, l2)
        c1.close(); c2.close()
    finally:
        g2.Free()
    # and the host twin inside the product library
    name = "layers.1.feed_forward.w2.weight"
    assert np.array_equal(L.synth.fill_host(args, name, SEED), tensors[name])


@pytest.mark.parametrize("prompt_len", [1, 5, 8])
def test_forward_strict_bit_exact(L, tiny, prompt_len):
    args, _, om, gm = tiny
    rng = np.random.default_rng(prompt_len)
    seq = 24
    toks = rng.integers(0, args["vocab_size"], size=seq).astype(np.int32)
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), acc_mode=L._capi.LNB_ACC_STRICT)
    osess = om.new_session(seq)
    try:
        # prefill
        exp, tr = osess.forward(toks[:prompt_len], 0, all_rows=True, trace=True)
        got = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[:prompt_len], L.ml.DT_INT32), 0).RawData
        assert np.array_equal(ctx.residual(prompt_len), tr[-1]), "residual stream after the last layer differs"
        assert np.array_equal(got, exp)
        # decode steps
        for pos in range(prompt_len, prompt_len + 6):
            exp = osess.forward(toks[pos:pos + 1], pos)
            got = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[pos:pos + 1], L.ml.DT_INT32), pos).RawData
            assert np.array_equal(got, exp), f"logits differ at position {pos}"
        for layer in range(args["n_layers"]):
            ok, ov = osess.cache(layer)
            assert np.array_equal(ctx.CacheK(layer).RawData, ok)
            assert np.array_equal(ctx.CacheV(layer).RawData, ov)
    finally:
        ctx.close(); osess.close()


def test_forward_per_layer_trace_strict(L, tiny):
    args, _, om, gm = tiny
    toks = np.array([1, 2, 3], np.int32)
    osess = om.new_session(8)
    _, tr = osess.forward(toks, 0, trace=True)
    osess.close()
    for n in range(1, args["n_layers"] + 1):
        ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(8), acc_mode=L._capi.LNB_ACC_STRICT)
        ctx.set_layer_limit(n)
        gm.Transformer.forward_argmax(ctx, toks, 0)
        assert np.array_equal(ctx.residual(3), tr[n]), f"layer {n}"
        ctx.close()


def test_forward_fast_within_tolerance(L, tiny):
    args, _, om, gm = tiny
    rng = np.random.default_rng(11)
    seq = 24
    toks = rng.integers(0, args["vocab_size"], size=seq).astype(np.int32)
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), acc_mode=L._capi.LNB_ACC_FAST)
    osess = om.new_session(seq)
    try:
        exp = osess.forward(toks[:8], 0)
        got = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[:8], L.ml.DT_INT32), 0).RawData
        assert np.abs(got - exp).max() <= 1e-2                      # north_star tolerance
        for pos in range(8, 16):
            exp = osess.forward(toks[pos:pos + 1], pos)
            got = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[pos:pos + 1], L.ml.DT_INT32), pos).RawData
            assert np.abs(got - exp).max() <= 1e-2
    finally:
        ctx.close(); osess.close()


def test_forward_errors_like_reference(L, tiny):
    args, _, _, gm = tiny
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(8))
    T = L.ml.Tensor
    with pytest.raises(L.ml.MlError, match="empty token array"):
        gm.Transformer.Forward(ctx, T(np.zeros(0, np.int32), L.ml.DT_INT32), 0)
    with pytest.raises(L._capi.LnbError, match="SequenceLength"):
        gm.Transformer.Forward(ctx, T(np.zeros(1, np.int32), L.ml.DT_INT32), 8)
    with pytest.raises(L._capi.LnbError, match="out of range"):
        gm.Transformer.Forward(ctx, T(np.array([args["vocab_size"]], np.int32), L.ml.DT_INT32), 0)
    with pytest.raises(L._capi.LnbError, match="startPos 0"):
        gm.Transformer.Forward(ctx, T(np.zeros(2, np.int32), L.ml.DT_INT32), 3)
    ctx.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_generate_loop_matches_oracle(L, tiny, mode):
    args, _, om, gm = tiny
    acc = L._capi.LNB_ACC_STRICT if mode == "strict" else L._capi.LNB_ACC_FAST
    prompt = [1, 50, 999, 7, 300, 12, 64, 2]
    seq = 40
    exp = om.generate(prompt, seq, stop_ids=(10**9,))
    eng = L.inference.InferenceEngine(gm, L.model.InferenceArgs(seq), acc_mode=acc)
    got_ref_api = [t for _, t in eng.GenerateTokens(prompt, use_reference_api=True)]
    got_fused = [t for _, t in eng.GenerateTokens(prompt, use_reference_api=False)]
    assert got_ref_api == got_fused
    assert len(got_fused) == seq - len(prompt)
    if mode == "strict":
        assert got_fused == list(exp)
    else:
        n = next((i for i, (a, b) in enumerate(zip(got_fused, exp)) if a != b), len(exp))
        assert n >= 1  # first token must agree; the full-size statistics live in test_gpu_8b.py


def test_generate_stops_on_eos(L, tiny):
    args, _, om, gm = tiny
    prompt = [1, 50, 999]
    free = list(om.generate(prompt, 20, stop_ids=(10**9,)))
    stop = free[4]
    gm.Vocabulary.StopTokenIds = (stop,)
    try:
        eng = L.inference.InferenceEngine(gm, L.model.InferenceArgs(20), acc_mode=L._capi.LNB_ACC_STRICT)
        out = list(eng.GenerateTokens(prompt))
        k = free.index(stop)
        assert [t for _, t in out] == free[:k + 1]
        assert out[-1][0] == L.inference.GSFinishedByReachingEOS
        assert list(om.generate(prompt, 20, stop_ids=(stop,))) == free[:k + 1]
    finally:
        gm.Vocabulary.StopTokenIds = L.synth.STOP_IDS


@pytest.mark.parametrize("use_graph", [False, True])
def test_device_resident_decode_equals_host_loop(L, tiny, use_graph):
    args, _, om, gm = tiny
    prompt = [1, 50, 999, 7, 300, 12, 64, 2]
    seq = 40
    exp = list(om.generate(prompt, seq, stop_ids=(10**9,)))
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), acc_mode=L._capi.LNB_ACC_STRICT)
    try:
        first, _ = gm.Transformer.forward_argmax(ctx, np.array(prompt, np.int32), 0)
        assert first == exp[0]
        toks, ms, graphed = ctx.decode_run(first, len(prompt), seq - len(prompt) - 1, use_graph=use_graph)
        assert list(toks) == exp[1:]
        assert ms > 0
        assert graphed == use_graph
        assert ctx.launch_count() > 0
    finally:
        ctx.close()


def test_prefill_tensor_core_path(L, tiny):
    """LNB_ACC_FAST with a prompt of >= 32 tokens runs the tcgen05 prefill path (gemm_tc.cuh + X8
    elementwise kernels); it must agree with the oracle up to fp32 summation order, and the KV cache it
    leaves must drive the ordinary decode path."""
    args, _, om, gm = tiny
    rng = np.random.default_rng(40)
    S, seq = 40, 48
    toks = rng.integers(0, args["vocab_size"], size=seq).astype(np.int32)
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), max_rows=S, acc_mode=L._capi.LNB_ACC_FAST)
    osess = om.new_session(seq)
    try:
        exp, tr = osess.forward(toks[:S], 0, all_rows=True, trace=True)
        nxt, got = gm.Transformer.forward_argmax(ctx, toks[:S], 0, want_logits="all")
        assert got.shape == exp.shape
        assert np.abs(got - exp).max() <= 1e-2
        assert (np.argmax(got, -1) == np.argmax(exp, -1)).mean() >= 0.95
        assert nxt == int(np.argmax(got[-1]))
        from tests.helpers import f32
        r_got, r_exp = f32(ctx.residual(S)).reshape(S, -1), f32(tr[-1]).reshape(S, -1)
        assert np.abs(r_got - r_exp).max() <= 2.0 ** -5 * max(1.0, np.abs(r_exp).max())   # a few bf16 ulps of the largest entry
        assert (r_got != r_exp).mean() < 0.3
        for layer in range(args["n_layers"]):
            ok, ov = osess.cache(layer)
            for got_c, exp_c in ((ctx.CacheK(layer).RawData[:S], ok[:S]), (ctx.CacheV(layer).RawData[:S], ov[:S])):
                g_, e_ = f32(got_c).reshape(-1), f32(exp_c).reshape(-1)
                assert np.abs(g_ - e_).max() <= 2.0 ** -6 * max(1.0, np.abs(e_).max())
        for pos in range(S, S + 4):     # decode continues on the cache written by the prefill kernels
            e = osess.forward(toks[pos:pos + 1], pos)
            g = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[pos:pos + 1], L.ml.DT_INT32), pos).RawData
            assert np.abs(g - e).max() <= 1e-2
    finally:
        ctx.close(); osess.close()


def test_prefill_strict_stays_bit_exact_for_long_prompts(L, tiny):
    args, _, om, gm = tiny
    rng = np.random.default_rng(41)
    S = 36
    toks = rng.integers(0, args["vocab_size"], size=S).astype(np.int32)
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(40), max_rows=S, acc_mode=L._capi.LNB_ACC_STRICT)
    osess = om.new_session(40)
    try:
        exp = osess.forward(toks, 0, all_rows=True)
        got = gm.Transformer.Forward(ctx, L.ml.Tensor(toks, L.ml.DT_INT32), 0).RawData
        assert np.array_equal(got, exp)
    finally:
        ctx.close(); osess.close()


def test_cpp_host_mirror_generates_the_oracle_tokens(L):
    """host/lnb_generate: the C++ mirror of src/ml + src/model + src/inference (reference-shaped loop:
    Forward -> logits to host -> Slice -> ml.Argmax) over the C-ABI, on the device-generated tiny model."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "host", "lnb_generate")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "host"), "-s"])
    out = subprocess.run([exe, "24", "strict", "tiny"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("tokens:")][0]
    got = [int(t) for t in line.split()[1:]]
    args = dict(L.synth.TINY)
    om = oracle_model(args, host_tensors(args, 7))
    exp = list(om.generate([1, 50, 999, 7, 300, 12, 64, 2], 24, stop_ids=(10**9,)))
    om.close()
    assert got == exp
    # the same through model.LoadModel: `--write-synthetic` writes params.json + consolidated.00.pth (host-only),
    # the second run maps that file and uploads it
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([exe, "--write-synthetic", d, "tiny"])
        out = subprocess.run([exe, "24", "strict", d], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        line = [l for l in out.stdout.splitlines() if l.startswith("tokens:")][0]
        assert [int(t) for t in line.split()[1:]] == exp
    # inference::GenerateTokensBatch: 3 prompts generated together, each gets its single-prompt stream
    out = subprocess.run([exe, "24", "strict", "tiny", "batch=3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    base = [1, 50, 999, 7, 300, 12, 64, 2]
    om = oracle_model(args, host_tensors(args, 7))
    for b in range(3):
        prompt = [base[0]] + [(t + 977 * b) % 1000 for t in base[1:]]
        line = [l for l in out.stdout.splitlines() if l.startswith(f"tokens[{b}]:")][0]
        assert [int(t) for t in line.split()[1:]] == list(om.generate(prompt, 24, stop_ids=(10**9,))), b
    om.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_batched_decode_equals_independent_contexts(L, tiny, mode):
    """config 5: n_seq sequences stepped together through one pass over the weights; STRICT: every sequence is
    bit-identical to its own oracle InferenceContext (different prompt lengths, different positions)."""
    args, _, om, gm = tiny
    acc = L._capi.LNB_ACC_STRICT if mode == "strict" else L._capi.LNB_ACC_FAST
    rng = np.random.default_rng(55)
    seq, nseq = 32, 5
    prompts = [rng.integers(0, args["vocab_size"], size=k).astype(np.int32) for k in (3, 8, 1, 6, 8)]
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), max_rows=8, acc_mode=acc, n_seq=nseq)
    osess = [om.new_session(seq) for _ in range(nseq)]
    try:
        cur, pos = [], []
        for i, p in enumerate(prompts):                      # prefill every sequence through the ordinary Forward
            ctx.set_active_sequence(i)
            nxt, lg = gm.Transformer.forward_argmax(ctx, p, 0, want_logits="last")
            exp = osess[i].forward(p, 0, all_rows=False)
            if mode == "strict":
                assert np.array_equal(lg, exp)
            cur.append(int(np.argmax(exp[0]))); pos.append(len(p))
            assert mode != "strict" or nxt == cur[-1]
        for step in range(6):
            nxt, lg = ctx.forward_batch(cur, pos, want_logits=True)
            for i in range(nseq):
                exp = osess[i].forward(np.array([cur[i]], np.int32), pos[i])
                if mode == "strict":
                    assert np.array_equal(lg[i], exp[0]), f"sequence {i} step {step}"
                    assert nxt[i] == O.argmax_f32(exp[0])
                else:
                    assert np.abs(lg[i] - exp[0]).max() <= 1e-2
                cur[i] = int(np.argmax(exp[0])); pos[i] += 1   # teacher-forced with the oracle's token
        if mode == "strict":
            for i in range(nseq):                            # caches of every sequence, all layers
                ctx.set_active_sequence(i)
                for layer in range(args["n_layers"]):
                    ok, ov = osess[i].cache(layer)
                    assert np.array_equal(ctx.CacheK(layer).RawData[:pos[i]], ok[:pos[i]])
                    assert np.array_equal(ctx.CacheV(layer).RawData[:pos[i]], ov[:pos[i]])
        with pytest.raises(L._capi.LnbError):
            ctx.forward_batch([1] * (nseq + 1), [0] * (nseq + 1))
        with pytest.raises(L._capi.LnbError):
            ctx.set_active_sequence(nseq)
    finally:
        ctx.close()
        for o in osess:
            o.close()


def test_load_model_from_checkpoint_directory(L, tiny, tmp_path):
    """model.LoadModel(modelDir) -- params.json + consolidated.00.pth go from the file mapping straight
    to HBM (lnb_model_load_pth) and the model is bit-identical to the one built from host tensors / the oracle"""
    import json
    from lnb_b200.torch_reader import TorchModelWriter, write_synthetic_checkpoint
    args, tensors, om, gm = tiny
    write_synthetic_checkpoint(str(tmp_path / "consolidated.00.pth"), L.synth.args_c(args), SEED)
    # a params.json that derives TINY's widths the reference's way: head_dim = dim / n_heads,
    # ffn = multiple_of * ceil(int(mult * int(2 * 4 dim / 3)) / multiple_of) = 256 * ceil(511 / 256) = 512
    (tmp_path / "params.json").write_text(json.dumps({
        "dim": args["dim"], "n_layers": args["n_layers"], "n_heads": args["n_heads"], "n_kv_heads": args["n_kv_heads"],
        "ffn_dim_multiplier": 0.75, "multiple_of": 256, "norm_eps": 1e-05, "rope_theta": 500000.0, "use_scaled_rope": True}))
    d = L.model.load_model_args(str(tmp_path), max_seq_len=args["max_seq_len"])
    assert d["ffn_dim"] == args["ffn_dim"] and d["head_dim"] == args["head_dim"] and d["vocab_size"] == -1
    m = L.model.LoadModel(str(tmp_path), max_seq_len=args["max_seq_len"])     # vocab from tok_embeddings rows
    try:
        assert m.ModelArgs.VocabSize == args["vocab_size"]
        toks = np.array([3, 77, 1000, 5, 9], np.int32)
        ctx = L.model.InferenceContext(m.Transformer, L.model.InferenceArgs(16), acc_mode=L._capi.LNB_ACC_STRICT)
        so = om.new_session(16)
        got = m.Transformer.Forward(ctx, L.ml.Tensor(toks, L.ml.DT_INT32), 0).RawData
        assert np.array_equal(got, so.forward(toks, 0))
        nxt, lg = m.Transformer.forward_argmax(ctx, np.array([int(np.argmax(got[-1]))], np.int32), 5, want_logits="last")
        assert np.array_equal(lg, so.forward(np.array([int(np.argmax(got[-1]))], np.int32), 5))
        ctx.close(); so.close()
    finally:
        m.Free()
    # a checkpoint with a missing tensor / a wrong shape / a non-bf16 tensor is refused the reference's way
    def variant(mutate):
        w = TorchModelWriter(str(tmp_path / "consolidated.00.pth"))
        for name, arr in tensors.items():
            r = mutate(name, arr)
            if r is not None:
                w.Add(name, *r)
        w.Finish()
    variant(lambda n, a: None if n == "layers.1.ffn_norm.weight" else (a,))
    with pytest.raises(L._capi.LnbError) as e:
        L.model.LoadModel(str(tmp_path), max_seq_len=args["max_seq_len"])
    assert "missing" in str(e.value)
    variant(lambda n, a: (a[:-1],) if n == "norm.weight" else (a,))
    with pytest.raises(L._capi.LnbError) as e:
        L.model.LoadModel(str(tmp_path), max_seq_len=args["max_seq_len"])
    assert "norm.weight" in str(e.value) and "shape" in str(e.value)
    variant(lambda n, a: (a.astype(np.float32),) if n == "norm.weight" else (a,))
    with pytest.raises(L._capi.LnbError) as e:
        L.model.LoadModel(str(tmp_path), max_seq_len=args["max_seq_len"])
    assert "float32" in str(e.value)


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_generate_tokens_batch_equals_per_prompt_generation(L, tiny, mode):
    """the batched generate loop -- every prompt gets exactly the stream the single-prompt loop of
    generateTokensInternal gives it (oracle: orc_generate), with different prompt lengths and one sequence that
    stops early on an EOS id while the others run to SequenceLength"""
    args, _, om, gm = tiny
    prompts = [[1, 50, 999, 7, 300, 12, 64, 2], [5, 6, 7], [1000], [3, 3, 3, 3, 3]]
    seq = 24
    free = [list(om.generate(p, seq, stop_ids=(10**9,))) for p in prompts]
    stop = int(free[1][4])                                   # sequence 1 will hit it at its 5th token at the latest
    exp = [list(om.generate(p, seq, stop_ids=(stop,))) for p in prompts]
    assert len(exp[1]) <= 5 and any(len(e) == seq - len(p) for e, p in zip(exp, prompts))
    saved = gm.Vocabulary.StopTokenIds
    gm.Vocabulary.StopTokenIds = (stop,)
    try:
        acc = L._capi.LNB_ACC_STRICT if mode == "strict" else L._capi.LNB_ACC_FAST
        eng = L.inference.InferenceEngine(gm, L.model.InferenceArgs(seq), acc_mode=acc)
        got, last_state = [[] for _ in prompts], {}
        for i, state, tok in eng.GenerateTokensBatch(prompts):
            assert last_state.get(i, L.inference.GSInProgress) == L.inference.GSInProgress      # nothing after the end
            got[i].append(tok); last_state[i] = state
        if mode == "strict":
            assert got == exp
            for i, e in enumerate(exp):
                assert last_state[i] == (L.inference.GSFinishedByReachingEOS if e[-1] == stop else L.inference.GSFinishedByReachingSeqLen)
        else:   # FAST may flip a near-tie; the streams must at least be complete and well-formed
            assert all(len(g) >= 1 and s != L.inference.GSInProgress for g, s in zip(got, last_state.values()))
        one = [tok for _, _, tok in eng.GenerateTokensBatch([prompts[2]])]                       # n = 1 delegates
        assert mode != "strict" or one == exp[2]
        with pytest.raises(L.ml.MlError):
            list(eng.GenerateTokensBatch([list(range(seq))]))
    finally:
        gm.Vocabulary.StopTokenIds = saved


def test_captured_decode_graph_follows_active_sequence_and_layer_limit(L, tiny):
    """round-1 advisor finding: the decode graph bakes in the active sequence's cache pointers and the layer count;
    changing either must re-capture, not silently replay the first capture"""
    args, _, om, gm = tiny
    seq = 32
    prompts = [[1, 50, 999, 7], [3, 4, 5, 6, 7, 8]]
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), max_rows=8, acc_mode=L._capi.LNB_ACC_STRICT, n_seq=2)
    try:
        firsts = []
        for i, p in enumerate(prompts):
            ctx.set_active_sequence(i)
            first, _ = gm.Transformer.forward_argmax(ctx, np.array(p, np.int32), 0)
            firsts.append(first)
        for i, p in enumerate(prompts):                      # graph captured on sequence 0, then used on sequence 1
            ctx.set_active_sequence(i)
            toks, _, graphed = ctx.decode_run(firsts[i], len(p), 8, use_graph=True)
            assert graphed
            exp = list(om.generate(p, len(p) + 9, stop_ids=(10**9,)))
            assert [firsts[i]] + list(toks) == exp, f"sequence {i}"
        ctx.set_active_sequence(0)
        ctx.set_layer_limit(1)                               # one layer only: the graph path must agree with the eager path
        a, _, _ = ctx.decode_run(firsts[0], len(prompts[0]), 4, use_graph=True)
        b, _, _ = ctx.decode_run(firsts[0], len(prompts[0]), 4, use_graph=False)
        assert list(a) == list(b)
        ctx.set_layer_limit(0)
        c, _, _ = ctx.decode_run(firsts[0], len(prompts[0]), 8, use_graph=True)
        assert [firsts[0]] + list(c) == list(om.generate(prompts[0], len(prompts[0]) + 9, stop_ids=(10**9,)))
    finally:
        ctx.close()


def test_forward_returns_a_device_handle_that_reads_like_the_host_tensor(L, tiny):
    """Transformer.Forward keeps the [S, vocab] logits in HBM (ml.DeviceLogits): Size / Slice / ml.Argmax work on the handle,
    RawData copies the rows out, and the handle dies with the context's next Forward (inference.go:202-216 never needs more)"""
    args, _, om, gm = tiny
    toks = np.array([1, 50, 999, 7, 300], np.int32)
    ctx = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(16), acc_mode=L._capi.LNB_ACC_STRICT)
    osess = om.new_session(16)
    try:
        exp = osess.forward(toks, 0, all_rows=True)
        lg = gm.Transformer.Forward(ctx, L.ml.Tensor(toks, L.ml.DT_INT32), 0)
        assert isinstance(lg, L.ml.DeviceLogits) and lg.Size == [5, args["vocab_size"]] and lg.on_device()
        last = lg.Slice([4], [5])
        assert last.Size == [1, args["vocab_size"]] and last.on_device()
        assert int(L.ml.Argmax(last, 1).Item()) == O.argmax_f32(exp[4])          # fused argmax of the kept last row
        mid = lg.Slice([1], [4])
        assert L.ml.Argmax(mid, 1).RawData.tolist() == [O.argmax_f32(exp[r]) for r in (1, 2, 3)]   # device argmax of kept rows
        assert np.array_equal(mid.RawData, exp[1:4]) and not mid.on_device()      # materialised on first touch
        assert np.array_equal(lg.RawData, exp)
        lg2 = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[:1] + 1, L.ml.DT_INT32), 5)
        stale = lg.Slice([0], [1])                                                # lg itself was read: still a host tensor
        assert np.array_equal(stale.RawData, exp[0:1])
        lg3 = gm.Transformer.Forward(ctx, L.ml.Tensor(toks[:1] + 2, L.ml.DT_INT32), 6)
        with pytest.raises(L.ml.MlError):
            lg2.RawData                                                           # never read, and a newer Forward ran
        assert lg3.Size == [1, args["vocab_size"]]
    finally:
        ctx.close()
        osess.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_chunked_prefill_extension_equals_one_shot_prefill(L, tiny, mode):
    """SURVEY 8f-4 (beyond the reference): S > 1 at startPos > 0 with the [S,T] causal mask.  The library refuses it by
    default exactly like the reference's [S,S] mask would fail; with the switch on, a prompt fed in chunks leaves the caches
    and logits of the one-shot prefill (bit-identical in STRICT, and equal to the oracle's extension)."""
    args, _, om, gm = tiny
    acc = L._capi.LNB_ACC_STRICT if mode == "strict" else L._capi.LNB_ACC_FAST
    rng = np.random.default_rng(8)
    prompt = rng.integers(0, args["vocab_size"], size=19).astype(np.int32)
    seq = 40
    one = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), max_rows=24, acc_mode=acc)
    chk = L.model.InferenceContext(gm.Transformer, L.model.InferenceArgs(seq), max_rows=24, acc_mode=acc)
    o1, o2 = om.new_session(seq), om.new_session(seq)
    try:
        with pytest.raises(L._capi.LnbError, match="startPos 0"):
            gm.Transformer.forward_argmax(chk, prompt[8:12], 8)
        chk.allow_chunked_prefill(True)
        n1, l1 = gm.Transformer.forward_argmax(one, prompt, 0, want_logits="last")
        exp = o1.forward(prompt, 0, all_rows=False)
        pos = 0
        for size in (8, 5, 6):                                  # 8 + 5 + 6 = 19, unequal chunks
            n2, l2 = gm.Transformer.forward_argmax(chk, prompt[pos:pos + size], pos, want_logits="all")
            e2 = o2.forward(prompt[pos:pos + size], pos, all_rows=True) if pos == 0 else o2.forward_chunk(prompt[pos:pos + size], pos)
            if mode == "strict":
                assert np.array_equal(l2, e2), f"chunk at {pos}"
            else:
                assert float(np.abs(l2 - e2).max()) <= 1e-2
            pos += size
        if mode == "strict":
            assert np.array_equal(l2[-1:], l1) and np.array_equal(l1, exp) and n1 == n2      # chunked == one-shot == reference
            for layer in range(args["n_layers"]):
                assert np.array_equal(chk.CacheK(layer).RawData[:19], one.CacheK(layer).RawData[:19])
                assert np.array_equal(chk.CacheV(layer).RawData[:19], one.CacheV(layer).RawData[:19])
        # the generate loop with a chunked prompt emits the same stream (a prompt longer than max_rows becomes possible)
        eng = L.inference.InferenceEngine(gm, L.model.InferenceArgs(seq), acc_mode=acc, max_rows=8)
        got = [t for _, t in eng.GenerateTokens(list(prompt), prefill_chunk=8)]
        if mode == "strict":
            assert got == list(om.generate(prompt, seq, stop_ids=gm.Vocabulary.StopTokenIds))
    finally:
        one.close(); chk.close(); o1.close(); o2.close()
