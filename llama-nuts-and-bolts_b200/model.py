"""Host mirror of the reference's src/model package for the forward path:
ModelArgs (modelargs.go), Model / LlamaTransformer (llamatransformer.go:16-113,145-180) and
InferenceContext (inferencecontext.go).  The transformer's weights, KV cache and activations
live in HBM behind liblnb.so; `LlamaTransformer.Forward` is ONE call through the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi, ml, synth
from ._capi import LNB_ACC_FAST, LNB_ACC_STRICT, check, lib, ptr


@dataclass
class ModelArgs:  # src/model/modelargs.go:10-44 (+ derived HeadDim / FFN width)
    Dim: int = 4096
    N_Layers: int = 32
    N_Heads: int = 32
    N_KVHeads: int = 8
    VocabSize: int = 128256
    MultipleOf: int = 1024
    FFNDimMultiplier: float = 1.3
    NormEpsilon: float = 1e-5
    RopeTheta: float = 500000.0
    UseScaledRope: bool = True
    MaxSequenceLength: int = 2048
    HeadDim: int = field(default=0)
    FFNDim: int = field(default=0)

    def __post_init__(self):
        if not self.HeadDim:
            self.HeadDim = self.Dim // self.N_Heads  # llamatransformer.go:73
        if not self.FFNDim:  # llamatransformer.go:569-577
            hidden = int(2 * (4 * self.Dim) / 3)
            hidden = int(self.FFNDimMultiplier * hidden)
            self.FFNDim = self.MultipleOf * ((hidden + self.MultipleOf - 1) // self.MultipleOf)

    def to_c_dict(self) -> dict:
        return dict(dim=self.Dim, n_layers=self.N_Layers, n_heads=self.N_Heads, n_kv_heads=self.N_KVHeads,
                    head_dim=self.HeadDim, ffn_dim=self.FFNDim, vocab_size=self.VocabSize,
                    max_seq_len=self.MaxSequenceLength, norm_eps=self.NormEpsilon, rope_theta=self.RopeTheta,
                    use_scaled_rope=int(self.UseScaledRope))

    @classmethod
    def from_c_dict(cls, d: dict) -> "ModelArgs":
        return cls(Dim=d["dim"], N_Layers=d["n_layers"], N_Heads=d["n_heads"], N_KVHeads=d["n_kv_heads"],
                   VocabSize=d["vocab_size"], NormEpsilon=d["norm_eps"], RopeTheta=d["rope_theta"],
                   UseScaledRope=bool(d["use_scaled_rope"]), MaxSequenceLength=d["max_seq_len"],
                   HeadDim=d["head_dim"], FFNDim=d["ffn_dim"])


@dataclass
class Vocabulary:  # src/model/vocabulary.go (ids only; the tokenizer is out of scope)
    PadId: int = synth.PAD_ID
    StopTokenIds: tuple = synth.STOP_IDS


class LlamaTransformer:
    """model.LlamaTransformer: owns the device-side weights of one tensor-parallel rank."""

    def __init__(self, args: ModelArgs, device: int = 0, tp_rank: int = 0, tp_size: int = 1, nccl_id: bytes | None = None):
        self.args = args
        self.device, self.tp_rank, self.tp_size = device, tp_rank, tp_size
        self._cargs = _capi.ModelArgsC(**args.to_c_dict())
        h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        check(lib.lnb_model_create(C.byref(self._cargs), device, tp_rank, tp_size, idbuf, C.byref(h)))
        self.h = h
        self.finalized = False

    # --- weights -----------------------------------------------------------------------------
    def upload_tensor(self, name: str, data: np.ndarray):
        """getTensor + device upload (src/model/loader.go:183-197): `data` is the full bf16 tensor."""
        data = np.ascontiguousarray(data, np.uint16)
        shape = (C.c_int64 * data.ndim)(*data.shape)
        check(lib.lnb_model_upload_tensor(self.h, name.encode(), ptr(data, _capi.u16p), shape, data.ndim))

    def init_synthetic(self, seed: int = synth.SEED):
        check(lib.lnb_model_init_synthetic(self.h, seed))

    def load_pth(self, path: str) -> int:
        """every tensor the architecture names, from the checkpoint mapping straight to HBM (lnb_model_load_pth);
        returns how many were uploaded"""
        n = C.c_int(0)
        check(lib.lnb_model_load_pth(self.h, path.encode(), C.byref(n)))
        return n.value

    def finalize(self):
        check(lib.lnb_model_finalize(self.h))
        self.finalized = True

    def rope_table(self, rows: int | None = None) -> np.ndarray:
        rows = rows or self.args.MaxSequenceLength * 2
        out = np.empty((rows, self.args.HeadDim // 2, 2), np.float32)
        check(lib.lnb_model_get_rope_table(self.h, ptr(out, _capi.f32p), rows))
        return out

    def silu_table(self) -> np.ndarray:
        out = np.empty(65536, np.uint16)
        check(lib.lnb_model_get_silu_table(self.h, ptr(out, _capi.u16p)))
        return out

    # --- forward -----------------------------------------------------------------------------
    def Forward(self, infContext: "InferenceContext", inputTokens: ml.Tensor, startPos: int) -> ml.Tensor:
        """LlamaTransformer.Forward (llamatransformer.go:145-180): returns f32 logits [S, vocab]."""
        if inputTokens.DataType is not ml.DT_INT32 or inputTokens.RawData.ndim != 1:
            raise ml.MlError("inputTokens must be a 1-D Int32 tensor")
        S = inputTokens.RawData.shape[0]
        if S == 0:
            raise ml.MlError("empty token array")
        # the [S, vocab] f32 result stays in HBM behind a handle; reading it (RawData) copies it to the host
        gen = C.c_int64(0)
        check(lib.lnb_forward_device(infContext.h, ptr(inputTokens.RawData, _capi.i32p), S, startPos, S, None, C.byref(gen)))
        infContext.generation = gen.value
        return ml.DeviceLogits(infContext, gen.value, 0, S, self.args.VocabSize)

    def forward_argmax(self, infContext: "InferenceContext", tokens, startPos: int, want_logits: str = "none"):
        """Fast path of the generate loop: forward + last-row argmax in one call.
        want_logits: "none" | "last" | "all"."""
        tokens = np.ascontiguousarray(tokens, np.int32)
        S = tokens.shape[0]
        nxt = C.c_int32(-1)
        logits = None
        if want_logits != "none":
            logits = np.empty((S if want_logits == "all" else 1, self.args.VocabSize), np.float32)
        check(lib.lnb_forward(infContext.h, ptr(tokens, _capi.i32p), S, startPos,
                              ptr(logits, _capi.f32p) if logits is not None else None,
                              1 if want_logits == "all" else 0, C.byref(nxt)))
        return nxt.value, logits

    def close(self):
        if self.h:
            lib.lnb_model_destroy(self.h)
            self.h = None


@dataclass
class InferenceArgs:  # src/common/inferenceargs.go:3-11
    SequenceLength: int = 0


class InferenceContext:
    """model.InferenceContext (inferencecontext.go:8-46): the KV cache (here: in HBM)."""

    def __init__(self, transformer: LlamaTransformer, inferenceArgs: InferenceArgs, logFn=None,
                 max_rows: int = 8, acc_mode: int = LNB_ACC_FAST, n_seq: int = 1):
        self.transformer = transformer
        self.SequenceLength = inferenceArgs.SequenceLength if inferenceArgs.SequenceLength > 0 \
            else transformer.args.MaxSequenceLength
        self.logFn = logFn
        self.acc_mode = acc_mode
        self.max_rows = max_rows
        self.n_seq = n_seq
        h = C.c_void_p()
        if n_seq > 1:   # n_seq reference InferenceContexts that share the weights (batched decode)
            check(lib.lnb_session_create_batch(transformer.h, self.SequenceLength, n_seq, max_rows, acc_mode, C.byref(h)))
        else:
            check(lib.lnb_session_create(transformer.h, self.SequenceLength, max_rows, acc_mode, C.byref(h)))
        self.h = h
        self.generation = 0        # forward calls made through lnb_forward_device (validity of ml.DeviceLogits handles)
        self.pre_close_hook = None  # tensor-parallel sessions: barrier before the peer-mapped region is freed

    def Logf(self, fmt, *v):
        if self.logFn:
            self.logFn(fmt, *v)

    # CacheK / CacheV of the reference are host tensors; here they are read back on demand
    def CacheK(self, layer: int) -> ml.Tensor:
        return self._cache(_capi.LNB_BUF_CACHE_K, layer)

    def CacheV(self, layer: int) -> ml.Tensor:
        return self._cache(_capi.LNB_BUF_CACHE_V, layer)

    def _cache(self, which, layer):
        a = self.transformer.args
        kv_l = a.N_KVHeads // self.transformer.tp_size
        out = np.empty((self.SequenceLength, kv_l, a.HeadDim), np.uint16)
        check(lib.lnb_session_read(self.h, which, layer, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return ml.Tensor(out, ml.DT_BF16)

    def residual(self, rows: int) -> np.ndarray:
        out = np.empty((rows, self.transformer.args.Dim), np.uint16)
        check(lib.lnb_session_read(self.h, _capi.LNB_BUF_RESIDUAL, 0, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_active_sequence(self, seq: int):
        check(lib.lnb_session_set_active_sequence(self.h, seq))

    def forward_batch(self, tokens, positions, want_logits: bool = False):
        """one decode step for len(tokens) independent sequences: (greedy tokens [n], logits [n, vocab] or None)"""
        tokens = np.ascontiguousarray(tokens, np.int32)
        positions = np.ascontiguousarray(positions, np.int32)
        n = tokens.shape[0]
        nxt = np.empty(n, np.int32)
        logits = np.empty((n, self.transformer.args.VocabSize), np.float32) if want_logits else None
        check(lib.lnb_forward_batch(self.h, ptr(tokens, _capi.i32p), ptr(positions, _capi.i32p), n,
                                    ptr(logits, _capi.f32p) if want_logits else None, ptr(nxt, _capi.i32p)))
        return nxt, logits

    def allow_chunked_prefill(self, on: bool = True):
        """EXTENSION (not in the reference): accept S > 1 at startPos > 0 with the [S,T] causal mask"""
        check(lib.lnb_session_set_chunked_prefill(self.h, int(on)))

    def set_layer_limit(self, n: int):
        check(lib.lnb_session_set_layer_limit(self.h, n))

    def launch_count(self) -> int:
        return int(lib.lnb_session_launch_count(self.h))

    def decode_run(self, first_token: int, start_pos: int, n_steps: int, use_graph: bool = True):
        """device-resident greedy decode (lnb_decode_run): returns (tokens, ms, used_graph)"""
        toks = np.empty(n_steps, np.int32)
        ms = C.c_float(0)
        rc = check(lib.lnb_decode_run(self.h, first_token, start_pos, n_steps, int(use_graph),
                                      ptr(toks, _capi.i32p), C.byref(ms)))
        return toks, ms.value, bool(rc)

    def enable_peer_allreduce(self, all_gather_bytes):
        """Switch this session from NCCL to the fused peer-memory all-reduce.  `all_gather_bytes(b: bytes) ->
        list[bytes]` must return every rank's contribution ordered by rank (e.g. torch.distributed)."""
        buf = C.create_string_buffer(64)
        check(lib.lnb_session_p2p_export(self.h, buf))
        handles = all_gather_bytes(buf.raw)
        blob = b"".join(handles)
        check(lib.lnb_session_p2p_import(self.h, C.create_string_buffer(blob, len(blob)), len(handles)))

    def bench_kernel(self, kind: int, reps: int = 3):
        """(ms per launch, algorithmic weight bytes per launch, launches timed) -- bench.py roofline"""
        ms, nb, nl = C.c_float(0), C.c_int64(0), C.c_int32(0)
        check(lib.lnb_session_bench_kernel(self.h, kind, reps, C.byref(ms), C.byref(nb), C.byref(nl)))
        return ms.value, nb.value, nl.value

    def close(self):
        if self.h:
            if self.pre_close_hook is not None:
                self.pre_close_hook(self)   # peers may still be storing into this session's all-reduce region
            lib.lnb_session_destroy(self.h)
            self.h = None

    def uses_engine(self) -> bool:
        """True: S=1 steps of this context run as the persistent decode engine (one launch per decode_run / Forward)"""
        return bool(check(lib.lnb_session_decode_engine(self.h)))

    def engine_profile(self, batch=False):
        """LNB_ENGINE_PROF=1: {section: (mean cycles, max cycles)} of the decode engine's consumer thread 0 since the last call
        (batch=True: the sections of the batch engine, engine_batch.cuh BatchParams.prof)"""
        out = (C.c_double * 24)()
        check(lib.lnb_session_engine_profile(self.h, out))
        names = ["grid_barrier", "prologue", "main_loop", "epilogue", "attention", "peer_reduce", "input_wait", "scan_maps", "scan_scans",
                 "scan_walk", "scan_count"]
        if batch:
            names = ["grid_barrier", "setup", "x_tile|mma:acc_wait", "stage_wait", "fma|mma:w_wait", "epilogue", "scale", "attention", "other",
                     "tiles", "mma:x_wait", "mma:issue"]
        return {n: (out[2 * i], out[2 * i + 1]) for i, n in enumerate(names)}

    def disable_peer_allreduce(self):
        """back to ncclAllReduce (e.g. after LNB_ETIMEOUT); every rank must do the same"""
        check(lib.lnb_session_p2p_disable(self.h))


class Model:
    """model.Model (src/model/model.go:43-107): args + vocabulary + transformer."""

    def __init__(self, args: ModelArgs, device: int = 0, tp_rank: int = 0, tp_size: int = 1, nccl_id=None):
        self.ModelArgs = args
        self.Vocabulary = Vocabulary()
        self.Transformer = LlamaTransformer(args, device, tp_rank, tp_size, nccl_id)

    def Free(self):
        self.Transformer.close()


def LoadSyntheticModel(args_dict: dict | None = None, seed: int = synth.SEED, device: int = 0, tp_rank: int = 0,
                       tp_size: int = 1, nccl_id=None) -> Model:
    """Stand-in for model.LoadModel (src/model/loader.go:18-70) when no checkpoint exists:
    random-init weights of the given architecture, generated directly in HBM."""
    args = ModelArgs.from_c_dict(args_dict or synth.LLAMA31_8B)
    m = Model(args, device, tp_rank, tp_size, nccl_id)
    m.Transformer.init_synthetic(seed)
    m.Transformer.finalize()
    return m


def LoadModelFromTensors(args_dict: dict, tensors: dict[str, np.ndarray], device: int = 0, tp_rank: int = 0,
                         tp_size: int = 1, nccl_id=None) -> Model:
    """The reference's flow (loader.go:22-70): tensors come from host memory (the checkpoint mmap)
    and are bound by name + shape."""
    args = ModelArgs.from_c_dict(args_dict)
    m = Model(args, device, tp_rank, tp_size, nccl_id)
    for name, arr in tensors.items():
        m.Transformer.upload_tensor(name, arr)
    m.Transformer.finalize()
    return m


def load_model_args(modelDir: str, max_seq_len: int = 2048) -> dict:
    """loadModelArgs (src/model/loader.go:72-82): params.json -> model args incl. the derived widths"""
    import os
    c = _capi.ModelArgsC()
    check(lib.lnb_model_args_from_params_json(os.path.join(modelDir, "params.json").encode(), max_seq_len, C.byref(c)))
    return {f: getattr(c, f) for f, _ in _capi.ModelArgsC._fields_}


def LoadModel(modelDir: str, device: int = 0, tp_rank: int = 0, tp_size: int = 1, nccl_id=None, max_seq_len: int = 2048) -> Model:
    """model.LoadModel (src/model/loader.go:18-70): <modelDir>/consolidated.00.pth + params.json.
    tokenizer.model is optional here: when present it becomes model.Vocabulary and supplies / checks VocabSize like
    the reference (loader.go:84-120); when absent VocabSize falls back to the rows of tok_embeddings.weight."""
    import os
    from .torch_reader import TorchModelReader
    path = os.path.join(modelDir, "consolidated.00.pth")
    d = load_model_args(modelDir, max_seq_len)
    vocab = None
    vocab_path = os.path.join(modelDir, "tokenizer.model")
    if os.path.exists(vocab_path):                      # loadVocab + checkModelArgs (loader.go:84-120)
        from . import vocabulary
        vocab = vocabulary.Load(vocab_path)
        if d["vocab_size"] < 1:
            d["vocab_size"] = len(vocab)
        elif d["vocab_size"] != len(vocab):
            raise _capi.LnbError(-1, f"error while checking config and model: [VocabSize={d['vocab_size']} and vocabulary model "
                                     f"length={len(vocab)} aren't equal]")
    if d["vocab_size"] < 1:
        with TorchModelReader(path) as r:
            t = r.Load().get("tok_embeddings.weight")
        if t is None:
            raise _capi.LnbError(-1, 'tensor "tok_embeddings.weight" not found')
        d["vocab_size"] = t.Size[0]
    m = Model(ModelArgs.from_c_dict(d), device, tp_rank, tp_size, nccl_id)
    if vocab is not None:
        m.Vocabulary = vocab
    try:
        m.Transformer.load_pth(path)
        m.Transformer.finalize()
    except Exception:
        m.Free()
        raise
    return m


def applyScaling(freqs: ml.Tensor) -> ml.Tensor:
    """applyScaling (src/model/llamatransformer.go:662-692): the Llama-3.1 inverse-frequency remap, evaluated in
    float32 per element on a bf16 vector"""
    scaleFactor, lowFreqFactor, highFreqFactor, oldContextLen = np.float32(8), np.float32(1), np.float32(4), np.float32(8192)
    lowFreqWavelen, highFreqWavelen = oldContextLen / lowFreqFactor, oldContextLen / highFreqFactor
    twoPi = np.float32(2.0 * np.pi)
    out = []
    for freq in freqs.to_f32_array():
        freq = np.float32(freq)
        wavelen = twoPi / freq
        if wavelen < highFreqWavelen:
            new = freq
        elif wavelen > lowFreqWavelen:
            new = freq / scaleFactor
        else:
            smooth = (oldContextLen / wavelen - lowFreqFactor) / (highFreqFactor - lowFreqFactor)
            new = np.float32((np.float32(1) - smooth) * freq) / scaleFactor + np.float32(smooth * freq)
        out.append(np.float32(new))
    return ml.Tensor.from_f32(np.array(out, np.float32), ml.DT_BF16)


def precomputeFreqsCis(dim: int, end: int, theta: float, useScaled: bool) -> ml.Tensor:
    """precomputeFreqsCis (llamatransformer.go:694-751) from the ml builders, exactly as the Go host does it; the
    result ([end, dim/2] complex64) can be handed to lnb_model_set_rope_table, which otherwise builds the same bits"""
    freqs = ml.ARange(0, dim, 2, ml.DT_BF16)
    val = freqs.to_f32_array()
    freqs = ml.Tensor.from_f32((1.0 / np.power(theta, (val / np.float32(dim)).astype(np.float64))).astype(np.float32), ml.DT_BF16)
    t = ml.ARange(0, end, 1, ml.DT_BF16)
    if useScaled:
        freqs = applyScaling(freqs)
    freqs = ml.Outer(t, freqs)
    return ml.Polar(ml.OnesLike(freqs), freqs)
