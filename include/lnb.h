/*
 * lnb.h -- C-ABI of liblnb.so: the B200 (sm_100a) forward path that sits under the
 * Go API of adalkiran/llama-nuts-and-bolts.
 *
 * The reference has no FFI or plugin interface (it is pure Go, CPU only); the
 * drop-in boundary is its exported Go API, and this header is what a cgo shim in
 * src/ml and src/model binds (INTEGRATION.md shows the stubs).  Every entry point
 * cites the reference symbol (file:line, relative to the reference repo) whose
 * arithmetic it replaces.  Plain C types only; no pointer is retained past a call
 * (cgo pointer rule) -- weights are copied to HBM at upload.
 *
 * Error convention (reference: `(nil, error)` from every op, e.g.
 * src/ml/operations_impl.go:428-446): every function returns 0 on success or a
 * negative LNB_E* code; lnb_last_error() returns the thread-local message.
 *
 * Numeric contract: bf16 storage, f32 accumulate,
 * f32 -> bf16 by TRUNCATION after every op (src/dtype/bfloat16.go:59-61), f64
 * softmax without max-subtraction, table SiLU, RoPE through f64 intermediates.
 * The accumulation ORDER of the long sums is selectable:
 *   LNB_ACC_STRICT  k = 0,1,2,... sequentially per output, exactly the reference's
 *                   order -> results are bit-identical to the Go CPU path.
 *   LNB_ACC_FAST    each output's sum is split into interleaved partial sums that
 *                   are combined in a fixed documented order -> same
 *                   values up to fp32 summation order; faster (HBM-bound).
 */
#ifndef LNB_H
#define LNB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNB_VERSION 100

enum {
  LNB_OK = 0,
  LNB_EINVAL = -1,   /* bad argument / shape / dtype (reference: fmt.Errorf shape errors) */
  LNB_ECUDA = -2,    /* CUDA runtime failure (sticky on the handle) */
  LNB_ENCCL = -3,    /* NCCL failure */
  LNB_ESTATE = -4,   /* call order (e.g. forward before finalize) */
  LNB_ENOMEM = -5,
  LNB_ETIMEOUT = -6  /* a tensor-parallel peer did not deliver its all-reduce words in time (late / dead rank) */
};

enum { LNB_ACC_STRICT = 0, LNB_ACC_FAST = 1 };

const char* lnb_last_error(void);
int lnb_version(void);
/* number of visible CUDA devices, or a negative error */
int lnb_device_count(void);

/* ------------------------------------------------------------------------------------
 * Model-level API -- replaces model.NewLlamaTransformer / LlamaTransformer.Forward
 * ---------------------------------------------------------------------------------- */

/* mirrors model.ModelArgs (src/model/modelargs.go:10-64) plus the derived FFN width
 * (src/model/llamatransformer.go:569-577) */
typedef struct {
  int32_t dim, n_layers, n_heads, n_kv_heads, head_dim, ffn_dim, vocab_size, max_seq_len;
  float norm_eps;
  double rope_theta;
  int32_t use_scaled_rope;
} lnb_model_args;

typedef struct lnb_model lnb_model;     /* device-resident weights + tables of ONE rank */
typedef struct lnb_session lnb_session; /* model.InferenceContext: KV cache + activations */

/* Creates the device-side model of tensor-parallel rank `tp_rank` of `tp_size` on CUDA
 * device `device`.  tp_size==1: nccl_unique_id may be NULL.  tp_size>1: one process per
 * rank; nccl_unique_id is the 128-byte id from lnb_nccl_unique_id() of rank 0, sent by
 * the host (torch.distributed / Go) to every rank.
 * Replaces: model.NewLlamaTransformer (src/model/llamatransformer.go:64-113). */
int lnb_model_create(const lnb_model_args* args, int device, int tp_rank, int tp_size,
                     const void* nccl_unique_id, lnb_model** out);
int lnb_nccl_unique_id(void* out128);
/* Host-only arithmetic (no CUDA call): the [row0:row0+rows, col0:col0+cols] window of checkpoint
 * tensor `name` that rank tp_rank of tp_size keeps (Megatron split: wq/wk/wv/w1/w3/output by rows,
 * wo/w2 by columns, norms and embeddings replicated). */
int lnb_tp_shard_window(const lnb_model_args* args, const char* name, int tp_rank, int tp_size,
                        int64_t* row0, int64_t* col0, int64_t* rows, int64_t* cols);

/* Upload one checkpoint tensor by its reference name ("layers.7.attention.wq.weight",
 * "tok_embeddings.weight", ...; names and shapes: llamatransformer.go:84-105,191,202,
 * 273-282,580-586; getTensor shape check: src/model/loader.go:183-197).  `host` is the
 * FULL bf16 row-major [out,in] tensor exactly as it sits in the checkpoint mmap
 * (src/torch/types.go:51-56); the library copies (and, for tp_size>1, slices) it into
 * its HBM layout and keeps no reference to `host`. */
int lnb_model_upload_tensor(lnb_model* m, const char* name, const uint16_t* host, const int64_t* shape, int ndim);

/* ---- checkpoint files --------------------------------------------------------
 * Native reader / writer for PyTorch ".pth" archives (zip + pickle), host-only.
 * Replaces: torch.NewTorchModelReader / Load / persistentLoad (src/torch/torchmodelreader.go:21-145),
 * rebuild_tensor_v2 + TorchStorage.Load (src/torch/types.go:23-56), the unpickler (src/pickle/picklereader.go, pickledispatch.go)
 * and the read-only mmap (src/common/memorymapper_unix.go:21-45). */
typedef struct lnb_pth lnb_pth;
enum { LNB_PTH_BF16 = 0, LNB_PTH_F16 = 1, LNB_PTH_F32 = 2, LNB_PTH_F64 = 3, LNB_PTH_I8 = 4, LNB_PTH_U8 = 5,
       LNB_PTH_I16 = 6, LNB_PTH_I32 = 7, LNB_PTH_I64 = 8, LNB_PTH_BOOL = 9 };
/* maps the file, indexes the zip, unpickles the single *.pkl (torchmodelreader.go:39-66) */
int lnb_pth_open(const char* path, lnb_pth** out);
int lnb_pth_close(lnb_pth* f);
/* number of tensors in the top-level dict (pickle.PickleDict keys, torchmodelreader.go:57-63) */
int lnb_pth_tensor_count(const lnb_pth* f);
/* name / dtype / shape (<= 8 dims) / byte range inside the file of tensor `index` (dict order).
 * Returns 0, 1 when the tensor is not contiguous, or a negative error.  Any out pointer may be NULL. */
int lnb_pth_tensor_info(const lnb_pth* f, int index, const char** name, int* dtype, int* ndim, int64_t* shape,
                        int64_t* file_offset, int64_t* nbytes);
/* pointer into the read-only mapping (the reference's Tensor.RawData aliasing the mmap, types.go:51-56);
 * valid until lnb_pth_close */
const void* lnb_pth_tensor_data(const lnb_pth* f, int index);
/* model.LoadModelEx's tensor half (src/model/loader.go:22-41) fused with the by-name binding of
 * NewLlamaTransformer (llamatransformer.go:84-105): every tensor the architecture names goes from the page
 * cache to HBM (tp_size>1: only this rank's window); other entries are ignored.  lnb_model_finalize still
 * reports missing tensors. */
int lnb_model_load_pth(lnb_model* m, const char* path, int* n_uploaded);
/* loadModelArgsFromFile (src/model/modelargs.go:52-64) with the defaults of NewModelArgs (:29-44) and the
 * derived HeadDim / N_KVHeads / FFN width (llamatransformer.go:73-82,569-577).  vocab_size stays -1 when
 * params.json has none (the reference takes it from the tokenizer, loader.go:106-111). */
int lnb_model_args_from_params_json(const char* path, int max_seq_len, lnb_model_args* out);
/* writer: one storage per tensor at offset 0, pickle protocol 2 restricted to the opcodes the reference's
 * unpickler dispatches (src/pickle/pickledispatch.go:52-77) -- readable by torch.load and by the Go loader */
typedef struct lnb_pth_writer lnb_pth_writer;
int lnb_pth_writer_create(const char* path, lnb_pth_writer** out);
int lnb_pth_writer_add(lnb_pth_writer* w, const char* name, int dtype, const void* data, const int64_t* shape, int ndim);
/* writes data.pkl + the zip directory and frees the writer (also on failure) */
int lnb_pth_writer_finish(lnb_pth_writer* w);
/* the synthetic checkpoint of lnb_model_init_synthetic as a consolidated.00.pth (same bits, host-only) */
int lnb_pth_write_synthetic(const char* path, const lnb_model_args* args, uint64_t seed);

/* ---- tokenizer ------------------------------------------------------------------
 * tiktoken vocabulary + Llama-3 BPE, host-only.  Replaces tiktoken.Load (src/tiktoken/tiktokenreader.go:12-85),
 * model.NewVocabulary incl. its split regexp (src/model/vocabulary.go:23-50), InferenceEngine.TokenizeString /
 * bytePairMerge / Tokenize (src/inference/tokenize.go:27-193) and the byte concatenation under
 * TokenBatchToString (:239-258).  The console's emoji alias annotation (src/inference/emoji.go) is not built. */
typedef struct lnb_vocab lnb_vocab;
/* tokenizer.model: "<base64 token> <rank>" lines; the 256 special tokens are appended after the ranks */
int lnb_vocab_load(const char* tokenizer_model_path, lnb_vocab** out);
int lnb_vocab_destroy(lnb_vocab* v);
/* writes a stand-in tokenizer.model with n_mergeable ranks (256 bytes + unique 3-byte fillers that never merge) so that a
 * synthetic model directory passes checkModelArgs' VocabSize == vocabulary length (src/model/loader.go:98-120) */
int lnb_vocab_write_synthetic(const char* path, int n_mergeable);

/* len(Vocabulary.IdToToken) (vocabulary.go:27) */
int lnb_vocab_size(const lnb_vocab* v);
/* Vocabulary.TokenToId[token]; *id = -1 when absent.  `token` is raw bytes (pieces need not be valid UTF-8) */
int lnb_vocab_token_id(const lnb_vocab* v, const void* token, int token_len, int32_t* id);
/* Vocabulary.IdToToken[id]: pointer into the vocabulary, valid until lnb_vocab_destroy */
int lnb_vocab_token_bytes(const lnb_vocab* v, int32_t id, const void** bytes, int* len);
/* BeginOfSentenceId, EndOfSentenceId, PadId (-1) and the two StopTokenIds <|eom_id|>, <|eot_id|> (tiktokenreader.go:74-82) */
int lnb_vocab_special_ids(const lnb_vocab* v, int32_t* bos, int32_t* eos, int32_t* pad, int32_t* stop2);
/* Vocabulary.SplitRegexp.FindAllString (src/model/vocabulary.go:36, used at tokenize.go:180): byte offsets one past
 * each piece of `text`, by a hand-written matcher with Go/RE2 semantics (leftmost-first, ASCII \s, \p{L} / \p{N} of
 * Unicode 15.0, simple case folding in the contraction alternative) */
int lnb_split_pieces(const char* text, int64_t text_len, int64_t* ends, int cap, int* n_out);
/* InferenceEngine.TokenizeString (tokenize.go:175-193).  *n_out = tokens produced (also when cap is too small) */
int lnb_tokenize_string(const lnb_vocab* v, const char* text, int64_t text_len, int32_t* out, int cap, int* n_out);
/* InferenceEngine.Tokenize (tokenize.go:27-95): <|begin_of_text|>, every non-empty part wrapped in header tokens and
 * closed with <|eot_id|>, then the open assistant header */
int lnb_tokenize_prompt(const lnb_vocab* v, const char* const* headers, const char* const* contents, int n_parts,
                        int32_t* out, int cap, int* n_out);
/* bytes of the tokens up to the first PadId (TokenBatchToString :239-258 without the emoji annotation) */
int lnb_detokenize(const lnb_vocab* v, const int32_t* ids, int n, char* out, int64_t cap, int64_t* n_out);

/* Random-init every tensor directly in HBM with the synthetic generator of DESIGN.md
 * (no checkpoint exists in the build environment).  Same bits as oracle's
 * orc_synth_fill for the same seed. */
int lnb_model_init_synthetic(lnb_model* m, uint64_t seed);
/* host-side twin of the generator, for callers that want the same bytes in host memory */
int lnb_synth_fill_host(uint64_t seed, const char* name, float scale, float offset, int64_t n, uint16_t* out);
/* scale/offset used by lnb_model_init_synthetic for a tensor name */
int lnb_synth_spec(const lnb_model_args* args, const char* name, float* scale, float* offset);

/* Optional: override the tables the library builds itself at finalize.
 * cis: [rows][head_dim/2][2] f32 == model.precomputeFreqsCis output
 * (llamatransformer.go:694-751); silu_bf16: ml.TABLE_SILU truncated to bf16
 * (src/ml/activations.go:11-25,38). */
int lnb_model_set_rope_table(lnb_model* m, const float* cis, int rows);
int lnb_model_set_silu_table(lnb_model* m, const uint16_t* silu_bf16_65536);
/* read back the tables in use (tests compare them with the oracle's) */
int lnb_model_get_rope_table(lnb_model* m, float* cis_out, int rows);
int lnb_model_get_silu_table(lnb_model* m, uint16_t* out65536);

/* checks all 291 tensors are present, builds tables, makes the model immutable */
int lnb_model_finalize(lnb_model* m);
int lnb_model_destroy(lnb_model* m); /* hook for model.Model.Free (src/model/model.go:56) */

/* model.NewInferenceContext (src/model/inferencecontext.go:17-46): KV cache of seq_len rows
 * per layer, zero-filled.  max_rows = largest S a forward call may carry (>=1).
 * acc_mode = LNB_ACC_STRICT / LNB_ACC_FAST. */
int lnb_session_create(lnb_model* m, int seq_len, int max_rows, int acc_mode, lnb_session** out);
int lnb_session_destroy(lnb_session* s);

/* LlamaTransformer.Forward (src/model/llamatransformer.go:145-180) + the last-row
 * ml.Argmax of the generate loop (src/inference/inference.go:207-216).
 *   tokens[S] int32, start_pos as in the reference (S>1 requires start_pos==0, F11).
 *   logits : NULL, or host [S, vocab] f32 if all_rows, host [1, vocab] (last row) otherwise
 *   argmax_last : NULL, or receives the greedy token of the last row.
 * One call = one H2D of tokens, the whole forward on the device, one D2H of what was asked. */
int lnb_forward(lnb_session* s, const int32_t* tokens, int S, int start_pos,
                float* logits, int all_rows, int32_t* argmax_last);

/* Batched decode (BASELINE.json configs[4], "concurrent prompts"): a session with n_seq (<= 8) independent
 * sequences = n_seq reference InferenceContexts sharing one set of weights (the reference has no batch
 * dimension).  lnb_session_set_active_sequence selects which sequence lnb_forward /
 * lnb_decode_run / lnb_session_read address (e.g. to prefill each prompt); lnb_forward_batch then advances
 * all sequences by one token in ONE pass over the weights: row i = sequence i, token tokens[i] at position
 * positions[i].  Every row gets exactly the arithmetic of its own S=1 Forward. */
int lnb_session_create_batch(lnb_model* m, int seq_len, int n_seq, int max_rows, int acc_mode, lnb_session** out);
int lnb_session_set_active_sequence(lnb_session* s, int seq);
int lnb_forward_batch(lnb_session* s, const int32_t* tokens, const int32_t* positions, int n,
                      float* logits /* NULL or [n, vocab] */, int32_t* argmax_out /* [n] */);

/* LlamaTransformer.Forward whose result stays in HBM: the host mirror returns an ml.Tensor that carries a handle
 * (session, generation) instead of S x vocab floats, Tensor.Slice keeps the handle, and ml.Argmax on it runs on the
 * device -- the reference's call sequence Forward -> Slice(last row) -> Argmax (src/inference/inference.go:202-216)
 * moves 4 bytes per token over PCIe.  Touching the tensor's data calls lnb_session_logits_read.
 * rows_kept: 1 (last row) or S.  A handle is valid until the session's next forward (LNB_ESTATE afterwards).
 * Tensor-parallel: logits_read / logits_argmax are collectives (every rank makes the same call). */
int lnb_forward_device(lnb_session* s, const int32_t* tokens, int S, int start_pos, int rows_kept,
                       int32_t* argmax_last /* may be NULL */, int64_t* generation_out);
int lnb_session_logits_read(lnb_session* s, int64_t generation, int row0, int rows, float* host /* [rows, vocab] */);
int lnb_session_logits_argmax(lnb_session* s, int64_t generation, int row0, int rows, int32_t* out /* [rows] */);

/* EXTENSION beyond the reference (chunked prefill): accept Forward calls of S > 1 tokens at startPos > 0 and
 * mask them with the [S,T] causal mask (row s sees keys t <= startPos + s).  The reference's [S,S] mask
 * (src/model/llamatransformer.go:128-136, added to [32,S,T] scores at :469-473) only broadcasts for startPos 0, which is
 * why the library refuses such calls by default, like the Go code would fail.  Chunked and one-shot prefill give
 * bit-identical caches and logits in LNB_ACC_STRICT. */
int lnb_session_set_chunked_prefill(lnb_session* s, int on);

/* Device-resident greedy decode: runs n_steps consecutive S=1 forwards starting with
 * `first_token` at position start_pos, feeding each argmax back on the device (no host
 * sync inside), optionally as CUDA-graph replays.  tokens_out[n_steps]; ms_out = device
 * time of the n_steps steps measured with CUDA events on the session stream.
 * This is the timed region of bench.py's `value`. */
int lnb_decode_run(lnb_session* s, int32_t first_token, int start_pos, int n_steps, int use_graph,
                   int32_t* tokens_out, float* ms_out);

/* Fused collective for tensor-parallel decode: instead of ncclAllReduce after Wo / w2, every rank's GEMV
 * epilogue stores its fp32 partials straight into all peers' memory over NVLink and a small kernel
 * reduces the N slots in rank order (bit-identical on every rank).  Setup: every rank exports the 64-byte
 * CUDA IPC handle of its session region, the host gathers the tp_size handles, every rank imports them
 * (index = rank).  Sessions must be created and driven identically on all ranks.  Without this call the
 * session uses NCCL (exactly one all-reduce after Wo and one after w2, as north_star prescribes). */
int lnb_session_p2p_export(lnb_session* s, void* handle64);
int lnb_session_p2p_import(lnb_session* s, const void* handles /* tp_size x 64 bytes */, int n);
/* Every in-kernel wait for a peer is bounded (LNB_P2P_TIMEOUT_MS, default 1500): a late or dead rank turns the
 * running call into LNB_ETIMEOUT on its peers instead of hanging their GPUs.  After that error the session's peer
 * path is poisoned; lnb_session_p2p_disable switches the session back to ncclAllReduce (all ranks must do the same). */
int lnb_session_p2p_disable(lnb_session* s);

/* debugging / parity probes */
enum { LNB_BUF_RESIDUAL = 0, LNB_BUF_CACHE_K = 1, LNB_BUF_CACHE_V = 2, LNB_BUF_LOGITS = 3 };
/* copies a device buffer of the session to host: RESIDUAL [S,dim] bf16 (after the last
 * forward; with lnb_session_set_layer_limit, after that many layers), CACHE_K/V of `layer`
 * [seq_len, n_kv_local, head_dim] bf16, LOGITS [rows, vocab_local] f32. */
int lnb_session_read(lnb_session* s, int which, int layer, void* host, int64_t nbytes);
int lnb_session_set_layer_limit(lnb_session* s, int n_layers_to_run); /* <=0: all */
/* Which decode path does the session use?  1 = the persistent decode engine (csrc/engine.cuh: one kernel launch runs n
 * complete S=1 forwards), 0 = the kernel chain (one launch per projection / attention / norm scale; lnb_last_error()
 * tells why).  LNB_ENGINE=1 / 0 forces the choice; the default is the engine except for single-GPU LNB_ACC_FAST. */
int lnb_session_decode_engine(lnb_session* s);
/* profiling aid of the persistent decode engine (LNB_ENGINE_PROF=1 when the session first decodes): cycles of consumer
 * thread 0 per section, {mean, max} over the CTAs, reset on read.  out[24]. */
int lnb_session_engine_profile(lnb_session* s, double* out16);
/* number of kernels launched by this session since creation (bench.py gpu_launches) */
int64_t lnb_session_launch_count(lnb_session* s);
int lnb_session_sync(lnb_session* s);
/* Measurement hook for bench.py's roofline: times ONE projection kernel of the decode step in
 * isolation (all layers' weights back to back so every launch streams from HBM; `reps` sweeps;
 * CUDA events on the session stream).  kind: 0 wq|wk|wv, 1 wo, 2 w1|w3, 3 w2, 4 LM head. */
int lnb_session_bench_kernel(lnb_session* s, int kind, int reps, float* ms_per_launch, int64_t* bytes_per_launch,
                             int* launches);

/* ------------------------------------------------------------------------------------
 * Op-level API -- what the src/ml shims bind.  Host pointers in and out; each call
 * stages through HBM and runs the same kernels the model path uses.
 * ---------------------------------------------------------------------------------- */

/* ml.LinearTransformation, BF16 (src/ml/operations_impl.go:427-447,
 * operations_lineartransform.go:37-70,145-207): x[S,K], w[N,K] -> out[S,N] */
int lnb_op_linear_bf16(const uint16_t* x, const uint16_t* w, uint16_t* out, int S, int K, int N, int acc_mode);
/* ml.MatMul, BF16 (operations_impl.go:449-476, operations_matmul.go:24-60,136-182):
 * a[B,M,K] x b[B,K,N] -> out[B,M,N] */
int lnb_op_matmul_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int B, int M, int K, int N);
/* RMSNorm.Forward (src/model/llamatransformer.go:633-660) */
int lnb_op_rmsnorm_bf16(const uint16_t* x, const uint16_t* w, uint16_t* out, int S, int D, float eps, int acc_mode);
/* the scale of RMSNorm.Forward alone (llamatransformer.go:641-656: Pow, Mean, AddScalar, RSqrt):
 * r[s] = f32(1 / sqrt(f64( (sequential fp32 sum_k x[s,k]^2) / D + eps ))), bit-exact.
 * algo: 0 = what the model path uses (env LNB_RMS_ALGO=chain|scan|seg overrides the built-in choice),
 * 1 = one-thread FADD chain, 2 = iterative binade scan, 3 = one-pass predict / fold / walk
 * (csrc/seqsum.cuh; 2 and 3 return LNB_EINVAL when D does not fit their shape), 4 = the variant that runs inside the
 * persistent decode engine (256 threads, runs folded per warp; csrc/engine.cuh).  Identical bits. */
int lnb_op_rms_scale_f32(const uint16_t* x, float* r, int S, int D, float eps, int algo);
/* applyRotaryEmbeddings for one tensor (llamatransformer.go:753-790): x[S,H,hd] */
int lnb_op_rope_bf16(const uint16_t* x, const float* cis, uint16_t* out, int S, int H, int hd, int start_pos);
/* the attention core of LlamaAttention.Forward (llamatransformer.go:402-514):
 * q[S,n_heads,hd] rotated, cache_k/v[T,n_kv,hd] -> out[S,n_heads*hd] */
int lnb_op_attention_bf16(const uint16_t* q, const uint16_t* cache_k, const uint16_t* cache_v, uint16_t* out,
                          int S, int T, int n_heads, int n_kv, int hd, int causal_mask, int acc_mode);
/* ml.Silu (activations.go:27-50), ml.Add / ml.MultiplyElementwise same-shape bf16
 * (operations_impl.go:307-335,367-395) */
int lnb_op_silu_bf16(const uint16_t* x, uint16_t* out, int64_t n);
int lnb_op_add_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n);
int lnb_op_mul_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n);
/* ml.Softmax on f32 rows (operations_impl.go:478-511), ml.Argmax (:513-548) */
int lnb_op_softmax_f32(const float* x, float* out, int rows, int cols);
int lnb_op_argmax_f32(const float* x, int rows, int cols, int32_t* out);
/* ml.Fwd_Get_Rows (operations_impl.go:142-173) */
int lnb_op_get_rows_bf16(const uint16_t* emb, const int32_t* tokens, uint16_t* out, int S, int vocab, int dim);

#ifdef __cplusplus
}
#endif
#endif /* LNB_H */
