// lnb_host.hpp -- C++ host mirror of the reference's Go packages above the C-ABI (include/lnb.h).
//
// The reference is compiled Go; the build image has no Go toolchain, so this header restates, in C++17,
// the host-side types and call sequence a Go build would keep: ml::Tensor + the ml ops on the forward
// path, model::ModelArgs / LlamaTransformer / InferenceContext, and inference::InferenceEngine with the
// generate loop of src/inference/inference.go:173-254.  Same names, argument meaning and error texts as
// the Go code (errors are exceptions instead of `(nil, error)`).  Header-only; links against liblnb.so.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../include/lnb.h"

namespace lnb_host {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(int rc) {
  if (rc < 0) throw Error(std::string("lnb: ") + lnb_last_error());
}

namespace ml {  // ---- src/ml ---------------------------------------------------------------------------
enum class DataType { BF16, F32, INT32 };  // datatype.go:11-34
inline size_t item_size(DataType t) { return t == DataType::BF16 ? 2 : 4; }

struct Tensor {  // tensor.go:11-19 (contiguous row-major, Size + RawData)
  std::vector<int> Size;
  DataType DT = DataType::BF16;
  std::vector<uint8_t> RawData;
  std::string Name;
  Tensor() = default;
  Tensor(std::vector<int> size, DataType dt) : Size(std::move(size)), DT(dt) { RawData.assign(GetElementCount() * item_size(dt), 0); }
  size_t GetElementCount() const {
    size_t n = 1;
    for (int d : Size) n *= (size_t)d;
    return n;
  }
  template <class T> T* data() { return reinterpret_cast<T*>(RawData.data()); }
  template <class T> const T* data() const { return reinterpret_cast<const T*>(RawData.data()); }
};

inline uint16_t Float32ToBFloat16bits(float f) {  // src/dtype/bfloat16.go:59-61 (truncation)
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);
}
inline float BFloat16bitsToFloat32(uint16_t b) {  // :55-57
  uint32_t u = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

inline int AccMode = LNB_ACC_STRICT;

inline Tensor LinearTransformation(const Tensor& input, const Tensor& weights) {  // operations_impl.go:427-447
  if (input.DT != weights.DT) throw Error("tensors are not in same data type");
  if (input.Size.size() != 2 || weights.Size.size() != 2) throw Error("LinearTransformation needs two matrices");
  if (input.Size[1] != weights.Size[1])
    throw Error("columns size " + std::to_string(input.Size[1]) + " of input tensor should be equal with " +
                std::to_string(weights.Size[1]) + " input features count of weights tensor");
  if (input.DT != DataType::BF16) throw Error("unsupported tensor datatype");
  Tensor dst({input.Size[0], weights.Size[0]}, DataType::BF16);
  check(lnb_op_linear_bf16(input.data<uint16_t>(), weights.data<uint16_t>(), dst.data<uint16_t>(), input.Size[0], input.Size[1],
                           weights.Size[0], AccMode));
  return dst;
}

inline Tensor Argmax(const Tensor& input) {  // operations_impl.go:513-548 (last dimension)
  if (input.DT != DataType::F32) throw Error("unsupported tensor datatype");
  const int cols = input.Size.back();
  const int rows = (int)(input.GetElementCount() / cols);
  Tensor dst(std::vector<int>(input.Size.begin(), input.Size.end() - 1), DataType::INT32);
  if (dst.Size.empty()) dst = Tensor({1}, DataType::INT32);
  check(lnb_op_argmax_f32(input.data<float>(), rows, cols, dst.data<int32_t>()));
  return dst;
}
}  // namespace ml

namespace model {  // ---- src/model ------------------------------------------------------------------------
struct ModelArgs : lnb_model_args {  // modelargs.go:10-44 (+ derived HeadDim / FFN width, llamatransformer.go:569-577)
  static ModelArgs Llama31_8B() {
    ModelArgs a{};
    a.dim = 4096; a.n_layers = 32; a.n_heads = 32; a.n_kv_heads = 8; a.head_dim = 128; a.ffn_dim = 14336;
    a.vocab_size = 128256; a.max_seq_len = 2048; a.norm_eps = 1e-5f; a.rope_theta = 500000.0; a.use_scaled_rope = 1;
    return a;
  }
};
struct Vocabulary {  // vocabulary.go (ids only)
  int32_t PadId = -1;
  std::vector<int32_t> StopTokenIds{128008, 128009};  // src/tiktoken/tiktokenreader.go:81
};

// src/tiktoken + model.Vocabulary's tables + the tokenizer half of src/inference/tokenize.go over lnb_vocab
struct PromptPart {  // inference.PromptPart (tokenize.go:21-25)
  std::string Header, Content;
};
class Tokenizer {
 public:
  explicit Tokenizer(const std::string& vocabFilePath) { check(lnb_vocab_load(vocabFilePath.c_str(), &h_)); }   // tiktoken.Load
  ~Tokenizer() { lnb_vocab_destroy(h_); }
  Tokenizer(const Tokenizer&) = delete;
  int Size() const { return lnb_vocab_size(h_); }
  Vocabulary GetVocabulary() const {   // PadId / StopTokenIds of NewVocabulary (vocabulary.go:23-50)
    Vocabulary v;
    int32_t stop[2];
    check(lnb_vocab_special_ids(h_, nullptr, nullptr, &v.PadId, stop));
    v.StopTokenIds = {stop[0], stop[1]};
    return v;
  }
  std::vector<int32_t> TokenizeString(const std::string& text) const {   // tokenize.go:175-193
    std::vector<int32_t> out(text.size() + 8);
    int n = 0;
    check(lnb_tokenize_string(h_, text.data(), (int64_t)text.size(), out.data(), (int)out.size(), &n));
    out.resize((size_t)n);
    return out;
  }
  std::vector<int32_t> Tokenize(const std::vector<PromptPart>& parts) const {   // tokenize.go:27-95
    std::vector<const char*> hs, cs;
    size_t cap = 32;
    for (const auto& p : parts) {
      hs.push_back(p.Header.c_str());
      cs.push_back(p.Content.c_str());
      cap += p.Header.size() + p.Content.size() + 16;
    }
    std::vector<int32_t> out(cap);
    int n = 0;
    check(lnb_tokenize_prompt(h_, hs.data(), cs.data(), (int)parts.size(), out.data(), (int)out.size(), &n));
    out.resize((size_t)n);
    return out;
  }
  // Vocabulary.IdToToken[id] (raw bytes; a piece need not be valid UTF-8)
  std::string IdToToken(int32_t id) const {
    const void* p = nullptr;
    int n = 0;
    check(lnb_vocab_token_bytes(h_, id, &p, &n));
    return std::string((const char*)p, (size_t)n);
  }
  // TokenToString (tokenize.go:195-237) without the emoji annotation: a piece that is not valid UTF-8 on its own is
  // collected in waitingBytes and released one rune per call once the bytes decode; returns addedToWaiting
  bool TokenToString(int32_t id, std::string& waitingBytes, std::string& text) const {
    const std::string piece = IdToToken(id);
    text.clear();
    if (Utf8Valid(piece)) { text = piece; return false; }
    waitingBytes += piece;
    if (!Utf8Valid(waitingBytes)) return true;
    const size_t n = Utf8RuneLen((unsigned char)waitingBytes[0]);
    text = waitingBytes.substr(0, n);
    waitingBytes.erase(0, n);
    return false;
  }
  static size_t Utf8RuneLen(unsigned char c) { return c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : 4; }
  static bool Utf8Valid(const std::string& s) {   // utf8.Valid
    for (size_t i = 0; i < s.size();) {
      const unsigned char c = (unsigned char)s[i];
      size_t n;
      uint32_t cp, min;
      if (c < 0x80) { i++; continue; }
      if ((c & 0xe0) == 0xc0) { n = 2; cp = c & 0x1f; min = 0x80; }
      else if ((c & 0xf0) == 0xe0) { n = 3; cp = c & 0x0f; min = 0x800; }
      else if ((c & 0xf8) == 0xf0) { n = 4; cp = c & 0x07; min = 0x10000; }
      else return false;
      if (i + n > s.size()) return false;
      for (size_t k = 1; k < n; k++) {
        if (((unsigned char)s[i + k] & 0xc0) != 0x80) return false;
        cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3f);
      }
      if (cp < min || cp > 0x10ffff || (cp >= 0xd800 && cp <= 0xdfff)) return false;
      i += n;
    }
    return true;
  }
  std::string TokenBatchToString(const std::vector<int32_t>& ids) const {   // tokenize.go:239-258 (bytes; no emoji annotation)
    std::string out(64 + 128 * ids.size(), '\0');
    int64_t n = 0;
    int rc = lnb_detokenize(h_, ids.data(), (int)ids.size(), &out[0], (int64_t)out.size(), &n);
    if (rc != 0 && n > (int64_t)out.size()) {
      out.assign((size_t)n, '\0');
      rc = lnb_detokenize(h_, ids.data(), (int)ids.size(), &out[0], (int64_t)out.size(), &n);
    }
    check(rc);
    out.resize((size_t)n);
    return out;
  }

 private:
  lnb_vocab* h_ = nullptr;
};

class LlamaTransformer {  // llamatransformer.go:16-113
 public:
  ModelArgs args;
  LlamaTransformer(const ModelArgs& a, int device = 0) : args(a) { check(lnb_model_create(&args, device, 0, 1, nullptr, &h_)); }
  ~LlamaTransformer() { lnb_model_destroy(h_); }
  LlamaTransformer(const LlamaTransformer&) = delete;
  // getTensor + upload (loader.go:183-197)
  void UploadTensor(const std::string& name, const uint16_t* data, const std::vector<int64_t>& shape) {
    check(lnb_model_upload_tensor(h_, name.c_str(), data, shape.data(), (int)shape.size()));
  }
  void InitSynthetic(uint64_t seed) { check(lnb_model_init_synthetic(h_, seed)); }
  // torch.TorchModelReader.Load + the by-name binding of NewLlamaTransformer: checkpoint mapping -> HBM
  int LoadPth(const std::string& path) {
    int n = 0;
    check(lnb_model_load_pth(h_, path.c_str(), &n));
    return n;
  }
  void Finalize() { check(lnb_model_finalize(h_)); }
  lnb_model* handle() { return h_; }

 private:
  lnb_model* h_ = nullptr;
};

// src/torch: TorchModelReader over the native reader of liblnb.so (RawData aliases the read-only file mapping)
class TorchModelReader {  // torchmodelreader.go:13-66
 public:
  struct Entry {
    std::string Name;
    int DType = 0;                 // LNB_PTH_*
    std::vector<int64_t> Size;
    const void* RawData = nullptr;
    int64_t ByteCount = 0;
    bool Contiguous = true;
  };
  explicit TorchModelReader(const std::string& modelFilePath) { check(lnb_pth_open(modelFilePath.c_str(), &h_)); }
  ~TorchModelReader() { lnb_pth_close(h_); }
  TorchModelReader(const TorchModelReader&) = delete;
  std::vector<Entry> Load() const {
    std::vector<Entry> out;
    const int n = lnb_pth_tensor_count(h_);
    for (int i = 0; i < n; i++) {
      Entry e;
      const char* name = nullptr;
      int nd = 0;
      int64_t shape[8], off = 0;
      const int rc = lnb_pth_tensor_info(h_, i, &name, &e.DType, &nd, shape, &off, &e.ByteCount);
      check(rc < 0 ? rc : 0);
      e.Contiguous = (rc == 0);
      e.Name = name;
      e.Size.assign(shape, shape + nd);
      e.RawData = lnb_pth_tensor_data(h_, i);
      out.push_back(std::move(e));
    }
    return out;
  }

 private:
  lnb_pth* h_ = nullptr;
};

// loadModelArgsFromFile (modelargs.go:52-64) + derived widths; vocab_size < 1 -> rows of tok_embeddings.weight
// (the reference takes it from tokenizer.model, loader.go:103-111; the tokenizer is out of scope here)
inline ModelArgs LoadModelArgs(const std::string& modelDir, int maxSeqLen = 2048) {
  ModelArgs a{};
  check(lnb_model_args_from_params_json((modelDir + "/params.json").c_str(), maxSeqLen, &a));
  if (a.vocab_size < 1) {
    TorchModelReader r(modelDir + "/consolidated.00.pth");
    for (const auto& e : r.Load())
      if (e.Name == "tok_embeddings.weight" && !e.Size.empty()) a.vocab_size = (int32_t)e.Size[0];
    if (a.vocab_size < 1) throw Error("tensor \"tok_embeddings.weight\" not found");
  }
  return a;
}

// model.LoadModel (loader.go:18-70): <modelDir>/params.json + consolidated.00.pth -> a finalized transformer
inline std::unique_ptr<LlamaTransformer> LoadModel(const std::string& modelDir, int device = 0, int maxSeqLen = 2048) {
  auto t = std::make_unique<LlamaTransformer>(LoadModelArgs(modelDir, maxSeqLen), device);
  t->LoadPth(modelDir + "/consolidated.00.pth");
  t->Finalize();
  return t;
}

class InferenceContext {  // inferencecontext.go:8-46 (KV cache in HBM)
 public:
  int SequenceLength;
  InferenceContext(LlamaTransformer& t, int sequenceLength, int accMode = LNB_ACC_STRICT, int maxRows = 8)
      : SequenceLength(sequenceLength > 0 ? sequenceLength : t.args.max_seq_len), vocab_(t.args.vocab_size) {
    check(lnb_session_create(t.handle(), SequenceLength, maxRows, accMode, &h_));
  }
  // n reference InferenceContexts that share the weights (batched decode): own KV cache and position each
  InferenceContext(LlamaTransformer& t, int sequenceLength, int accMode, int maxRows, int nSeq)
      : SequenceLength(sequenceLength > 0 ? sequenceLength : t.args.max_seq_len), vocab_(t.args.vocab_size) {
    check(lnb_session_create_batch(t.handle(), SequenceLength, nSeq, maxRows, accMode, &h_));
  }
  ~InferenceContext() { lnb_session_destroy(h_); }
  void SetActiveSequence(int seq) { check(lnb_session_set_active_sequence(h_, seq)); }
  // Forward + last-row ml.Argmax in one call (inference.go:202-216), 4-byte read-back
  int32_t ForwardArgmax(const std::vector<int32_t>& tokens, int startPos) {
    int32_t next = -1;
    check(lnb_forward(h_, tokens.data(), (int)tokens.size(), startPos, nullptr, 0, &next));
    return next;
  }
  // one decode step of every sequence: tokens[i] at positions[i] -> greedy next token of sequence i
  std::vector<int32_t> ForwardBatch(const std::vector<int32_t>& tokens, const std::vector<int32_t>& positions) {
    std::vector<int32_t> next(tokens.size(), -1);
    check(lnb_forward_batch(h_, tokens.data(), positions.data(), (int)tokens.size(), nullptr, next.data()));
    return next;
  }
  InferenceContext(const InferenceContext&) = delete;
  // LlamaTransformer.Forward (llamatransformer.go:145-180): tokens [S] -> f32 logits [S, vocab]
  ml::Tensor Forward(const ml::Tensor& inputTokens, int startPos) {
    if (inputTokens.DT != ml::DataType::INT32 || inputTokens.Size.size() != 1) throw Error("inputTokens must be a 1-D Int32 tensor");
    const int S = inputTokens.Size[0];
    if (S == 0) throw Error("empty token array");
    ml::Tensor logits({S, vocab_}, ml::DataType::F32);
    check(lnb_forward(h_, inputTokens.data<int32_t>(), S, startPos, logits.data<float>(), 1, nullptr));
    return logits;
  }
  // The same Forward with the logits left in HBM (lnb_forward_device): the returned handle stands for the [rowsKept, vocab]
  // tensor; ArgmaxRows = Slice + ml.Argmax on the device (4 bytes per row back), ReadLogits = touching the tensor's data.
  // A handle is valid until the context's next forward.
  struct DeviceLogits { int64_t generation = 0; int rows = 0; };
  DeviceLogits ForwardDevice(const std::vector<int32_t>& tokens, int startPos, bool allRows = false, int32_t* argmaxLast = nullptr) {
    DeviceLogits d;
    d.rows = allRows ? (int)tokens.size() : 1;
    check(lnb_forward_device(h_, tokens.data(), (int)tokens.size(), startPos, d.rows, argmaxLast, &d.generation));
    return d;
  }
  std::vector<int32_t> ArgmaxRows(const DeviceLogits& d, int row0, int rows) {
    std::vector<int32_t> out((size_t)rows, -1);
    check(lnb_session_logits_argmax(h_, d.generation, row0, rows, out.data()));
    return out;
  }
  ml::Tensor ReadLogits(const DeviceLogits& d, int row0, int rows) {
    ml::Tensor t({rows, vocab_}, ml::DataType::F32);
    check(lnb_session_logits_read(h_, d.generation, row0, rows, t.data<float>()));
    return t;
  }
  // chunked prefill (EXTENSION): allow S > 1 calls at startPos > 0 with the [S,T] causal mask
  void AllowChunkedPrefill(bool on) { check(lnb_session_set_chunked_prefill(h_, on ? 1 : 0)); }
  // 1 = the persistent decode engine serves this context's S=1 calls, 0 = the kernel chain
  bool UsesDecodeEngine() { return lnb_session_decode_engine(h_) == 1; }
  // after LNB_ETIMEOUT on the peer all-reduce: back to ncclAllReduce (every rank must do the same)
  void DisablePeerAllReduce() { check(lnb_session_p2p_disable(h_)); }
  lnb_session* handle() { return h_; }

 private:
  lnb_session* h_ = nullptr;
  int vocab_;
};
}  // namespace model

namespace inference {  // ---- src/inference ------------------------------------------------------------------
enum GenerationState { GSInProgress = 0, GSFinishedByReachingEOS = 1, GSFinishedByReachingSeqLen = 2 };

// generateTokensInternal (inference.go:173-254); `emit` plays generatedTokensCh
inline void GenerateTokens(model::LlamaTransformer& transformer, const model::Vocabulary& vocab, int sequenceLength, int accMode,
                           const std::vector<int32_t>& promptTokens, const std::function<void(GenerationState, int32_t)>& emit) {
  // the prefill call carries the whole prompt: size the session's row buffers for it
  model::InferenceContext infContext(transformer, sequenceLength, accMode, std::max<int>(8, (int)promptTokens.size()));
  const int promptLength = (int)promptTokens.size();
  if (promptLength >= infContext.SequenceLength)
    throw Error("context SequenceLength " + std::to_string(infContext.SequenceLength) + " must be higher than prompt tokens length " +
                std::to_string(promptLength));
  std::vector<int32_t> tokens(infContext.SequenceLength, vocab.PadId);  // ml.Full(..., PadId) :181
  std::copy(promptTokens.begin(), promptTokens.end(), tokens.begin());
  int prevPos = 0;
  for (int curPos = promptLength; curPos < infContext.SequenceLength; curPos++) {  // :194
    ml::Tensor slice({curPos - prevPos}, ml::DataType::INT32);                     // tokens.Slice([prevPos],[curPos]) :195
    std::memcpy(slice.RawData.data(), tokens.data() + prevPos, (size_t)(curPos - prevPos) * 4);
    ml::Tensor logits = infContext.Forward(slice, prevPos);                        // :202
    const int rows = logits.Size[0], V = logits.Size[1];
    ml::Tensor last({1, V}, ml::DataType::F32);                                    // logits.Slice(last row) :207
    std::memcpy(last.RawData.data(), logits.data<float>() + (size_t)(rows - 1) * V, (size_t)V * 4);
    int32_t nextTokenId = ml::Argmax(last).data<int32_t>()[0];                     // :211
    if (tokens[curPos] != vocab.PadId) nextTokenId = tokens[curPos];               // :218-226
    tokens[curPos] = nextTokenId;
    bool eos = false;
    for (int32_t s : vocab.StopTokenIds) eos |= (s == nextTokenId);
    prevPos = curPos;
    if (eos) { emit(GSFinishedByReachingEOS, nextTokenId); break; }                 // :233-240
    if (curPos + 1 == infContext.SequenceLength) { emit(GSFinishedByReachingSeqLen, nextTokenId); break; }
    emit(GSInProgress, nextTokenId);
  }
}

// The consumer InferenceEngine.TokenizeBatch (src/inference/tokenize.go:97-107) never got in the reference
// (SURVEY 8f-4): every prompt runs the loop above -- own KV cache, positions and stop condition -- but all
// sequences advance together, one pass over the weights per step.  emit(sequence, state, token).
inline void GenerateTokensBatch(model::LlamaTransformer& transformer, const model::Vocabulary& vocab, int sequenceLength, int accMode,
                                const std::vector<std::vector<int32_t>>& prompts,
                                const std::function<void(int, GenerationState, int32_t)>& emit) {
  const int n = (int)prompts.size();
  if (n < 1) throw Error("empty prompt batch");
  model::InferenceContext infContext(transformer, sequenceLength, accMode, std::max(8, n), n);
  const int seqLen = infContext.SequenceLength;
  for (const auto& p : prompts)
    if ((int)p.size() >= seqLen)
      throw Error("context SequenceLength " + std::to_string(seqLen) + " must be higher than prompt tokens length " + std::to_string(p.size()));
  std::vector<int32_t> cur(n), pos(n);
  std::vector<char> done(n, 0);
  int n_done = 0;
  auto record = [&](int i, int32_t tok, int curPos) {   // :218-248
    cur[i] = tok;
    pos[i] = curPos;
    bool eos = false;
    for (int32_t s : vocab.StopTokenIds) eos |= (s == tok);
    GenerationState st = eos ? GSFinishedByReachingEOS : (curPos + 1 == seqLen ? GSFinishedByReachingSeqLen : GSInProgress);
    if (st != GSInProgress) { done[i] = 1; n_done++; }
    emit(i, st, tok);
  };
  for (int i = 0; i < n; i++) {                           // prefill on sequence i's cache
    infContext.SetActiveSequence(i);
    record(i, infContext.ForwardArgmax(prompts[i], 0), (int)prompts[i].size());
  }
  while (n_done < n) {
    // finished sequences are stepped again at their last position (a no-op for their state)
    const std::vector<int32_t> next = infContext.ForwardBatch(cur, pos);
    for (int i = 0; i < n; i++)
      if (!done[i]) record(i, next[i], pos[i] + 1);
  }
}
}  // namespace inference
}  // namespace lnb_host
