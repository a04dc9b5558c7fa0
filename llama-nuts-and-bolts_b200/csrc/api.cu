// api.cu -- the C-ABI of liblnb.so (include/lnb.h): model / session lifecycle, the fused
// forward pass, the device-resident decode loop and the op-level entry points.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cerrno>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lnb.h"
#include "gemm_tc.cuh"
#include "sdpa_tc.cuh"
#include "gemv.cuh"
#include "kernels.cuh"
#include "seqsum.cuh"
#include "engine.cuh"
#include "engine_batch.cuh"
#include "pth.hpp"
#include "tokenizer.hpp"

namespace lnb {
void build_rope_table(int dim, int end, double theta, bool use_scaled, std::vector<float>& out);
void build_silu_table(std::vector<uint16_t>& out);
}  // namespace lnb

using namespace lnb;

// ------------------------------------------------------------------------------------------
// errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(LNB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

extern "C" const char* lnb_last_error(void) { return g_err.c_str(); }
extern "C" int lnb_version(void) { return LNB_VERSION; }
extern "C" int lnb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(LNB_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (single-GPU use never needs the library)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSum_ = 0, ncclMax_ = 2 };
enum { ncclUint64_ = 5, ncclFloat32_ = 7 };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail(LNB_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                       \
  if (!g_nccl.field) return fail(LNB_ENCCL, "libnccl lacks %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
#define NC(call)                                                                                \
  do {                                                                                          \
    int r_ = (call);                                                                            \
    if (r_ != 0) return fail(LNB_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_));     \
  } while (0)

extern "C" int lnb_nccl_unique_id(void* out128) {
  if (!out128) return fail(LNB_EINVAL, "out128 is NULL");
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return 0;
}

// ------------------------------------------------------------------------------------------
// model
struct LayerW {
  uint16_t *attn_norm = nullptr, *wqkv = nullptr, *wo = nullptr, *ffn_norm = nullptr, *w13 = nullptr, *w2 = nullptr;
  unsigned have = 0;  // bit per checkpoint tensor (9)
};
enum { T_ATTN_NORM = 0, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_EMBD, T_NORM, T_OUTPUT };

struct lnb_model {
  lnb_model_args a;
  int device = 0, tp_rank = 0, tp_size = 1;
  int q_dim = 0, kv_dim = 0, q_l = 0, kv_l = 0, ffn_l = 0, vocab_l = 0;
  uint16_t *tok_embd = nullptr, *norm = nullptr, *output = nullptr;
  bool have_embd = false, have_norm = false, have_output = false, finalized = false;
  std::vector<LayerW> layers;
  float* cis = nullptr;
  int cis_rows = 0;
  uint16_t* silu_tab = nullptr;
  bool cis_set = false, silu_set = false;
  ncclComm_t comm = nullptr;
  void* staging = nullptr;
  size_t staging_bytes = 0;
  int sm_count = 148;
};

static int parse_name(const lnb_model* m, const char* name, int* kind, int* layer) {
  *layer = -1;
  if (!strcmp(name, "tok_embeddings.weight")) { *kind = T_EMBD; return 0; }
  if (!strcmp(name, "norm.weight")) { *kind = T_NORM; return 0; }
  if (!strcmp(name, "output.weight")) { *kind = T_OUTPUT; return 0; }
  int l = -1;
  char rest[96];
  if (sscanf(name, "layers.%d.%95s", &l, rest) != 2 || l < 0 || l >= m->a.n_layers) return -1;
  *layer = l;
  static const char* names[9] = {"attention_norm.weight", "attention.wq.weight", "attention.wk.weight",
                                 "attention.wv.weight",   "attention.wo.weight", "ffn_norm.weight",
                                 "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight"};
  for (int i = 0; i < 9; i++)
    if (!strcmp(rest, names[i])) { *kind = i; return 0; }
  return -1;
}

// full checkpoint shape of a tensor kind (llamatransformer.go:84-105,191,202,273-282,580-586)
static void full_shape(const lnb_model* m, int kind, int64_t* rows, int64_t* cols) {
  const lnb_model_args& a = m->a;
  switch (kind) {
    case T_ATTN_NORM: case T_FFN_NORM: case T_NORM: *rows = a.dim; *cols = 1; break;
    case T_WQ: *rows = m->q_dim; *cols = a.dim; break;
    case T_WK: case T_WV: *rows = m->kv_dim; *cols = a.dim; break;
    case T_WO: *rows = a.dim; *cols = m->q_dim; break;
    case T_W1: case T_W3: *rows = a.ffn_dim; *cols = a.dim; break;
    case T_W2: *rows = a.dim; *cols = a.ffn_dim; break;
    case T_EMBD: case T_OUTPUT: *rows = a.vocab_size; *cols = a.dim; break;
  }
}

// where a checkpoint tensor lands in the HBM layout of this rank
struct Placement {
  uint16_t* dst;
  int64_t row0, col0;  // window of the full tensor owned by this rank
  int rows, cols;      // window size
  bool panel_major;
  int dpanel0, dpanel_stride;
};
static Placement placement(lnb_model* m, int kind, int layer) {
  Placement p{};
  const lnb_model_args& a = m->a;
  const int r = m->tp_rank;
  p.dpanel0 = 0;
  p.dpanel_stride = 1;
  p.panel_major = true;
  LayerW* L = layer >= 0 ? &m->layers[layer] : nullptr;
  switch (kind) {
    case T_ATTN_NORM: p.dst = L->attn_norm; p.rows = 1; p.cols = a.dim; p.panel_major = false; break;
    case T_FFN_NORM: p.dst = L->ffn_norm; p.rows = 1; p.cols = a.dim; p.panel_major = false; break;
    case T_NORM: p.dst = m->norm; p.rows = 1; p.cols = a.dim; p.panel_major = false; break;
    case T_EMBD: p.dst = m->tok_embd; p.rows = a.vocab_size; p.cols = a.dim; p.panel_major = false; break;
    case T_WQ: p.dst = L->wqkv; p.row0 = (int64_t)r * m->q_l; p.rows = m->q_l; p.cols = a.dim; break;
    case T_WK: p.dst = L->wqkv; p.row0 = (int64_t)r * m->kv_l; p.rows = m->kv_l; p.cols = a.dim; p.dpanel0 = m->q_l / 8; break;
    case T_WV: p.dst = L->wqkv; p.row0 = (int64_t)r * m->kv_l; p.rows = m->kv_l; p.cols = a.dim; p.dpanel0 = (m->q_l + m->kv_l) / 8; break;
    case T_WO: p.dst = L->wo; p.col0 = (int64_t)r * m->q_l; p.rows = a.dim; p.cols = m->q_l; break;
    case T_W1: p.dst = L->w13; p.row0 = (int64_t)r * m->ffn_l; p.rows = m->ffn_l; p.cols = a.dim; p.dpanel0 = 0; p.dpanel_stride = 2; break;
    case T_W3: p.dst = L->w13; p.row0 = (int64_t)r * m->ffn_l; p.rows = m->ffn_l; p.cols = a.dim; p.dpanel0 = 2; p.dpanel_stride = 2; break;
    case T_W2: p.dst = L->w2; p.col0 = (int64_t)r * m->ffn_l; p.rows = a.dim; p.cols = m->ffn_l; break;
    case T_OUTPUT: p.dst = m->output; p.row0 = (int64_t)r * m->vocab_l; p.rows = m->vocab_l; p.cols = a.dim; break;
  }
  return p;
}

static void mark_present(lnb_model* m, int kind, int layer) {
  if (kind == T_EMBD) m->have_embd = true;
  else if (kind == T_NORM) m->have_norm = true;
  else if (kind == T_OUTPUT) m->have_output = true;
  else m->layers[layer].have |= 1u << kind;
}

// Pure host arithmetic (no CUDA): which window of checkpoint tensor `name` rank tp_rank owns.
extern "C" int lnb_tp_shard_window(const lnb_model_args* args, const char* name, int tp_rank, int tp_size,
                                   int64_t* row0, int64_t* col0, int64_t* rows, int64_t* cols) {
  if (!args || !name || !row0 || !col0 || !rows || !cols) return fail(LNB_EINVAL, "NULL argument");
  if (tp_size < 1 || tp_rank < 0 || tp_rank >= tp_size) return fail(LNB_EINVAL, "bad tp rank %d / size %d", tp_rank, tp_size);
  const lnb_model_args& a = *args;
  if (a.n_kv_heads % tp_size || a.ffn_dim % tp_size || a.vocab_size % tp_size)
    return fail(LNB_EINVAL, "tp_size %d must divide n_kv_heads, ffn_dim and vocab_size", tp_size);
  lnb_model tmp;
  tmp.a = a;
  tmp.tp_rank = tp_rank;
  tmp.tp_size = tp_size;
  tmp.q_dim = a.n_heads * a.head_dim;
  tmp.kv_dim = a.n_kv_heads * a.head_dim;
  tmp.q_l = tmp.q_dim / tp_size;
  tmp.kv_l = tmp.kv_dim / tp_size;
  tmp.ffn_l = a.ffn_dim / tp_size;
  tmp.vocab_l = a.vocab_size / tp_size;
  tmp.layers.resize(a.n_layers);
  int kind, layer;
  if (parse_name(&tmp, name, &kind, &layer)) return fail(LNB_EINVAL, "unknown tensor name \"%s\"", name);
  Placement p = placement(&tmp, kind, layer);
  int64_t fr, fc;
  full_shape(&tmp, kind, &fr, &fc);
  *row0 = p.row0; *col0 = p.col0;
  if (fc == 1) { *rows = fr; *cols = 1; }  // 1-D tensors are replicated
  else { *rows = p.rows; *cols = p.cols; }
  return 0;
}

extern "C" int lnb_model_create(const lnb_model_args* args, int device, int tp_rank, int tp_size,
                                const void* nccl_unique_id, lnb_model** out) {
  if (!args || !out) return fail(LNB_EINVAL, "args/out is NULL");
  const lnb_model_args& a = *args;
  if (a.dim <= 0 || a.n_layers <= 0 || a.n_heads <= 0 || a.n_kv_heads <= 0 || a.head_dim <= 0 || a.ffn_dim <= 0 ||
      a.vocab_size <= 0 || a.max_seq_len <= 0)
    return fail(LNB_EINVAL, "non-positive model dimension");
  if (a.n_heads % a.n_kv_heads) return fail(LNB_EINVAL, "n_heads %d not a multiple of n_kv_heads %d", a.n_heads, a.n_kv_heads);
  if (a.head_dim % 2) return fail(LNB_EINVAL, "head_dim must be even");
  if (tp_size < 1 || tp_rank < 0 || tp_rank >= tp_size) return fail(LNB_EINVAL, "bad tp rank %d / size %d", tp_rank, tp_size);
  if (a.n_kv_heads % tp_size || a.ffn_dim % tp_size || a.vocab_size % tp_size)
    return fail(LNB_EINVAL, "tp_size %d must divide n_kv_heads, ffn_dim and vocab_size", tp_size);
  const int q_dim = a.n_heads * a.head_dim, kv_dim = a.n_kv_heads * a.head_dim;
  const int q_l = q_dim / tp_size, kv_l = kv_dim / tp_size, ffn_l = a.ffn_dim / tp_size, vocab_l = a.vocab_size / tp_size;
  if (a.dim % 32 || q_l % 32 || kv_l % 32 || ffn_l % 16 || vocab_l % 16)
    return fail(LNB_EINVAL, "dimension not tileable: dim %d q_l %d kv_l %d ffn_l %d vocab_l %d (need multiples of 32/32/32/16/16)",
                a.dim, q_l, kv_l, ffn_l, vocab_l);
  if (tp_size > 1 && !nccl_unique_id) return fail(LNB_EINVAL, "tp_size>1 needs an nccl unique id");
  CU(cudaSetDevice(device));
  lnb_model* m = new lnb_model();
  m->a = a;
  m->device = device;
  m->tp_rank = tp_rank;
  m->tp_size = tp_size;
  m->q_dim = q_dim; m->kv_dim = kv_dim; m->q_l = q_l; m->kv_l = kv_l; m->ffn_l = ffn_l; m->vocab_l = vocab_l;
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
  m->layers.resize(a.n_layers);
  auto alloc16 = [&](uint16_t** p, size_t elems) -> cudaError_t { return cudaMalloc((void**)p, elems * 2); };
  cudaError_t e = cudaSuccess;
  e = alloc16(&m->tok_embd, (size_t)a.vocab_size * a.dim);
  if (e == cudaSuccess) e = alloc16(&m->norm, a.dim);
  if (e == cudaSuccess) e = alloc16(&m->output, (size_t)vocab_l * a.dim);
  for (int l = 0; l < a.n_layers && e == cudaSuccess; l++) {
    LayerW& L = m->layers[l];
    e = alloc16(&L.attn_norm, a.dim);
    if (e == cudaSuccess) e = alloc16(&L.ffn_norm, a.dim);
    if (e == cudaSuccess) e = alloc16(&L.wqkv, (size_t)(q_l + 2 * kv_l) * a.dim);
    if (e == cudaSuccess) e = alloc16(&L.wo, (size_t)a.dim * q_l);
    if (e == cudaSuccess) e = alloc16(&L.w13, (size_t)2 * ffn_l * a.dim);
    if (e == cudaSuccess) e = alloc16(&L.w2, (size_t)a.dim * ffn_l);
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->silu_tab, 65536 * 2);
  if (e != cudaSuccess) {
    lnb_model_destroy(m);
    return fail(LNB_ENOMEM, "cudaMalloc of model weights failed: %s", cudaGetErrorString(e));
  }
  if (tp_size > 1) {
    int rc = nccl_load();
    if (rc) { lnb_model_destroy(m); return rc; }
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, 128);
    int r = g_nccl.CommInitRank(&m->comm, tp_size, id, tp_rank);
    if (r != 0) { lnb_model_destroy(m); return fail(LNB_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
  }
  *out = m;
  return 0;
}

extern "C" int lnb_model_destroy(lnb_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  if (m->comm) g_nccl.CommDestroy(m->comm);
  cudaFree(m->tok_embd); cudaFree(m->norm); cudaFree(m->output);
  for (auto& L : m->layers) {
    cudaFree(L.attn_norm); cudaFree(L.ffn_norm); cudaFree(L.wqkv); cudaFree(L.wo); cudaFree(L.w13); cudaFree(L.w2);
  }
  cudaFree(m->cis); cudaFree(m->silu_tab); cudaFree(m->staging);
  delete m;
  return 0;
}

static int ensure_staging(lnb_model* m, size_t bytes) {
  if (m->staging_bytes >= bytes) return 0;
  if (m->staging) cudaFree(m->staging);
  m->staging = nullptr;
  m->staging_bytes = 0;
  CU(cudaMalloc(&m->staging, bytes));
  m->staging_bytes = bytes;
  return 0;
}

extern "C" int lnb_model_upload_tensor(lnb_model* m, const char* name, const uint16_t* host, const int64_t* shape, int ndim) {
  if (!m || !name || !host || !shape) return fail(LNB_EINVAL, "NULL argument");
  if (m->finalized) return fail(LNB_ESTATE, "model is finalized");
  int kind, layer;
  if (parse_name(m, name, &kind, &layer)) return fail(LNB_EINVAL, "unknown tensor name \"%s\"", name);
  int64_t rows, cols;
  full_shape(m, kind, &rows, &cols);
  // same check as getTensor (src/model/loader.go:183-197)
  bool ok = (cols == 1) ? (ndim == 1 && shape[0] == rows) : (ndim == 2 && shape[0] == rows && shape[1] == cols);
  if (!ok) return fail(LNB_EINVAL, "tensor \"%s\": unexpected shape (expected [%lld%s%lld])", name, (long long)rows,
                       cols == 1 ? "] / [" : ", ", (long long)cols);
  CU(cudaSetDevice(m->device));
  Placement p = placement(m, kind, layer);
  const int64_t ld = (cols == 1) ? rows : cols;
  if (!p.panel_major) {
    CU(cudaMemcpy(p.dst, host, (size_t)p.rows * p.cols * 2, cudaMemcpyHostToDevice));
  } else {
    int rc = ensure_staging(m, (size_t)p.rows * p.cols * 2);
    if (rc) return rc;
    CU(cudaMemcpy2D(m->staging, (size_t)p.cols * 2, host + p.row0 * ld + p.col0, (size_t)ld * 2, (size_t)p.cols * 2,
                    p.rows, cudaMemcpyHostToDevice));
    retile_kernel<<<m->sm_count * 8, 256>>>((const uint16_t*)m->staging, p.cols, 0, 0, p.rows, p.cols, p.dst, p.dpanel0,
                                            p.dpanel_stride);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  mark_present(m, kind, layer);
  return 0;
}

// ---- checkpoint files: .pth reader / writer, params.json --------------------------
struct lnb_pth {
  lnb::PthFile f;
};
struct lnb_pth_writer {
  lnb::PthWriter w;
};

extern "C" int lnb_pth_open(const char* path, lnb_pth** out) {
  if (!path || !out) return fail(LNB_EINVAL, "NULL argument");
  lnb_pth* h = new lnb_pth();
  std::string err;
  if (!h->f.open(path, err)) {
    delete h;
    return fail(LNB_EINVAL, "%s", err.c_str());
  }
  *out = h;
  return 0;
}
extern "C" int lnb_pth_close(lnb_pth* f) {
  delete f;
  return 0;
}
extern "C" int lnb_pth_tensor_count(const lnb_pth* f) {
  if (!f) return fail(LNB_EINVAL, "NULL argument");
  return (int)f->f.tensors().size();
}
extern "C" int lnb_pth_tensor_info(const lnb_pth* f, int index, const char** name, int* dtype, int* ndim, int64_t* shape,
                                   int64_t* file_offset, int64_t* nbytes) {
  if (!f) return fail(LNB_EINVAL, "NULL argument");
  if (index < 0 || index >= (int)f->f.tensors().size()) return fail(LNB_EINVAL, "tensor index %d out of range", index);
  const lnb::PthTensor& t = f->f.tensors()[index];
  if (t.shape.size() > 8) return fail(LNB_EINVAL, "tensor \"%s\" has more than 8 dimensions", t.name.c_str());
  if (name) *name = t.name.c_str();
  if (dtype) *dtype = t.dtype;
  if (ndim) *ndim = (int)t.shape.size();
  if (shape) for (size_t i = 0; i < t.shape.size(); i++) shape[i] = t.shape[i];
  if (file_offset) *file_offset = t.file_offset;
  if (nbytes) *nbytes = t.nbytes;
  return t.contiguous ? 0 : 1;
}
extern "C" const void* lnb_pth_tensor_data(const lnb_pth* f, int index) {
  if (!f || index < 0 || index >= (int)f->f.tensors().size()) {
    fail(LNB_EINVAL, "tensor index %d out of range", index);
    return nullptr;
  }
  return f->f.data(f->f.tensors()[index]);
}

// torch.NewTorchModelReader + Load + the name / shape binding of NewLlamaTransformer, in one pass over the
// mapping: every tensor the architecture names is copied (TP: sliced) from the page cache into HBM.
extern "C" int lnb_model_load_pth(lnb_model* m, const char* path, int* n_uploaded) {
  if (!m || !path) return fail(LNB_EINVAL, "NULL argument");
  if (m->finalized) return fail(LNB_ESTATE, "model is finalized");
  lnb::PthFile f;
  std::string err;
  if (!f.open(path, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  int n = 0;
  for (const lnb::PthTensor& t : f.tensors()) {
    int kind, layer;
    if (parse_name(m, t.name.c_str(), &kind, &layer)) continue;   // e.g. "rope.freqs": the reference never asks for it either
    if (t.dtype != lnb::PTH_BF16)
      return fail(LNB_EINVAL, "tensor \"%s\" is %s; the reference reads torch.BFloat16Storage only (src/torch/types.go:9-21)", t.name.c_str(),
                  lnb::pth_dtype_name(t.dtype));
    if (!t.contiguous) return fail(LNB_EINVAL, "tensor \"%s\" is not contiguous", t.name.c_str());
    if (t.shape.size() > 2) return fail(LNB_EINVAL, "tensor \"%s\": unexpected rank %zu", t.name.c_str(), t.shape.size());
    int rc = lnb_model_upload_tensor(m, t.name.c_str(), (const uint16_t*)f.data(t), t.shape.data(), (int)t.shape.size());
    if (rc) return rc;
    n++;
  }
  if (n_uploaded) *n_uploaded = n;
  return 0;
}

// loadModelArgsFromFile (src/model/modelargs.go:52-64) + the derived fields (llamatransformer.go:73-82,569-577)
extern "C" int lnb_model_args_from_params_json(const char* path, int max_seq_len, lnb_model_args* out) {
  if (!path || !out) return fail(LNB_EINVAL, "NULL argument");
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(LNB_EINVAL, "open %s: %s", path, strerror(errno));
  std::string text;
  char buf[4096];
  size_t k;
  while ((k = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, k);
  fclose(fp);
  lnb::ParamsJson pj;
  std::string err;
  if (!lnb::parse_params_json(text, pj, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  if (pj.dim <= 0 || pj.n_heads <= 0 || pj.multiple_of <= 0) return fail(LNB_EINVAL, "params.json: non-positive dimension");
  lnb_model_args a{};
  a.dim = pj.dim;
  a.n_layers = pj.n_layers;
  a.n_heads = pj.n_heads;
  a.n_kv_heads = pj.n_kv_heads < 0 ? pj.n_heads : pj.n_kv_heads;
  a.head_dim = pj.dim / pj.n_heads;
  int hidden = 4 * pj.dim;
  hidden = (int)(2 * hidden / 3);
  if (pj.ffn_dim_multiplier > -1) hidden = (int)(pj.ffn_dim_multiplier * (double)hidden);
  a.ffn_dim = pj.multiple_of * ((hidden + pj.multiple_of - 1) / pj.multiple_of);
  a.vocab_size = pj.vocab_size;          // -1 when absent: the reference takes it from the tokenizer (loader.go:108-110)
  a.max_seq_len = max_seq_len > 0 ? max_seq_len : 2048;
  a.norm_eps = pj.norm_eps;
  a.rope_theta = pj.rope_theta > 0 ? pj.rope_theta : 500000.0;
  a.use_scaled_rope = pj.use_scaled_rope ? 1 : 0;
  *out = a;
  return 0;
}

extern "C" int lnb_pth_writer_create(const char* path, lnb_pth_writer** out) {
  if (!path || !out) return fail(LNB_EINVAL, "NULL argument");
  lnb_pth_writer* w = new lnb_pth_writer();
  std::string err;
  if (!w->w.open(path, err)) {
    delete w;
    return fail(LNB_EINVAL, "%s", err.c_str());
  }
  *out = w;
  return 0;
}
extern "C" int lnb_pth_writer_add(lnb_pth_writer* w, const char* name, int dtype, const void* data, const int64_t* shape, int ndim) {
  if (!w || !name || (!data && ndim > 0) || (!shape && ndim > 0)) return fail(LNB_EINVAL, "NULL argument");
  if (ndim < 0 || ndim > 8) return fail(LNB_EINVAL, "bad rank %d", ndim);
  std::string err;
  std::vector<int64_t> sh(shape, shape + ndim);
  if (!w->w.add(name, dtype, data, sh, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  return 0;
}
extern "C" int lnb_pth_writer_finish(lnb_pth_writer* w) {
  if (!w) return fail(LNB_EINVAL, "NULL argument");
  std::string err;
  const bool ok = w->w.finish(err);
  delete w;
  return ok ? 0 : fail(LNB_EINVAL, "%s", err.c_str());
}

// ---- tokenizer: tiktoken vocabulary + BPE, host-only ----------------------------------
struct lnb_vocab {
  lnb::Vocab v;
};
extern "C" int lnb_vocab_load(const char* tokenizer_model_path, lnb_vocab** out) {
  if (!tokenizer_model_path || !out) return fail(LNB_EINVAL, "NULL argument");
  lnb_vocab* h = new lnb_vocab();
  std::string err;
  if (!h->v.load(tokenizer_model_path, err)) {
    delete h;
    return fail(LNB_EINVAL, "%s", err.c_str());
  }
  *out = h;
  return 0;
}
extern "C" int lnb_vocab_write_synthetic(const char* path, int n_mergeable) {
  if (!path) return fail(LNB_EINVAL, "NULL argument");
  std::string err;
  if (!lnb::write_synthetic_vocab(path, n_mergeable, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  return 0;
}
extern "C" int lnb_vocab_destroy(lnb_vocab* v) {
  delete v;
  return 0;
}
extern "C" int lnb_vocab_size(const lnb_vocab* v) {
  if (!v) return fail(LNB_EINVAL, "NULL argument");
  return v->v.size();
}
extern "C" int lnb_vocab_token_id(const lnb_vocab* v, const void* token, int token_len, int32_t* id) {
  if (!v || !token || token_len < 0 || !id) return fail(LNB_EINVAL, "bad argument");
  *id = v->v.id_of(std::string((const char*)token, (size_t)token_len));
  return 0;
}
extern "C" int lnb_vocab_token_bytes(const lnb_vocab* v, int32_t id, const void** bytes, int* len) {
  if (!v || !bytes || !len) return fail(LNB_EINVAL, "NULL argument");
  if (id < 0 || id >= v->v.size()) return fail(LNB_EINVAL, "token id %d out of range", id);
  *bytes = v->v.id_to_token[(size_t)id].data();
  *len = (int)v->v.id_to_token[(size_t)id].size();
  return 0;
}
extern "C" int lnb_vocab_special_ids(const lnb_vocab* v, int32_t* bos, int32_t* eos, int32_t* pad, int32_t* stop2) {
  if (!v) return fail(LNB_EINVAL, "NULL argument");
  if (bos) *bos = v->v.bos_id;
  if (eos) *eos = v->v.eos_id;
  if (pad) *pad = v->v.pad_id;
  if (stop2) { stop2[0] = v->v.stop_ids[0]; stop2[1] = v->v.stop_ids[1]; }
  return 0;
}
static int copy_ids(const std::vector<int32_t>& ids, int32_t* out, int cap, int* n_out) {
  if (n_out) *n_out = (int)ids.size();
  if ((int)ids.size() > cap) return fail(LNB_EINVAL, "token buffer too small: need %zu, have %d", ids.size(), cap);
  if (!ids.empty()) memcpy(out, ids.data(), ids.size() * 4);
  return 0;
}
extern "C" int lnb_tokenize_string(const lnb_vocab* v, const char* text, int64_t text_len, int32_t* out, int cap, int* n_out) {
  if (!v || (!text && text_len > 0) || text_len < 0 || (!out && cap > 0)) return fail(LNB_EINVAL, "bad argument");
  std::vector<int32_t> ids;
  v->v.tokenize_string(std::string(text ? text : "", (size_t)text_len), ids);
  return copy_ids(ids, out, cap, n_out);
}
extern "C" int lnb_split_pieces(const char* text, int64_t text_len, int64_t* ends, int cap, int* n_out) {
  if ((!text && text_len > 0) || text_len < 0 || (!ends && cap > 0)) return fail(LNB_EINVAL, "bad argument");
  const std::string t(text ? text : "", (size_t)text_len);
  int n = 0;
  for (size_t p = 0; p < t.size();) {
    p = lnb::Vocab::next_piece(t, p);
    if (n < cap) ends[n] = (int64_t)p;
    n++;
  }
  if (n_out) *n_out = n;
  return n > cap ? fail(LNB_EINVAL, "piece buffer too small: need %d, have %d", n, cap) : 0;
}
extern "C" int lnb_tokenize_prompt(const lnb_vocab* v, const char* const* headers, const char* const* contents, int n_parts,
                                   int32_t* out, int cap, int* n_out) {
  if (!v || n_parts < 0 || (n_parts > 0 && (!headers || !contents)) || (!out && cap > 0)) return fail(LNB_EINVAL, "bad argument");
  std::vector<lnb::PromptPart> parts;
  for (int i = 0; i < n_parts; i++) {
    if (!headers[i] || !contents[i]) return fail(LNB_EINVAL, "NULL prompt part");
    parts.push_back(lnb::PromptPart{headers[i], contents[i]});
  }
  std::vector<int32_t> ids;
  std::string err;
  if (!v->v.tokenize_prompt(parts, ids, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  return copy_ids(ids, out, cap, n_out);
}
extern "C" int lnb_detokenize(const lnb_vocab* v, const int32_t* ids, int n, char* out, int64_t cap, int64_t* n_out) {
  if (!v || (!ids && n > 0) || n < 0 || (!out && cap > 0)) return fail(LNB_EINVAL, "bad argument");
  std::string s;
  if (!v->v.detokenize(ids, n, s)) return fail(LNB_EINVAL, "token id out of range");
  if (n_out) *n_out = (int64_t)s.size();
  if ((int64_t)s.size() > cap) return fail(LNB_EINVAL, "text buffer too small: need %zu", s.size());
  if (!s.empty()) memcpy(out, s.data(), s.size());
  return 0;
}

// ---- synthetic checkpoint ------------------------------------------------------------------
static uint64_t fnv1a64(const char* s) {
  uint64_t h = 0xcbf29ce484222325ULL;
  for (; *s; s++) { h ^= (uint8_t)*s; h *= 0x100000001b3ULL; }
  return h;
}
static void synth_spec_kind(const lnb_model_args& a, int kind, float* scale, float* offset) {
  const float s3 = 1.7320508075688772f;
  *offset = 0.f;
  switch (kind) {
    case T_EMBD: *scale = 1.0f; break;
    case T_ATTN_NORM: case T_FFN_NORM: case T_NORM: *scale = 0.1f; *offset = 1.0f; break;
    case T_W2: *scale = s3 / sqrtf((float)a.ffn_dim); break;
    case T_OUTPUT: *scale = 0.25f * s3 / sqrtf((float)a.dim); break;
    case T_WO: *scale = s3 / sqrtf((float)(a.n_heads * a.head_dim)); break;
    default: *scale = s3 / sqrtf((float)a.dim); break;
  }
}
extern "C" int lnb_synth_spec(const lnb_model_args* args, const char* name, float* scale, float* offset) {
  if (!args || !name || !scale || !offset) return fail(LNB_EINVAL, "NULL argument");
  lnb_model tmp;
  tmp.a = *args;
  int kind, layer;
  if (parse_name(&tmp, name, &kind, &layer)) return fail(LNB_EINVAL, "unknown tensor name \"%s\"", name);
  synth_spec_kind(*args, kind, scale, offset);
  return 0;
}
static void synth_fill_range(uint64_t s, float scale, float offset, int64_t i0, int64_t i1, uint16_t* out) {
  for (int64_t i = i0; i < i1; i++) {
    uint64_t z = s + ((uint64_t)i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    volatile float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    volatile float w = 2.0f * u - 1.0f;
    volatile float v = w * scale;
    volatile float v2 = v + offset;
    float vv = v2;
    uint32_t b;
    memcpy(&b, &vv, 4);
    out[i] = (uint16_t)(b >> 16);
  }
}
extern "C" int lnb_synth_fill_host(uint64_t seed, const char* name, float scale, float offset, int64_t n, uint16_t* out) {
  if (!name || !out) return fail(LNB_EINVAL, "NULL argument");
  const uint64_t s = seed ^ fnv1a64(name);
  // counter-based generator: ranges are independent, so large tensors are filled by a few threads
  const unsigned hw = std::thread::hardware_concurrency();
  const int nt = (n < (1 << 20)) ? 1 : (int)std::min<unsigned>(16u, hw ? hw : 1u);
  if (nt <= 1) {
    synth_fill_range(s, scale, offset, 0, n, out);
    return 0;
  }
  std::vector<std::thread> th;
  const int64_t step = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const int64_t a = (int64_t)t * step, b = std::min<int64_t>(n, a + step);
    if (a < b) th.emplace_back(synth_fill_range, s, scale, offset, a, b, out);
  }
  for (auto& x : th) x.join();
  return 0;
}
extern "C" int lnb_model_init_synthetic(lnb_model* m, uint64_t seed) {
  if (!m) return fail(LNB_EINVAL, "model is NULL");
  if (m->finalized) return fail(LNB_ESTATE, "model is finalized");
  CU(cudaSetDevice(m->device));
  static const char* lnames[9] = {"attention_norm.weight", "attention.wq.weight", "attention.wk.weight",
                                  "attention.wv.weight",   "attention.wo.weight", "ffn_norm.weight",
                                  "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight"};
  auto fill = [&](const char* name, int kind, int layer) -> int {
    int64_t rows, cols;
    full_shape(m, kind, &rows, &cols);
    const int64_t ld = (cols == 1) ? rows : cols;
    Placement p = placement(m, kind, layer);
    float scale, offset;
    synth_spec_kind(m->a, kind, &scale, &offset);
    synth_fill_kernel<<<m->sm_count * 8, 256>>>(seed ^ fnv1a64(name), scale, offset, ld, p.row0, p.col0, p.rows, p.cols,
                                                p.dst, p.panel_major ? 1 : 0, p.dpanel0, p.dpanel_stride);
    CU(cudaGetLastError());
    mark_present(m, kind, layer);
    return 0;
  };
  int rc;
  if ((rc = fill("tok_embeddings.weight", T_EMBD, -1))) return rc;
  if ((rc = fill("norm.weight", T_NORM, -1))) return rc;
  if ((rc = fill("output.weight", T_OUTPUT, -1))) return rc;
  char name[128];
  for (int l = 0; l < m->a.n_layers; l++)
    for (int k = 0; k < 9; k++) {
      snprintf(name, sizeof(name), "layers.%d.%s", l, lnames[k]);
      if ((rc = fill(name, k, l))) return rc;
    }
  CU(cudaDeviceSynchronize());
  return 0;
}

// The synthetic checkpoint as a file the UNMODIFIED reference loader can read (SURVEY 8d "Synthetic weights":
// "same bytes feed oracle, GPU path and (if ever available) the Go binary via a .pth writer").  Tensor order
// follows Meta's consolidated.00.pth (tok_embeddings, layers.*, norm, output); host-only, no CUDA call.
extern "C" int lnb_pth_write_synthetic(const char* path, const lnb_model_args* args, uint64_t seed) {
  if (!path || !args) return fail(LNB_EINVAL, "NULL argument");
  const lnb_model_args& a = *args;
  if (a.dim <= 0 || a.n_layers <= 0 || a.n_heads <= 0 || a.n_kv_heads <= 0 || a.head_dim <= 0 || a.ffn_dim <= 0 || a.vocab_size <= 0)
    return fail(LNB_EINVAL, "non-positive model dimension");
  lnb_model tmp;
  tmp.a = a;
  tmp.q_dim = a.n_heads * a.head_dim;
  tmp.kv_dim = a.n_kv_heads * a.head_dim;
  lnb::PthWriter w;
  std::string err;
  if (!w.open(path, err)) return fail(LNB_EINVAL, "%s", err.c_str());
  std::vector<uint16_t> buf;
  auto emit = [&](const char* name, int kind) -> int {
    int64_t rows, cols;
    full_shape(&tmp, kind, &rows, &cols);
    float scale, offset;
    synth_spec_kind(a, kind, &scale, &offset);
    buf.resize((size_t)(rows * cols));
    int rc = lnb_synth_fill_host(seed, name, scale, offset, rows * cols, buf.data());
    if (rc) return rc;
    std::vector<int64_t> shape;
    shape.push_back(rows);
    if (cols != 1) shape.push_back(cols);
    if (!w.add(name, lnb::PTH_BF16, buf.data(), shape, err)) return fail(LNB_EINVAL, "%s", err.c_str());
    return 0;
  };
  static const char* lnames[9] = {"attention_norm.weight", "attention.wq.weight", "attention.wk.weight",
                                  "attention.wv.weight",   "attention.wo.weight", "ffn_norm.weight",
                                  "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight"};
  int rc;
  if ((rc = emit("tok_embeddings.weight", T_EMBD))) return rc;
  char name[128];
  for (int l = 0; l < a.n_layers; l++)
    for (int k = 0; k < 9; k++) {
      snprintf(name, sizeof(name), "layers.%d.%s", l, lnames[k]);
      if ((rc = emit(name, k))) return rc;
    }
  if ((rc = emit("norm.weight", T_NORM))) return rc;
  if ((rc = emit("output.weight", T_OUTPUT))) return rc;
  if (!w.finish(err)) return fail(LNB_EINVAL, "%s", err.c_str());
  return 0;
}

// ---- tables ---------------------------------------------------------------------------------
extern "C" int lnb_model_set_rope_table(lnb_model* m, const float* cis, int rows) {
  if (!m || !cis || rows <= 0) return fail(LNB_EINVAL, "bad argument");
  if (m->finalized) return fail(LNB_ESTATE, "model is finalized");
  CU(cudaSetDevice(m->device));
  cudaFree(m->cis);
  m->cis = nullptr;
  const size_t bytes = (size_t)rows * (m->a.head_dim / 2) * 2 * sizeof(float);
  CU(cudaMalloc((void**)&m->cis, bytes));
  CU(cudaMemcpy(m->cis, cis, bytes, cudaMemcpyHostToDevice));
  m->cis_rows = rows;
  m->cis_set = true;
  return 0;
}
extern "C" int lnb_model_set_silu_table(lnb_model* m, const uint16_t* tab) {
  if (!m || !tab) return fail(LNB_EINVAL, "bad argument");
  if (m->finalized) return fail(LNB_ESTATE, "model is finalized");
  CU(cudaSetDevice(m->device));
  CU(cudaMemcpy(m->silu_tab, tab, 65536 * 2, cudaMemcpyHostToDevice));
  m->silu_set = true;
  return 0;
}
extern "C" int lnb_model_get_rope_table(lnb_model* m, float* out, int rows) {
  if (!m || !out || !m->cis || rows > m->cis_rows) return fail(LNB_EINVAL, "bad argument / table not built");
  CU(cudaSetDevice(m->device));
  CU(cudaMemcpy(out, m->cis, (size_t)rows * (m->a.head_dim / 2) * 2 * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int lnb_model_get_silu_table(lnb_model* m, uint16_t* out) {
  if (!m || !out) return fail(LNB_EINVAL, "bad argument");
  CU(cudaSetDevice(m->device));
  CU(cudaMemcpy(out, m->silu_tab, 65536 * 2, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int lnb_model_finalize(lnb_model* m) {
  if (!m) return fail(LNB_EINVAL, "model is NULL");
  if (m->finalized) return 0;
  if (!m->have_embd) return fail(LNB_ESTATE, "missing tensor tok_embeddings.weight");
  if (!m->have_norm) return fail(LNB_ESTATE, "missing tensor norm.weight");
  if (!m->have_output) return fail(LNB_ESTATE, "missing tensor output.weight");
  for (int l = 0; l < m->a.n_layers; l++)
    if (m->layers[l].have != 0x1ffu) return fail(LNB_ESTATE, "layer %d: missing tensors (mask 0x%x)", l, m->layers[l].have);
  CU(cudaSetDevice(m->device));
  if (!m->cis_set) {
    std::vector<float> cis;
    const int rows = m->a.max_seq_len * 2;  // llamatransformer.go:109
    build_rope_table(m->a.head_dim, rows, m->a.rope_theta, m->a.use_scaled_rope != 0, cis);
    int rc = lnb_model_set_rope_table(m, cis.data(), rows);
    if (rc) return rc;
  }
  if (!m->silu_set) {
    std::vector<uint16_t> tab;
    build_silu_table(tab);
    int rc = lnb_model_set_silu_table(m, tab.data());
    if (rc) return rc;
  }
  if (m->staging) { cudaFree(m->staging); m->staging = nullptr; m->staging_bytes = 0; }
  CU(cudaDeviceSynchronize());
  m->finalized = true;
  return 0;
}

// ------------------------------------------------------------------------------------------
// GEMV dispatch
// LNB_ACC_STRICT, one activation row.  The accumulation chain is issue-bound on its scheduler
// (~2.4 instructions per k at one instruction per 2 cycles), so every chain warp needs an SMSP to
// itself: ONE CTA per SM (>113 KB of shared memory each) holding 1, 2 or 4 chain warps (warps 1..4
// sit on SMSPs 1,2,3,0; the producer warp 0 sleeps on its mbarrier), picked from N so that the grid
// stays close to one wave.  Two chain warps on one scheduler run at half speed each.
using CfgS1 = GemvCfg<32, 1, 1, 512, 3>;    // N <= 32*148 rows
using CfgS1M = GemvCfg<64, 1, 1, 512, 3>;   // N <= 64*148
using CfgS1W = GemvCfg<128, 1, 1, 256, 3>;  // larger N (w1|w3, LM head): several waves, HBM-bound
using CfgS2 = GemvCfg<32, 1, 2, 256, 4>;
using CfgS8 = GemvCfg<32, 1, 8, 256, 4>;
using CfgF1 = GemvCfg<32, 8, 1, 256, 4>;
using CfgF1L = GemvCfg<32, 8, 1, 512, 3>;  // at most one CTA per SM (wo, w2): larger bulk copies
using CfgF2 = GemvCfg<32, 8, 2, 256, 4>;
using CfgF8 = GemvCfg<32, 8, 8, 256, 4>;

static const size_t kMaxSmem = 227 * 1024;

struct Launcher {
  cudaStream_t stream;
  bool pdl;
  int64_t* counter;
};

template <class Cfg, int PRO, int EPI>
static int launch_gemv_cfg(const Launcher& L, const GemvParams& p) {
  auto kern = gemv_kernel<Cfg, PRO, EPI>;
  const size_t smem = Cfg::smem_bytes(p.K);
  // the opt-in is per function AND per device (context): track it per instantiation per device
  static std::atomic<uint64_t> attr_mask{0};
  {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_mask.load(std::memory_order_relaxed) & bit)) {
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
      attr_mask.fetch_or(bit, std::memory_order_relaxed);
    }
  }
  cudaLaunchConfig_t cfg{};
  const int n_panels = p.N / 8;
  cfg.gridDim = dim3((n_panels + Cfg::kP - 1) / Cfg::kP);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = L.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = L.pdl ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, kern, p));
  if (L.counter) (*L.counter)++;
  return 0;
}

// picks the row-block size: largest MB in {8,2,1} that is useful for M and whose smem fits
template <int PRO, int EPI>
static int launch_gemv(const Launcher& L, int mode, GemvParams p, int M_total) {
  if (p.N % 16 || p.K % 8) return fail(LNB_EINVAL, "gemv: N %d must be a multiple of 16 and K %d of 8", p.N, p.K);  // 16: panel pairs
  if ((EPI == EPI_QKV_ROPE || EPI == EPI_SWIGLU) && p.N % 32) return fail(LNB_EINVAL, "gemv: N %d must be a multiple of 32", p.N);
  const bool strict = (mode == LNB_ACC_STRICT);
  p.strict_norm = strict ? 1 : 0;
  int mb = 1;
  if (M_total >= 3 && (strict ? CfgS8::smem_bytes(p.K) : CfgF8::smem_bytes(p.K)) <= kMaxSmem) mb = 8;
  else if (M_total >= 2 && (strict ? CfgS2::smem_bytes(p.K) : CfgF2::smem_bytes(p.K)) <= kMaxSmem) mb = 2;
  if ((strict ? CfgS1::smem_bytes(p.K) : CfgF1::smem_bytes(p.K)) > kMaxSmem)
    return fail(LNB_EINVAL, "gemv: K %d too large for shared memory", p.K);
  const uint16_t* x0 = p.x;
  uint16_t* ob0 = p.out_bf16;
  float* of0 = p.out_f32;
  const uint16_t* res0 = p.res;
  const float* rs0 = p.rscale;
  const int moff0 = p.m_off;
  LnbDevState* st0 = p.st;
  const int arg_row = p.argmax_row;
  for (int m0 = 0; m0 < M_total; m0 += mb) {
    p.M = (M_total - m0 < mb) ? (M_total - m0) : mb;
    p.x = x0 + (size_t)m0 * p.ldx;
    p.out_bf16 = ob0 ? ob0 + (size_t)m0 * p.ldo : nullptr;
    p.out_f32 = of0 ? of0 + (size_t)m0 * p.ldo : nullptr;
    p.res = res0 ? res0 + (size_t)m0 * p.ldo : nullptr;
    p.rscale = rs0 ? rs0 + m0 : nullptr;
    p.m_off = moff0 + m0;
    if (EPI == EPI_LOGITS) {
      const bool covers = st0 && arg_row >= m0 && arg_row < m0 + p.M;
      p.st = covers ? st0 : nullptr;
      p.argmax_row = covers ? arg_row - m0 : -1;
    }
    int rc;
    if (strict) {
      if (mb == 1 && p.N > 64 * 148 && p.N % 128 == 0 && CfgS1W::smem_bytes(p.K) <= kMaxSmem) rc = launch_gemv_cfg<CfgS1W, PRO, EPI>(L, p);
      else if (mb == 1 && p.N > 32 * 148 && p.N % 64 == 0 && CfgS1M::smem_bytes(p.K) <= kMaxSmem) rc = launch_gemv_cfg<CfgS1M, PRO, EPI>(L, p);
      else if (mb == 1) rc = launch_gemv_cfg<CfgS1, PRO, EPI>(L, p);
      else if (mb == 8) rc = launch_gemv_cfg<CfgS8, PRO, EPI>(L, p);
      else rc = launch_gemv_cfg<CfgS2, PRO, EPI>(L, p);
    } else {
      if (mb == 8) rc = launch_gemv_cfg<CfgF8, PRO, EPI>(L, p);
      else if (mb == 2) rc = launch_gemv_cfg<CfgF2, PRO, EPI>(L, p);
      else if (p.N <= 32 * 148 && CfgF1L::smem_bytes(p.K) <= kMaxSmem) rc = launch_gemv_cfg<CfgF1L, PRO, EPI>(L, p);
      else rc = launch_gemv_cfg<CfgF1, PRO, EPI>(L, p);
    }
    if (rc) return rc;
  }
  return 0;
}

template <typename... Args>
static int launch_simple(const Launcher& L, void (*kern)(Args...), dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = L.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = L.pdl ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, kern, args...));
  if (L.counter) (*L.counter)++;
  return 0;
}

// LNB_ACC_STRICT RMSNorm scale r[row] (the reference's sequential fp32 sum of squares, bit-exact).  Three
// interchangeable kernels, identical bits: 1 = one-thread FADD chain (rms_scale_kernel), 2 = iterative binade
// scan, 3 = one-pass predict / fold / walk (both seqsum.cuh; need a row length that fits seq_scan_shape).
// algo 0 = the model path's choice: LNB_RMS_ALGO=chain|scan|seg, else kRmsDefaultAlgo.
enum { RMS_AUTO = 0, RMS_CHAIN = 1, RMS_SCAN = 2, RMS_SEG = 3, RMS_ENGINE = 4 };   // 4 = the decode engine's in-CTA variant (engine.cuh)
static const int kRmsDefaultAlgo = RMS_SEG;  // measured: 209 tok/s vs 195 (chain) vs 171 (iterative scan), STRICT 8B decode
static int rms_default_algo() {
  static const int a = [] {
    const char* e = getenv("LNB_RMS_ALGO");
    if (e && !strcmp(e, "chain")) return (int)RMS_CHAIN;
    if (e && !strcmp(e, "scan")) return (int)RMS_SCAN;
    if (e && !strcmp(e, "seg")) return (int)RMS_SEG;
    return kRmsDefaultAlgo;
  }();
  return a;
}
static int launch_rms_scale(const Launcher& L, const uint16_t* x, int ldx, float* r, int rows, int D, float eps, int algo = RMS_AUTO) {
  int ch = 0, nt = 0;
  if (algo == RMS_AUTO) algo = rms_default_algo();
  if (algo != RMS_CHAIN && !seq_scan_shape(D, &ch, &nt)) algo = RMS_CHAIN;
  if (algo == RMS_SCAN) {
    switch (ch) {
      case 2: return launch_simple(L, rms_scale_scan_kernel<2>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      case 4: return launch_simple(L, rms_scale_scan_kernel<4>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      case 8: return launch_simple(L, rms_scale_scan_kernel<8>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      default: return launch_simple(L, rms_scale_scan_kernel<16>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
    }
  }
  if (algo == RMS_SEG) {
    switch (ch) {
      case 2: return launch_simple(L, rms_scale_seg_kernel<2>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      case 4: return launch_simple(L, rms_scale_seg_kernel<4>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      case 8: return launch_simple(L, rms_scale_seg_kernel<8>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
      default: return launch_simple(L, rms_scale_seg_kernel<16>, dim3(rows), dim3(nt), 0, x, ldx, r, D, eps);
    }
  }
  return launch_simple(L, rms_scale_kernel, dim3(rows), dim3(128), (size_t)D * 4, x, ldx, r, D, eps);
}

// tensor-core GEMM launcher (prefill): grid = (N/128, ceil(M/128))
template <int EPI>
static int launch_gemm_tc(const Launcher& L, const GemmTcParams& p) {
  if (p.N % TC_BN || p.K % TC_KT) return fail(LNB_EINVAL, "gemm_tc: N %d and K %d must be multiples of 128", p.N, p.K);
  auto kern = gemm_tc_kernel<EPI>;
  static std::atomic<uint64_t> attr_mask{0};   // per device, see launch_gemv_cfg
  {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_mask.load(std::memory_order_relaxed) & bit)) {
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
      attr_mask.fetch_or(bit, std::memory_order_relaxed);
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((p.M + TC_BM - 1) / TC_BM, p.N / TC_BN);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = L.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = L.pdl ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, kern, p));
  if (L.counter) (*L.counter)++;
  return 0;
}
static bool tc_eligible(int M, int N, int K) { return M >= 32 && N % TC_BN == 0 && K % TC_KT == 0; }

// ------------------------------------------------------------------------------------------
// session
__global__ void set_state_kernel(LnbDevState* st, int pos, int n_rows, int next_token, int reset_step) {
  st->pos = pos;
  st->safe_rows = pos;  // rows before this call's first position come from earlier, synchronised calls
  st->n_rows = n_rows;
  st->amax_key = LNB_ARGMAX_EMPTY;
  st->done_ctr = 0;
  if (next_token >= 0) st->next_token = next_token;
  if (reset_step) st->step = 0;
}
__global__ void set_p2p_epoch_kernel(LnbDevState* st, uint32_t e) {
  st->ar_epoch = e;
  st->ar_error = 0;
  st->ar_done2 = 0;
}
// tensor-parallel tail of the LM head: decode the reduced key, advance the decode state
__global__ void publish_kernel(LnbDevState* st, int advance, int32_t* tok_out) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long key = st->amax_key;
  const int32_t tok = (key == LNB_ARGMAX_EMPTY) ? -1 : (int32_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
  st->next_token = tok;
  st->amax_key = LNB_ARGMAX_EMPTY;
  st->done_ctr = 0;
  if (advance) {
    if (tok_out) tok_out[st->step] = tok;
    st->step += 1;
    st->pos += 1;
  }
}

struct lnb_session {
  lnb_model* m = nullptr;
  int seq_len = 0, max_rows = 1, mode = LNB_ACC_FAST;
  cudaStream_t stream = nullptr;
  uint16_t *x = nullptr, *h1 = nullptr, *q = nullptr, *o = nullptr, *mbuf = nullptr;
  float* part = nullptr;
  float* rs = nullptr;  // [max_rows] strict-mode RMSNorm scales
  // tensor-core prefill path (LNB_ACC_FAST, S >= 32): X8-layout activations + raw GEMM outputs
  uint16_t *xn8 = nullptr, *o8 = nullptr, *m8 = nullptr, *qkv_raw = nullptr, *gu = nullptr;
  int mpad = 0;
  // peer-memory all-reduce (tp_size > 1, optional): own region + the peers' mapped regions
  uint8_t* p2p_region = nullptr;
  size_t p2p_bytes = 0;
  void* p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  LnbP2P p2p{};
  bool p2p_ready = false;
  float* logits = nullptr;
  size_t logits_rows = 0;
  float* logits_full = nullptr;  // tp>1: gathered [rows, vocab]
  std::vector<uint16_t*> ck, cv;
  int32_t* d_tokens = nullptr;
  LnbDevState* st = nullptr;
  int32_t* d_tok_out = nullptr;
  int32_t* h_pin = nullptr;
  uint32_t* h_err = nullptr;     // pinned mirror of st->ar_error (peer all-reduce timeout flag)
  // logits of the last forward stay in HBM (lnb_forward_device); the host reads rows / argmaxes on demand
  int kept_rows = 0;             // rows of s->logits that belong to the last forward (0: none)
  int32_t kept_last_token = -1;  // greedy token of the last kept row (fused argmax of the LM-head kernel)
  int64_t generation = 0;        // forward calls so far (a logits handle is valid for one generation)
  float* h_logits_pin = nullptr; // pinned staging for logits read-backs
  size_t h_logits_bytes = 0;
  unsigned long long* d_keys = nullptr;  // [max_rows] per-row argmax keys (tensor-parallel on-demand argmax)
  int layer_limit = 0;
  bool allow_chunked = false;    // S > 1 at startPos > 0 (chunked prefill, an extension the reference cannot express)
  // batched decode (BASELINE config 5): n_seq independent sequences share the weights; caches are
  // [n_seq][seq_len][kv]; single-sequence calls address the cache of `active_seq`
  int n_seq = 1, active_seq = 0;
  int32_t* d_pos_arr = nullptr;   // [n_seq] positions of the current batched step
  int32_t* d_next_arr = nullptr;  // [n_seq] greedy tokens of the current batched step
  bool sdpa_smem_decode = false;  // the whole K/V history of one KV head fits in shared memory
  size_t sdpa_decode_smem = 0;
  int64_t launches = 0;
  // persistent decode engine (engine.cuh): the phase list of one S=1 step in device memory
  EnginePhase* d_phases = nullptr;
  int n_phases = 0, phases_cap = 0;
  unsigned int* d_bar = nullptr;
  // tagged activation vectors of the engine ({tag:16 | bf16:16} words): residual stream x, post-attention stream h1,
  // q|k|v of the step, attention output o, FFN hidden m
  uint32_t *x_t = nullptr, *h1_t = nullptr, *qkv_t = nullptr, *o_t = nullptr, *m_t = nullptr;
  // batch engine (engine_batch.cuh): chunk-major bf16 activations [K/8][8][8] of up to 8 sequences, its phase list, argmax keys
  uint16_t *bx = nullptr, *bh1 = nullptr, *bq = nullptr, *bo = nullptr, *bm = nullptr, *bxn = nullptr;
  int batch_tc = 0;     // FAST batch engine: projections on the tensor cores (batch_engine_kernel<0>)
  BatchPhase* d_bphases = nullptr;
  int n_bphases = 0;
  unsigned long long* d_bkeys = nullptr;
  int bkey_layers = -1;
  const float* bkey_logits = nullptr;
  uint32_t eng_tag = 1;          // next free tag (tags of one launch: [eng_tag, eng_tag + n_steps * n_phases))
  bool last_was_engine = false;  // the residual stream of the last forward lives in x_t
  unsigned long long* d_prof = nullptr;   // LNB_ENGINE_PROF=1: per-CTA cycle counters of the engine's consumer thread 0
  int eng_state = 0;             // 0 = not probed, 1 = usable, -1 = this session uses the kernel chain
  bool eng_single_default = true; // single-sequence S=1 steps use the engine (false: single-GPU FAST, where the kernel chain is faster)
  std::string eng_why;           // why not
  int eng_key_seq = -1, eng_key_layers = -1, eng_key_kind = -2;
  const float* eng_key_logits = nullptr;
  cudaGraphExec_t graph = nullptr;
  bool graph_tried = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::mutex mu;
  int last_rows = 0;
};

static int session_create_impl(lnb_model* m, int seq_len, int max_rows, int acc_mode, int n_seq, lnb_session** out);
extern "C" int lnb_session_create(lnb_model* m, int seq_len, int max_rows, int acc_mode, lnb_session** out) {
  return session_create_impl(m, seq_len, max_rows, acc_mode, 1, out);
}
extern "C" int lnb_session_create_batch(lnb_model* m, int seq_len, int n_seq, int max_rows, int acc_mode, lnb_session** out) {
  if (n_seq < 1 || n_seq > 8) return fail(LNB_EINVAL, "n_seq must be 1..8");
  return session_create_impl(m, seq_len, max_rows < n_seq ? n_seq : max_rows, acc_mode, n_seq, out);
}
static int session_create_impl(lnb_model* m, int seq_len, int max_rows, int acc_mode, int n_seq, lnb_session** out) {
  if (!m || !out) return fail(LNB_EINVAL, "NULL argument");
  if (!m->finalized) return fail(LNB_ESTATE, "model is not finalized");
  if (seq_len <= 0 || max_rows <= 0) return fail(LNB_EINVAL, "seq_len and max_rows must be positive");
  if (seq_len > m->cis_rows) return fail(LNB_EINVAL, "seq_len %d exceeds the RoPE table (%d rows)", seq_len, m->cis_rows);
  if (acc_mode != LNB_ACC_STRICT && acc_mode != LNB_ACC_FAST) return fail(LNB_EINVAL, "bad acc_mode %d", acc_mode);
  if ((size_t)seq_len * 12 + (size_t)m->a.head_dim * 4 > 200 * 1024)
    return fail(LNB_EINVAL, "seq_len %d too long for the decode attention kernel", seq_len);
  CU(cudaSetDevice(m->device));
  CU(cudaFuncSetAttribute(sdpa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU(cudaFuncSetAttribute(sdpa_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
  CU(cudaFuncSetAttribute(sdpa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
  CU(cudaFuncSetAttribute(sdpa_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SX_SMEM));
  lnb_session* s = new lnb_session();
  s->m = m;
  s->seq_len = seq_len;
  s->max_rows = max_rows;
  s->mode = acc_mode;
  s->n_seq = n_seq;
  const lnb_model_args& a = m->a;
  {
    const int n_rep = a.n_heads / a.n_kv_heads;
    const size_t need = (size_t)seq_len * ((a.head_dim + 8) + a.head_dim) * 2 + (size_t)n_rep * ((size_t)a.head_dim * 4 + (size_t)seq_len * 12 + 5 * 8) + 256;
    if (need <= 200 * 1024 && n_rep <= 8 && a.head_dim <= 128 && a.head_dim % 8 == 0) {
      s->sdpa_smem_decode = true;
      s->sdpa_decode_smem = need;
      CU(cudaFuncSetAttribute(sdpa_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
  }
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  auto al = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  al((void**)&s->x, (size_t)max_rows * a.dim * 2);
  al((void**)&s->h1, (size_t)max_rows * a.dim * 2);
  al((void**)&s->q, (size_t)max_rows * m->q_l * 2);
  al((void**)&s->o, (size_t)max_rows * m->q_l * 2);
  al((void**)&s->mbuf, (size_t)max_rows * m->ffn_l * 2);
  al((void**)&s->part, (size_t)max_rows * a.dim * 4);
  al((void**)&s->rs, (size_t)max_rows * 4);
  if (max_rows >= 32 && acc_mode == LNB_ACC_FAST) {
    s->mpad = (max_rows + TC_BM - 1) / TC_BM * TC_BM;
    al((void**)&s->xn8, (size_t)s->mpad * a.dim * 2);
    al((void**)&s->o8, (size_t)s->mpad * m->q_l * 2);
    al((void**)&s->m8, (size_t)s->mpad * m->ffn_l * 2);
    al((void**)&s->qkv_raw, (size_t)max_rows * (m->q_l + 2 * m->kv_l) * 2);
    al((void**)&s->gu, (size_t)max_rows * 2 * m->ffn_l * 2);
  }
  al((void**)&s->d_tokens, (size_t)max_rows * 4);
  al((void**)&s->st, sizeof(LnbDevState));
  al((void**)&s->d_tok_out, (size_t)seq_len * 4);
  s->ck.assign(a.n_layers, nullptr);
  s->cv.assign(a.n_layers, nullptr);
  const size_t cbytes = (size_t)n_seq * seq_len * m->kv_l * 2;
  al((void**)&s->d_pos_arr, 8 * 4);
  al((void**)&s->d_next_arr, 8 * 4);
  for (int l = 0; l < a.n_layers; l++) {
    al((void**)&s->ck[l], cbytes);
    al((void**)&s->cv[l], cbytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->ck[l], 0, cbytes, s->stream);  // ml.Zeros, inferencecontext.go:31-43
    if (e == cudaSuccess) e = cudaMemsetAsync(s->cv[l], 0, cbytes, s->stream);
  }
  if (e == cudaSuccess) e = cudaMemsetAsync(s->st, 0, sizeof(LnbDevState), s->stream);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_pin, (size_t)(max_rows + seq_len + 16) * 4);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_err, 64);
  if (e == cudaSuccess) *s->h_err = 0;
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess) {
    lnb_session_destroy(s);
    return fail(LNB_ENOMEM, "session allocation failed: %s", cudaGetErrorString(e));
  }
  *out = s;
  return 0;
}

extern "C" int lnb_session_destroy(lnb_session* s) {
  if (!s) return 0;
  cudaSetDevice(s->m->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->graph) cudaGraphExecDestroy(s->graph);
  cudaFree(s->x); cudaFree(s->h1); cudaFree(s->q); cudaFree(s->o); cudaFree(s->mbuf); cudaFree(s->part); cudaFree(s->rs);
  cudaFree(s->xn8); cudaFree(s->o8); cudaFree(s->m8); cudaFree(s->qkv_raw); cudaFree(s->gu);
  for (int r = 0; r < 8; r++)
    if (s->p2p_peer[r]) cudaIpcCloseMemHandle(s->p2p_peer[r]);
  cudaFree(s->p2p_region);
  cudaFree(s->d_phases); cudaFree(s->d_bar); cudaFree(s->d_prof);
  cudaFree(s->x_t); cudaFree(s->h1_t); cudaFree(s->qkv_t); cudaFree(s->o_t); cudaFree(s->m_t);
  cudaFree(s->bx); cudaFree(s->bh1); cudaFree(s->bq); cudaFree(s->bo); cudaFree(s->bm); cudaFree(s->bxn); cudaFree(s->d_bphases); cudaFree(s->d_bkeys);
  cudaFree(s->logits); cudaFree(s->logits_full); cudaFree(s->d_tokens); cudaFree(s->st); cudaFree(s->d_tok_out); cudaFree(s->d_pos_arr); cudaFree(s->d_next_arr);
  for (auto p : s->ck) cudaFree(p);
  for (auto p : s->cv) cudaFree(p);
  if (s->h_pin) cudaFreeHost(s->h_pin);
  if (s->h_err) cudaFreeHost(s->h_err);
  if (s->h_logits_pin) cudaFreeHost(s->h_logits_pin);
  cudaFree(s->d_keys);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

// Peer-memory all-reduce setup.  Step 1 on every rank: export (allocates this session's region and
// returns its 64-byte CUDA IPC handle).  The host gathers the tp_size handles (any channel).  Step 2 on
// every rank: import all handles (index = rank).  From then on the decode path of this session pushes
// its Wo / w2 partials and argmax keys straight into the peers' memory instead of calling NCCL.
extern "C" int lnb_session_p2p_export(lnb_session* s, void* handle64) {
  if (!s || !handle64) return fail(LNB_EINVAL, "NULL argument");
  lnb_model* m = s->m;
  if (m->tp_size < 2 || m->tp_size > 8) return fail(LNB_EINVAL, "peer all-reduce needs 2..8 ranks");
  CU(cudaSetDevice(m->device));
  if (!s->p2p_region) {
    const size_t slot = (size_t)s->max_rows * m->a.dim;
    s->p2p_bytes = (size_t)2 * m->tp_size * slot * sizeof(uint2);
    CU(cudaMalloc((void**)&s->p2p_region, s->p2p_bytes));
    CU(cudaMemset(s->p2p_region, 0, s->p2p_bytes));
    CU(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, s->p2p_region));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int lnb_session_p2p_import(lnb_session* s, const void* handles, int n) {
  if (!s || !handles) return fail(LNB_EINVAL, "NULL argument");
  lnb_model* m = s->m;
  if (n != m->tp_size) return fail(LNB_EINVAL, "expected %d handles, got %d", m->tp_size, n);
  if (!s->p2p_region) return fail(LNB_ESTATE, "call lnb_session_p2p_export first");
  CU(cudaSetDevice(m->device));
  std::lock_guard<std::mutex> lk(s->mu);
  for (int r = 0; r < n; r++) {
    uint8_t* base;
    if (r == m->tp_rank) {
      base = s->p2p_region;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const uint8_t*)handles + (size_t)r * 64, 64);
      void* q = nullptr;
      CU(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      s->p2p_peer[r] = q;
      base = (uint8_t*)q;
    }
    s->p2p.data[r] = reinterpret_cast<uint2*>(base);
  }
  s->p2p.rank = m->tp_rank;
  s->p2p.n = n;
  s->p2p.slot_elems = s->max_rows * m->a.dim;
  {  // budget of one in-kernel wait for a peer (a late or dead rank yields LNB_ETIMEOUT instead of a hung GPU)
    const char* e = getenv("LNB_P2P_TIMEOUT_MS");
    const long ms = e ? atol(e) : 1500;
    s->p2p.timeout_ns = ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
  }
  // epochs start at 1 (the region is zero-initialised, so no word carries a live epoch)
  set_p2p_epoch_kernel<<<1, 1, 0, s->stream>>>(s->st, 1u);
  CU(cudaStreamSynchronize(s->stream));
  if (s->graph) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
  s->graph_tried = false;
  s->p2p_ready = true;
  return 0;
}

// the captured decode graph bakes in the active sequence's cache pointers and the layer count
static void drop_graph(lnb_session* s) {
  if (s->graph) {
    cudaSetDevice(s->m->device);
    cudaStreamSynchronize(s->stream);
    cudaGraphExecDestroy(s->graph);
    s->graph = nullptr;
  }
  s->graph_tried = false;
}
// Chunked prefill, an EXTENSION: the reference builds its causal mask as [S,S] and adds it to [32,S,T]
// scores (llamatransformer.go:128-136,469-473), which only broadcasts for startPos 0 -- so it can only prefill a prompt
// in one call.  With this switch a Forward of S > 1 tokens at startPos > 0 is accepted and uses the mask a correct [S,T]
// broadcast would give: query row s (absolute position startPos + s) sees the keys t <= startPos + s.  Every other
// operation is unchanged, so a prompt prefilled in chunks leaves bit-identical caches and logits (LNB_ACC_STRICT).
extern "C" int lnb_session_set_chunked_prefill(lnb_session* s, int on) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  std::lock_guard<std::mutex> lk(s->mu);
  s->allow_chunked = on != 0;
  return 0;
}
extern "C" int lnb_session_set_layer_limit(lnb_session* s, int n) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  std::lock_guard<std::mutex> lk(s->mu);
  if (n != s->layer_limit) drop_graph(s);
  s->layer_limit = n;
  return 0;
}
extern "C" int64_t lnb_session_launch_count(lnb_session* s) { return s ? s->launches : 0; }
extern "C" int lnb_session_sync(lnb_session* s) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  CU(cudaSetDevice(s->m->device));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

// peer all-reduce health: the flag travels with the call's own read-backs (enqueue before the stream sync, test after it)
static int p2p_err_enqueue(lnb_session* s) {
  if (s->p2p_ready) CU(cudaMemcpyAsync(s->h_err, &s->st->ar_error, 4, cudaMemcpyDeviceToHost, s->stream));
  return 0;
}
static int p2p_err_check(lnb_session* s) {
  if (!s->p2p_ready || *s->h_err == 0u) return 0;
  const uint32_t e = *s->h_err;
  return fail(LNB_ETIMEOUT, "peer all-reduce timed out after %llu ms: rank %d never received the words of rank %u for all-reduce #%u "
              "(a peer is late, dead or drives its session differently); the session's peer path is poisoned -- "
              "re-import the handles (lnb_session_p2p_import) or fall back with lnb_session_p2p_disable",
              (unsigned long long)(s->p2p.timeout_ns / 1000000ull), s->m->tp_rank, e & 15u, (e >> 4) & 0xffffffu);
}
extern "C" int lnb_session_p2p_disable(lnb_session* s) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  std::lock_guard<std::mutex> lk(s->mu);
  CU(cudaSetDevice(s->m->device));
  CU(cudaStreamSynchronize(s->stream));
  drop_graph(s);
  s->p2p_ready = false;   // the region stays mapped (peers may still write); every later call uses ncclAllReduce
  *s->h_err = 0;
  return 0;
}

static int ensure_logits(lnb_session* s, size_t rows) {
  if (s->logits_rows >= rows) return 0;
  cudaFree(s->logits);
  cudaFree(s->logits_full);
  s->logits = s->logits_full = nullptr;
  s->logits_rows = 0;
  CU(cudaMalloc((void**)&s->logits, rows * (size_t)s->m->vocab_l * 4));
  if (s->m->tp_size > 1) CU(cudaMalloc((void**)&s->logits_full, rows * (size_t)s->m->a.vocab_size * 4));
  s->logits_rows = rows;
  return 0;
}

// ---- persistent decode engine (engine.cuh) ---------------------------------------------------------------------
static float attn_scale_bf16(int head_dim) {  // dtype.BFloat16fromFloat32(float32(math.Sqrt(float64(HeadDim)))) :464
  float f = (float)sqrt((double)head_dim);
  uint32_t u;
  memcpy(&u, &f, 4);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  return f;
}
// Can this session's S=1 steps run as ONE persistent kernel?  (LNB_ENGINE=0 forces the kernel chain.)
static bool engine_probe(lnb_session* s) {
  if (s->eng_state) return s->eng_state > 0;
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  auto no = [&](const char* why) { s->eng_state = -1; s->eng_why = why; return false; };
  // Default (measured on B200, profiles/r02_*): the engine wins wherever launch count or the serial order dominates --
  // LNB_ACC_STRICT at every N (213 vs 207, 325 vs 282, 449 vs 352 tokens/s at 1 / 2 / 8 GPUs) and every host-driven loop --
  // and ties with the kernel chain in LNB_ACC_FAST (308 vs 330 at 1 GPU, 501 vs 484 at 2, 774 vs 800 at 8).  It is the
  // default except for single-GPU FAST sessions; LNB_ENGINE=1 / 0 forces it on / off.
  const char* e = getenv("LNB_ENGINE");
  if (e && !strcmp(e, "0")) return no("LNB_ENGINE=0");
  // single-sequence steps of single-GPU FAST sessions stay on the kernel chain by default (the batch engine does not care)
  s->eng_single_default = (e && !strcmp(e, "1")) || !(s->mode == LNB_ACC_FAST && m->tp_size == 1);
  const int kmax = std::max(a.dim, std::max(m->q_l, m->ffn_l));
  if ((size_t)kmax * 4 > (size_t)ENG_XMAX) return no("activation vector exceeds the engine's shared-memory work area");
  if (a.dim % 8 || m->q_l % 8 || m->ffn_l % 8) return no("widths must be multiples of 8");
  if ((size_t)a.dim * 8 > (size_t)ENG_XMAX) return no("dim too wide for the engine's RMSNorm prologue");
  const int n_rep = a.n_heads / a.n_kv_heads;
  if (a.head_dim > 128 || a.head_dim % 8 || n_rep > 8) return no("head shape");
  if (eng_sdpa_smem(s->seq_len, a.head_dim, n_rep) > (size_t)ENG_WORK) return no("SequenceLength too long for the in-engine attention phase");
  int ch, nt;
  if (s->mode == LNB_ACC_STRICT && !eng_scan_shape(a.dim, &ch, &nt)) return no("dim does not fit the engine's binade scan");
  if (m->q_l / a.head_dim > m->sm_count) return no("more query heads than SMs");
  int occ = 0;
  cudaError_t ce;
  if (s->mode == LNB_ACC_STRICT) {
    ce = cudaFuncSetAttribute(decode_engine_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENG_SMEM);
    if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_engine_kernel<1>, ENG_THREADS, ENG_SMEM);
  } else {
    ce = cudaFuncSetAttribute(decode_engine_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENG_SMEM);
    if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_engine_kernel<8>, ENG_THREADS, ENG_SMEM);
  }
  if (ce != cudaSuccess || occ < 1) { cudaGetLastError(); return no("the engine kernel does not fit an SM"); }
  if (cudaMalloc((void**)&s->d_bar, 256) != cudaSuccess) { cudaGetLastError(); return no("cudaMalloc"); }
  {
    cudaError_t e2 = cudaMalloc((void**)&s->x_t, (size_t)a.dim * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&s->h1_t, (size_t)a.dim * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&s->qkv_t, (size_t)(m->q_l + 2 * m->kv_l) * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&s->o_t, (size_t)m->q_l * 4);
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&s->m_t, (size_t)m->ffn_l * 4);
    if (e2 != cudaSuccess) { cudaGetLastError(); return no("cudaMalloc"); }
    s->eng_tag = 0;   // forces the zero-fill of the tagged vectors before the first launch
  }
  {
    const char* pe = getenv("LNB_ENGINE_PROF");
    if (pe && strcmp(pe, "0")) {
      const size_t nb = (size_t)m->sm_count * ENG_NPROF * 8;
      if (cudaMalloc((void**)&s->d_prof, nb) == cudaSuccess) cudaMemset(s->d_prof, 0, nb);
      else { s->d_prof = nullptr; cudaGetLastError(); }
    }
  }
  s->eng_state = 1;
  return true;
}
// the engine needs the peer all-reduce under tensor parallelism (NCCL calls cannot be issued from inside a kernel)
static bool engine_capable(lnb_session* s) { return engine_probe(s) && (s->m->tp_size == 1 || s->p2p_ready); }
static bool engine_ok(lnb_session* s) { return engine_capable(s) && s->eng_single_default; }

// LNB_ACC_FAST prompt attention on the tensor cores (sdpa_tc.cuh); LNB_SDPA_TC=0 keeps the FMA-pipe tile kernel
// exp of every bf16 value as f64, per device (sdpa_tc.cuh); NULL on failure
static const double* exp_table_device() {
  static std::mutex mu;
  static double* tab[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!tab[dev]) {
    double* t = nullptr;
    if (cudaMalloc((void**)&t, 65536 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    exp_tab_kernel<<<256, 256>>>(t);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); cudaFree(t); return nullptr; }
    tab[dev] = t;
  }
  return tab[dev];
}
static bool sdpa_tc_enabled() {
  const char* e = getenv("LNB_SDPA_TC");
  return !(e && !strcmp(e, "0"));
}
// default: the software-pipelined form (sdpa_tc2_kernel); LNB_SDPA_TC=1: the serial form (sdpa_tc_kernel)
static bool sdpa_tc_pipelined() {
  const char* e = getenv("LNB_SDPA_TC");
  return !(e && !strcmp(e, "1"));
}
static int eng_kt(int mode, int N, int K, int G) {
  const int per = (N / 8 + G - 1) / G;                      // most panels one CTA owns
  const int PT = mode == LNB_ACC_STRICT ? EngCfg<1>::kPT : EngCfg<8>::kPT;
  const int maxp = std::max(1, std::min(PT, per));
  // the largest of 512 / 256 / 128 / 64 whose tile (maxp panels x kt x 16 B) fits a 32 KB stage: few, large bulk copies, and
  // k-tiles the STRICT chain can unroll completely (eng_chain_tile<NG>)
  int kt = 512;
  while (kt > 64 && maxp * kt * 16 > EngCfg<1>::kStage) kt >>= 1;
  return std::min(kt, K);
}
// kind -1: the full step; 0..4: independent copies of one projection of every layer (kernel-alone timing, plain buffers)
static int engine_build(lnb_session* s, int kind, const float* logits_out) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  const int n_layers = (s->layer_limit > 0 && s->layer_limit < a.n_layers) ? s->layer_limit : a.n_layers;
  if (s->d_phases && s->eng_key_seq == s->active_seq && s->eng_key_layers == n_layers && s->eng_key_kind == kind &&
      s->eng_key_logits == logits_out)
    return 0;
  const int G = m->sm_count, mode = s->mode;
  const size_t cache_off = (size_t)s->active_seq * s->seq_len * m->kv_l;
  const bool full = kind < 0;
  std::vector<EnginePhase> ph;
  // which phase of the step wrote a tagged vector last (consumers wait for exactly that phase's tag)
  int w_x = -1, w_h1 = -1, w_qkv = -1, w_o = -1, w_m = -1;
  auto gemv = [&](int pro, int epi, const uint16_t* W, int N, int K, const uint16_t* x, const uint16_t* norm_w, uint16_t* out, int ldo,
                  const uint16_t* res) {
    EnginePhase e{};
    e.type = EP_GEMV; e.pro = pro; e.epi = epi; e.W = W; e.N = N; e.K = K; e.kt = eng_kt(mode, N, K, G); e.x = x; e.ldx = K;
    e.norm_w = norm_w; e.out_bf16 = out; e.ldo = ldo; e.res = res;
    return e;
  };
  auto input = [&](EnginePhase& e, const uint32_t* vec, int writer) {   // tagged input of the phase being appended
    if (full && writer >= 0) { e.x_t = vec; e.x_delta = (int)ph.size() - writer; }
  };
  auto resid = [&](EnginePhase& e, const uint32_t* vec, int writer) {
    if (full && writer >= 0) { e.res_t = vec; e.res_delta = (int)ph.size() - writer; }
  };
  const bool tp = m->tp_size > 1;
  for (int l = 0; l < n_layers; l++) {
    LayerW& W = m->layers[l];
    if (full || kind == 0) {  // attn_norm -> wq|wk|wv -> RoPE -> KV append
      EnginePhase e = gemv(PRO_RMSNORM, EPI_QKV_ROPE, W.wqkv, m->q_l + 2 * m->kv_l, a.dim, s->x, W.attn_norm, s->q, m->q_l, nullptr);
      e.q_dim = m->q_l; e.kv_dim = m->kv_l; e.cache_k = s->ck[l] + cache_off; e.cache_v = s->cv[l] + cache_off;
      if (full && l == 0) e.flags |= EF_X_TOKEN;
      else input(e, s->x_t, w_x);
      if (full) { e.out_t = s->qkv_t; w_qkv = (int)ph.size(); }
      ph.push_back(e);
    }
    if (full) {
      EnginePhase e{};
      e.type = EP_SDPA; e.kv_dim = m->kv_l; e.q_dim = m->q_l;
      e.cache_k = s->ck[l] + cache_off; e.cache_v = s->cv[l] + cache_off;
      input(e, s->qkv_t, w_qkv);
      e.out_t = s->o_t; w_o = (int)ph.size();
      ph.push_back(e);
    }
    if (full || kind == 1) {  // wo + residual
      EnginePhase e = gemv(PRO_PLAIN, (tp && full) ? EPI_P2P : EPI_RESID, W.wo, a.dim, m->q_l, s->o, nullptr, s->h1, a.dim, s->x);
      input(e, s->o_t, w_o);
      if (full && !tp) {
        if (l == 0) e.flags |= EF_RES_TOKEN; else resid(e, s->x_t, w_x);
        e.out_t = s->h1_t; w_h1 = (int)ph.size();
      }
      ph.push_back(e);
      if (tp && full) {
        EnginePhase r{};
        r.type = EP_REDUCE;
        if (l == 0) r.flags |= EF_RES_TOKEN; else resid(r, s->x_t, w_x);
        r.out_t = s->h1_t; w_h1 = (int)ph.size();
        ph.push_back(r);
      }
    }
    if (full || kind == 2) {  // ffn_norm -> w1|w3 -> SiLU * up
      EnginePhase e = gemv(PRO_RMSNORM, EPI_SWIGLU, W.w13, 2 * m->ffn_l, a.dim, s->h1, W.ffn_norm, s->mbuf, m->ffn_l, nullptr);
      input(e, s->h1_t, w_h1);
      if (full) { e.out_t = s->m_t; w_m = (int)ph.size(); }
      ph.push_back(e);
    }
    if (full || kind == 3) {  // w2 + residual
      EnginePhase e = gemv(PRO_PLAIN, (tp && full) ? EPI_P2P : EPI_RESID, W.w2, a.dim, m->ffn_l, s->mbuf, nullptr, s->x, a.dim, s->h1);
      input(e, s->m_t, w_m);
      if (full && !tp) { resid(e, s->h1_t, w_h1); e.out_t = s->x_t; w_x = (int)ph.size(); }
      ph.push_back(e);
      if (tp && full) {
        EnginePhase r{};
        r.type = EP_REDUCE;
        resid(r, s->h1_t, w_h1);
        r.out_t = s->x_t; w_x = (int)ph.size();
        ph.push_back(r);
      }
    }
  }
  if (full || kind == 4) {
    const int reps = kind == 4 ? 4 : 1;
    for (int i = 0; i < reps; i++) {
      EnginePhase e = gemv(PRO_RMSNORM, EPI_LOGITS, m->output, m->vocab_l, a.dim, s->x, m->norm, nullptr, m->vocab_l, nullptr);
      e.out_f32 = const_cast<float*>(logits_out); e.n_offset = m->tp_rank * m->vocab_l;
      if (full && n_layers > 0) input(e, s->x_t, w_x);
      if (full) e.flags |= EF_GRID_SYNC;
      ph.push_back(e);
    }
    if (full && tp) {
      EnginePhase r{};
      r.type = EP_ARGMAX;
      ph.push_back(r);
    }
  }
  if (!full) ph.back().flags |= EF_GRID_SYNC;   // one barrier per sweep keeps the CTAs of a timing run together
  if ((int)ph.size() > s->phases_cap) {
    cudaFree(s->d_phases);
    s->d_phases = nullptr;
    s->phases_cap = 0;
    CU(cudaMalloc((void**)&s->d_phases, ph.size() * sizeof(EnginePhase)));
    s->phases_cap = (int)ph.size();
  }
  CU(cudaStreamSynchronize(s->stream));   // a running engine may still read the old list
  CU(cudaMemcpy(s->d_phases, ph.data(), ph.size() * sizeof(EnginePhase), cudaMemcpyHostToDevice));
  s->n_phases = (int)ph.size();
  s->eng_key_seq = s->active_seq; s->eng_key_layers = n_layers; s->eng_key_kind = kind; s->eng_key_logits = logits_out;
  return 0;
}
// n_steps S=1 forwards; the first input token and position come from the device state (set_state_kernel).  One launch
// covers up to 65535 / n_phases steps (16-bit activation tags); longer runs continue from the device state.
static int engine_launch(lnb_session* s, int n_steps, bool advance) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  EngineParams P{};
  P.phases = s->d_phases; P.n_phases = s->n_phases; P.st = s->st;
  P.emb = m->tok_embd; P.dim = a.dim; P.head_dim = a.head_dim; P.n_rep = a.n_heads / a.n_kv_heads; P.seq_len = s->seq_len;
  P.cis = m->cis; P.silu_tab = m->silu_tab; P.eps = a.norm_eps; P.attn_scale = attn_scale_bf16(a.head_dim);
  P.strict = s->mode == LNB_ACC_STRICT ? 1 : 0;
  P.tok_out = s->d_tok_out; P.bar_ctr = s->d_bar;
  P.p2p = s->p2p; P.tp = m->tp_size;
  P.timeout_ns = s->p2p.timeout_ns ? s->p2p.timeout_ns : 1500000000ull;
  {
    const char* e = getenv("LNB_P2P_TIMEOUT_MS");
    if (e) P.timeout_ns = atol(e) > 0 ? (unsigned long long)atol(e) * 1000000ull : 0ull;
  }
  P.err_host = s->h_err;
  {
    // L2 prefetch per CTA and projection beyond the ring: 148 x 256 KB = 37 MB of the 126 MB L2 by default (LNB_ENGINE_PF_KB, 0 = off)
    static const int pf_kb = [] { const char* e = getenv("LNB_ENGINE_PF_KB"); return e ? atoi(e) : 0; }();
    P.pf_window = pf_kb > 0 ? (unsigned int)pf_kb * 1024u : 0u;
  }
  P.prof = s->d_prof;
  P.advance = advance ? 1 : 0;
  const int max_steps = std::max(1, 65000 / std::max(1, s->n_phases));
  if (!advance && n_steps > max_steps) return fail(LNB_EINVAL, "engine: too many steps for one launch");
  for (int done = 0; done < n_steps;) {
    const int n = std::min(n_steps - done, max_steps);
    if (s->eng_tag == 0 || (uint64_t)s->eng_tag + (uint64_t)n * s->n_phases > 65535u) {
      // the 16-bit tags wrap: kill every old tag first (0 is never issued)
      CU(cudaMemsetAsync(s->x_t, 0, (size_t)a.dim * 4, s->stream));
      CU(cudaMemsetAsync(s->h1_t, 0, (size_t)a.dim * 4, s->stream));
      CU(cudaMemsetAsync(s->qkv_t, 0, (size_t)(m->q_l + 2 * m->kv_l) * 4, s->stream));
      CU(cudaMemsetAsync(s->o_t, 0, (size_t)m->q_l * 4, s->stream));
      CU(cudaMemsetAsync(s->m_t, 0, (size_t)m->ffn_l * 4, s->stream));
      s->eng_tag = 1;
    }
    P.n_steps = n;
    P.tag_base = s->eng_tag;
    s->eng_tag += (uint32_t)n * (uint32_t)s->n_phases;
    CU(cudaMemsetAsync(s->d_bar, 0, 4, s->stream));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(m->sm_count);
    cfg.blockDim = dim3(ENG_THREADS);
    cfg.dynamicSmemBytes = ENG_SMEM;
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident (they wait for each other), or the launch fails
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (s->mode == LNB_ACC_STRICT) CU(cudaLaunchKernelEx(&cfg, decode_engine_kernel<1>, P));
    else CU(cudaLaunchKernelEx(&cfg, decode_engine_kernel<8>, P));
    s->launches++;
    done += n;
  }
  s->last_was_engine = true;
  return 0;
}
// after a failed stream sync: did the engine trap, and why?
static int engine_fault(lnb_session* s, cudaError_t e) {
  const uint32_t c = s->h_err ? *s->h_err : 0u;
  if ((c >> 28) == 0x8u)
    return fail(LNB_ETIMEOUT, "peer all-reduce timed out inside the decode engine: rank %d never received the words of rank %u for "
                "all-reduce #%u (a peer is late or dead); the kernel trapped, this process's CUDA context is gone (%s)",
                s->m->tp_rank, c & 15u, (c >> 4) & 0xffffffu, cudaGetErrorString(e));
  if ((c >> 28) == 0xCu) return fail(LNB_ECUDA, "decode engine: a CTA never reached grid barrier target %u (%s)", c & 0xffffffu, cudaGetErrorString(e));
  if ((c >> 28) == 0xDu) return fail(LNB_ECUDA, "decode engine: weight stage %u never completed (%s)", c & 0xffffffu, cudaGetErrorString(e));
  if ((c >> 28) == 0xEu) return fail(LNB_ECUDA, "decode engine: the activations tagged %u never arrived (%s)", c & 0xffffu, cudaGetErrorString(e));
  return fail(LNB_ECUDA, "stream synchronize failed: %s", cudaGetErrorString(e));
}

// ---- batch engine (engine_batch.cuh): one launch advances all sequences of a batched session by one token -------------------
static bool batch_engine_ok(lnb_session* s) {
  if (s->n_seq < 2 || !engine_capable(s)) return false;
  const lnb_model_args& a = s->m->a;
  if ((size_t)a.dim * 4 > 16 * 1024) return false;                      // norm weights as f32 in 16 KB of shared memory
  const char* e = getenv("LNB_BATCH_ENGINE");
  if (e && !strcmp(e, "0")) return false;
  return true;
}
static int batch_engine_build(lnb_session* s, const float* logits_out) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  const int n_layers = (s->layer_limit > 0 && s->layer_limit < a.n_layers) ? s->layer_limit : a.n_layers;
  if (s->d_bphases && s->bkey_layers == n_layers && s->bkey_logits == logits_out) return 0;
  if (!s->bx) {
    cudaError_t e = cudaMalloc((void**)&s->bx, (size_t)a.dim * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->bh1, (size_t)a.dim * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->bq, (size_t)m->q_l * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->bo, (size_t)m->q_l * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->bm, (size_t)m->ffn_l * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_bkeys, 64);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->bxn, (size_t)a.dim * 16);
    if (e != cudaSuccess) return fail(LNB_ENOMEM, "batch engine buffers: %s", cudaGetErrorString(e));
    CU(cudaMemset(s->bx, 0, (size_t)a.dim * 16)); CU(cudaMemset(s->bh1, 0, (size_t)a.dim * 16));
    CU(cudaMemset(s->bq, 0, (size_t)m->q_l * 16)); CU(cudaMemset(s->bo, 0, (size_t)m->q_l * 16));
    CU(cudaMemset(s->bm, 0, (size_t)m->ffn_l * 16)); CU(cudaMemset(s->bxn, 0, (size_t)a.dim * 16));
    // FAST: the tensor-core form (every K a multiple of 16; LNB_BATCH_TC=0 keeps the FMA-pipe kernel)
    {
      const char* e2 = getenv("LNB_BATCH_TC");
      s->batch_tc = s->mode != LNB_ACC_STRICT && !(e2 && !strcmp(e2, "0")) && a.dim % 16 == 0 && m->q_l % 16 == 0 && m->ffn_l % 16 == 0;
    }
    if (s->mode == LNB_ACC_STRICT) CU(cudaFuncSetAttribute(batch_engine_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENG_SMEM));
    else if (s->batch_tc) CU(cudaFuncSetAttribute(batch_engine_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM));
    else CU(cudaFuncSetAttribute(batch_engine_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ENG_SMEM));
  }
  const int G = m->sm_count, mode = s->mode;
  const bool tp = m->tp_size > 1;
  const bool tc = s->batch_tc != 0;
  std::vector<BatchPhase> ph;
  auto gemv = [&](int pro, int epi, const uint16_t* W, int N, int K, const uint16_t* x, const uint16_t* norm_w, uint16_t* out, int ldo,
                  const uint16_t* res) {
    BatchPhase e{};
    e.type = BP_GEMV; e.pro = pro; e.epi = epi; e.W = W; e.N = N; e.K = K; e.kt = eng_kt(mode, N, K, G); e.x = x;
    e.norm_w = norm_w; e.out = out; e.ldo = ldo; e.res = res;
    if (tc) {
      // tensor-core form: 128-row tiles; the k-tile grows as the CTA's share of the rows shrinks (about 32 KB per stage,
      // copies of 2 .. 8 KB); RMSNorm happened in the SCALE phase (x = the normalised row)
      const int per = std::max(1, std::min(16, (N / 8 + G - 1) / G));
      int kt = BT_KT_MAX;
      while (kt > 128 && per * kt * 16 > EngCfg<1>::kStage) kt >>= 1;
      e.kt = std::min(kt, K);
      if (pro == PRO_RMSNORM) { e.pro = PRO_PLAIN; e.x = s->bxn; e.norm_w = nullptr; }
    }
    return e;
  };
  auto scale = [&](const uint16_t* x, bool from_token, const uint16_t* norm_w) {
    BatchPhase e{};
    e.type = BP_SCALE; e.K = a.dim; e.x = x; e.out = s->bx; e.flags = from_token ? EF_X_TOKEN : 0;
    if (tc) { e.norm_w = norm_w; e.xn = s->bxn; }
    return e;
  };
  auto reduce = [&](const uint16_t* res, uint16_t* out) {
    BatchPhase e{};
    e.type = BP_REDUCE; e.res = res; e.out = out;
    return e;
  };
  for (int l = 0; l < n_layers; l++) {
    LayerW& W = m->layers[l];
    ph.push_back(scale(s->bx, l == 0, W.attn_norm));
    {
      BatchPhase e = gemv(PRO_RMSNORM, EPI_QKV_ROPE, W.wqkv, m->q_l + 2 * m->kv_l, a.dim, s->bx, W.attn_norm, s->bq, m->q_l, nullptr);
      e.q_dim = m->q_l; e.kv_dim = m->kv_l; e.cache_k = s->ck[l]; e.cache_v = s->cv[l];
      ph.push_back(e);
    }
    {
      BatchPhase e{};
      e.type = BP_SDPA; e.x = s->bq; e.out = s->bo; e.q_dim = m->q_l; e.kv_dim = m->kv_l; e.cache_k = s->ck[l]; e.cache_v = s->cv[l];
      ph.push_back(e);
    }
    ph.push_back(gemv(PRO_PLAIN, tp ? EPI_P2P : EPI_RESID, W.wo, a.dim, m->q_l, s->bo, nullptr, s->bh1, a.dim, s->bx));
    if (tp) ph.push_back(reduce(s->bx, s->bh1));
    ph.push_back(scale(s->bh1, false, W.ffn_norm));
    ph.push_back(gemv(PRO_RMSNORM, EPI_SWIGLU, W.w13, 2 * m->ffn_l, a.dim, s->bh1, W.ffn_norm, s->bm, m->ffn_l, nullptr));
    ph.push_back(gemv(PRO_PLAIN, tp ? EPI_P2P : EPI_RESID, W.w2, a.dim, m->ffn_l, s->bm, nullptr, s->bx, a.dim, s->bh1));
    if (tp) ph.push_back(reduce(s->bh1, s->bx));
  }
  ph.push_back(scale(s->bx, n_layers == 0, m->norm));
  {
    BatchPhase e = gemv(PRO_RMSNORM, EPI_LOGITS, m->output, m->vocab_l, a.dim, s->bx, m->norm, nullptr, m->vocab_l, nullptr);
    e.out_f32 = const_cast<float*>(logits_out); e.n_offset = m->tp_rank * m->vocab_l;
    ph.push_back(e);
  }
  {
    BatchPhase e{};
    e.type = BP_TOKENS;
    ph.push_back(e);
  }
  cudaFree(s->d_bphases);
  s->d_bphases = nullptr;
  CU(cudaMalloc((void**)&s->d_bphases, ph.size() * sizeof(BatchPhase)));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaMemcpy(s->d_bphases, ph.data(), ph.size() * sizeof(BatchPhase), cudaMemcpyHostToDevice));
  s->n_bphases = (int)ph.size();
  s->bkey_layers = n_layers; s->bkey_logits = logits_out;
  return 0;
}
// tokens / positions of the n sequences are already in s->d_tokens / s->d_pos_arr
static int batch_engine_launch(lnb_session* s, int n) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  BatchParams P{};
  P.phases = s->d_bphases; P.n_phases = s->n_bphases; P.n = n; P.st = s->st;
  P.emb = m->tok_embd; P.dim = a.dim; P.head_dim = a.head_dim; P.n_rep = a.n_heads / a.n_kv_heads; P.seq_len = s->seq_len;
  P.cis = m->cis; P.silu_tab = m->silu_tab; P.eps = a.norm_eps; P.attn_scale = attn_scale_bf16(a.head_dim);
  P.strict = s->mode == LNB_ACC_STRICT ? 1 : 0;
  P.tokens = s->d_tokens; P.pos_arr = s->d_pos_arr; P.cache_seq_stride = (long long)s->seq_len * m->kv_l;
  P.rscale = s->rs; P.keys = s->d_bkeys; P.next_arr = s->d_next_arr; P.bar_ctr = s->d_bar;
  P.p2p = s->p2p; P.tp = m->tp_size;
  P.timeout_ns = s->p2p.timeout_ns ? s->p2p.timeout_ns : 1500000000ull;
  {
    const char* e = getenv("LNB_P2P_TIMEOUT_MS");
    if (e) P.timeout_ns = atol(e) > 0 ? (unsigned long long)atol(e) * 1000000ull : 0ull;
  }
  P.err_host = s->h_err;
  P.prof = s->d_prof;
  *s->h_err = 0;
  CU(cudaMemsetAsync(s->d_bar, 0, 4, s->stream));
  CU(cudaMemsetAsync(s->d_bkeys, 0, 64, s->stream));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m->sm_count);
  cfg.blockDim = dim3(ENG_THREADS);
  cfg.dynamicSmemBytes = s->batch_tc ? BT_SMEM : ENG_SMEM;
  cfg.stream = s->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (s->mode == LNB_ACC_STRICT) CU(cudaLaunchKernelEx(&cfg, batch_engine_kernel<1>, P));
  else if (s->batch_tc) CU(cudaLaunchKernelEx(&cfg, batch_engine_kernel<0>, P));
  else CU(cudaLaunchKernelEx(&cfg, batch_engine_kernel<8>, P));
  s->launches++;
  s->last_was_engine = false;
  return 0;
}

static int sync_stream(lnb_session* s) {
  cudaError_t e = cudaStreamSynchronize(s->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return engine_fault(s, e);
  return 0;
}
// one S=1 forward through the engine; the token travels in the device state (set_state_kernel)
static int enqueue_decode_engine(lnb_session* s, int32_t token, int start_pos, int n_steps, bool advance, bool reset_step, const float* logits_out) {
  int rc = engine_build(s, -1, logits_out);
  if (rc) return rc;
  set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, start_pos, 1, token, reset_step ? 1 : 0);
  s->launches++;
  *s->h_err = 0;
  return engine_launch(s, n_steps, advance);
}

// Enqueue one LlamaTransformer.Forward (llamatransformer.go:145-180) for S rows on the
// session stream.  from_state_token: row 0's token is st->next_token (device-driven decode).
// logits_rows: 0 = none stored, 1 = last row, S = all rows.  advance: decode-loop bookkeeping.
static int enqueue_forward(lnb_session* s, int S, bool from_state_token, int logits_rows, bool advance, bool pdl, bool batch = false,
                           bool gather_logits = true) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  Launcher L{s->stream, pdl, &s->launches};
  const int mode = s->mode;
  const int32_t* pos_ptr = &s->st->pos;
  // single-sequence calls work on the cache of the active sequence; a batched step addresses all of them
  const size_t seq_stride = (size_t)s->seq_len * m->kv_l;
  const size_t cache_off = batch ? 0 : (size_t)s->active_seq * seq_stride;
  if (batch && !s->sdpa_smem_decode) return fail(LNB_EINVAL, "batched decode needs SequenceLength small enough for the shared-memory attention kernel");
  int rc;
  rc = launch_simple(L, gather_rows_kernel, dim3(S), dim3(256), 0, (const uint16_t*)m->tok_embd,
                     (const int32_t*)(from_state_token ? nullptr : s->d_tokens), (const LnbDevState*)s->st, s->x, a.dim);
  if (rc) return rc;
  const int n_layers = (s->layer_limit > 0 && s->layer_limit < a.n_layers) ? s->layer_limit : a.n_layers;
  const float scale = [&] {  // dtype.BFloat16fromFloat32(float32(math.Sqrt(float64(HeadDim)))) :464
    float f = (float)sqrt((double)a.head_dim);
    uint32_t u;
    memcpy(&u, &f, 4);
    u &= 0xffff0000u;
    memcpy(&f, &u, 4);
    return f;
  }();
  const size_t sdpa_smem = (size_t)s->seq_len * 12 + (size_t)a.head_dim * 4;
  const bool strict = (mode == LNB_ACC_STRICT);
  auto strict_scale = [&](const uint16_t* xin, int rows) -> int {
    if (!strict) return 0;
    return launch_rms_scale(L, xin, a.dim, s->rs, rows, a.dim, a.norm_eps);
  };
  for (int l = 0; l < n_layers; l++) {
    LayerW& W = m->layers[l];
    if ((rc = strict_scale(s->x, S))) return rc;
    {  // attn_norm -> wq|wk|wv -> RoPE -> KV append            (:222, :297-403)
      GemvParams p{};
      p.rscale = strict ? s->rs : nullptr;
      p.W = W.wqkv; p.N = m->q_l + 2 * m->kv_l; p.K = a.dim;
      p.x = s->x; p.ldx = a.dim; p.norm_w = W.attn_norm; p.eps = a.norm_eps;
      p.out_bf16 = s->q; p.ldo = m->q_l;
      p.q_dim = m->q_l; p.kv_dim = m->kv_l; p.head_dim = a.head_dim;
      p.cache_k = s->ck[l] + cache_off; p.cache_v = s->cv[l] + cache_off; p.cis = m->cis; p.pos_ptr = pos_ptr;
      if (batch) { p.pos_arr = s->d_pos_arr; p.cache_seq_stride = (long long)seq_stride; }
      if ((rc = launch_gemv<PRO_RMSNORM, EPI_QKV_ROPE>(L, mode, p, S))) return rc;
    }
    if ((S == 1 || batch) && s->sdpa_smem_decode) {
      const int n_rep = a.n_heads / a.n_kv_heads;
      rc = launch_simple(L, sdpa_decode_kernel, dim3(m->kv_l / a.head_dim, batch ? S : 1), dim3(128 * n_rep), s->sdpa_decode_smem,
                         (const uint16_t*)s->q, (const uint16_t*)(s->ck[l] + cache_off), (const uint16_t*)(s->cv[l] + cache_off), m->kv_l,
                         n_rep, a.head_dim, s->o, pos_ptr, s->seq_len, mode == LNB_ACC_STRICT ? 1 : 0, scale,
                         (const int32_t*)(batch ? s->d_pos_arr : nullptr), (long long)seq_stride, m->q_l, (const LnbDevState*)s->st);
    } else {
      rc = launch_simple(L, sdpa_kernel, dim3(m->q_l / a.head_dim, S), dim3(128), sdpa_smem, (const uint16_t*)s->q, m->q_l,
                         (const uint16_t*)(s->ck[l] + cache_off), (const uint16_t*)(s->cv[l] + cache_off), m->kv_l,
                         a.n_heads / a.n_kv_heads, a.head_dim, s->o, m->q_l, pos_ptr, 0, S, S > 1 ? 1 : 0,
                         mode == LNB_ACC_STRICT ? 1 : 0, scale, 0);
    }
    if (rc) return rc;
    {  // wo + residual                                             (:522, :232)
      GemvParams p{};
      p.W = W.wo; p.N = a.dim; p.K = m->q_l; p.x = s->o; p.ldx = m->q_l; p.ldo = a.dim;
      if (m->tp_size == 1) {
        p.out_bf16 = s->h1; p.res = s->x;
        if ((rc = launch_gemv<PRO_PLAIN, EPI_RESID>(L, mode, p, S))) return rc;
      } else {
        if (s->p2p_ready) {
          p.p2p = s->p2p; p.st = s->st;
          if ((rc = launch_gemv<PRO_PLAIN, EPI_P2P>(L, mode, p, S))) return rc;
          if ((rc = launch_simple(L, p2p_reduce_resid_kernel, dim3(std::min(16, (S * a.dim + 255) / 256)), dim3(256), 0, s->p2p, s->st,
                                  (const uint16_t*)s->x, s->h1, S * a.dim)))
            return rc;
        } else {
          p.out_f32 = s->part;
          if ((rc = launch_gemv<PRO_PLAIN, EPI_F32RAW>(L, mode, p, S))) return rc;
          NC(g_nccl.AllReduce(s->part, s->part, (size_t)S * a.dim, ncclFloat32_, ncclSum_, m->comm, s->stream));
          if ((rc = launch_simple(L, resid_from_f32_kernel, dim3(std::min(148, (S * a.dim + 255) / 256)), dim3(256), 0,
                                  (const float*)s->part, (const uint16_t*)s->x, s->h1, (int64_t)S * a.dim)))
            return rc;
        }
      }
    }
    if ((rc = strict_scale(s->h1, S))) return rc;
    {  // ffn_norm -> w1|w3 -> SiLU * up                             (:237, :601-614)
      GemvParams p{};
      p.rscale = strict ? s->rs : nullptr;
      p.W = W.w13; p.N = 2 * m->ffn_l; p.K = a.dim; p.x = s->h1; p.ldx = a.dim; p.norm_w = W.ffn_norm; p.eps = a.norm_eps;
      p.out_bf16 = s->mbuf; p.ldo = m->ffn_l; p.silu_tab = m->silu_tab;
      if ((rc = launch_gemv<PRO_RMSNORM, EPI_SWIGLU>(L, mode, p, S))) return rc;
    }
    {  // w2 + residual                                             (:619, :248)
      GemvParams p{};
      p.W = W.w2; p.N = a.dim; p.K = m->ffn_l; p.x = s->mbuf; p.ldx = m->ffn_l; p.ldo = a.dim;
      if (m->tp_size == 1) {
        p.out_bf16 = s->x; p.res = s->h1;
        if ((rc = launch_gemv<PRO_PLAIN, EPI_RESID>(L, mode, p, S))) return rc;
      } else {
        if (s->p2p_ready) {
          p.p2p = s->p2p; p.st = s->st;
          if ((rc = launch_gemv<PRO_PLAIN, EPI_P2P>(L, mode, p, S))) return rc;
          if ((rc = launch_simple(L, p2p_reduce_resid_kernel, dim3(std::min(16, (S * a.dim + 255) / 256)), dim3(256), 0, s->p2p, s->st,
                                  (const uint16_t*)s->h1, s->x, S * a.dim)))
            return rc;
        } else {
          p.out_f32 = s->part;
          if ((rc = launch_gemv<PRO_PLAIN, EPI_F32RAW>(L, mode, p, S))) return rc;
          NC(g_nccl.AllReduce(s->part, s->part, (size_t)S * a.dim, ncclFloat32_, ncclSum_, m->comm, s->stream));
          if ((rc = launch_simple(L, resid_from_f32_kernel, dim3(std::min(148, (S * a.dim + 255) / 256)), dim3(256), 0,
                                  (const float*)s->part, (const uint16_t*)s->h1, s->x, (int64_t)S * a.dim)))
            return rc;
        }
      }
    }
  }
  {  // output_norm -> output -> f32 logits (+ greedy argmax of the last row)   (:166-175; inference.go:207-216)
    const int rows = (logits_rows > 1) ? S : 1;  // rows of the head actually computed
    const int row0 = S - rows;
    if ((rc = strict_scale(s->x + (size_t)row0 * a.dim, rows))) return rc;
    GemvParams p{};
    p.rscale = strict ? s->rs : nullptr;
    p.W = m->output; p.N = m->vocab_l; p.K = a.dim;
    p.x = s->x + (size_t)row0 * a.dim; p.ldx = a.dim; p.norm_w = m->norm; p.eps = a.norm_eps;
    p.out_f32 = logits_rows > 0 ? s->logits : nullptr; p.ldo = m->vocab_l;
    p.n_offset = m->tp_rank * m->vocab_l;
    p.st = batch ? nullptr : s->st; p.argmax_row = batch ? -1 : rows - 1; p.m_off = 0;
    p.publish = (m->tp_size == 1) ? 1 : 0;
    p.advance = advance ? 1 : 0; p.tok_out = s->d_tok_out;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(L, mode, p, rows))) return rc;
    if (batch) {
      // every row is the last row of its own sequence: one greedy argmax per row on the (gathered) logits
      if (m->tp_size > 1)
        for (int r = 0; r < rows; r++)
          NC(g_nccl.AllGather(s->logits + (size_t)r * m->vocab_l, s->logits_full + (size_t)r * a.vocab_size, m->vocab_l, ncclFloat32_,
                              m->comm, s->stream));
      const float* lg = (m->tp_size > 1) ? s->logits_full : s->logits;
      if ((rc = launch_simple(L, argmax_f32_kernel, dim3(rows), dim3(256), 0, lg, a.vocab_size, s->d_next_arr))) return rc;
    } else if (m->tp_size > 1) {
      if (s->p2p_ready) {
        if ((rc = launch_simple(L, p2p_argmax_kernel, dim3(1), dim3(32), 0, s->p2p, s->st, advance ? 1 : 0, s->d_tok_out))) return rc;
      } else {
        NC(g_nccl.AllReduce(&s->st->amax_key, &s->st->amax_key, 1, ncclUint64_, ncclMax_, m->comm, s->stream));
        if ((rc = launch_simple(L, publish_kernel, dim3(1), dim3(1), 0, s->st, advance ? 1 : 0, s->d_tok_out))) return rc;
      }
      if (logits_rows > 0 && gather_logits)
        for (int r = 0; r < rows; r++)
          NC(g_nccl.AllGather(s->logits + (size_t)r * m->vocab_l, s->logits_full + (size_t)r * a.vocab_size, m->vocab_l,
                              ncclFloat32_, m->comm, s->stream));
    }
  }
  s->last_rows = S;
  s->last_was_engine = false;
  return 0;
}

// Prompt processing on the tensor cores (LNB_ACC_FAST, S >= 32): the same forward pass as
// enqueue_forward with every projection as a tcgen05 GEMM (gemm_tc.cuh) and the elementwise steps
// as separate memory-bound kernels that write the next GEMM's X8 operand directly.
static bool forward_tc_ok(const lnb_session* s, int S) {
  const lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  return s->mode == LNB_ACC_FAST && s->xn8 && S >= 32 && (m->q_l + 2 * m->kv_l) % TC_BN == 0 && a.dim % TC_BN == 0 &&
         (2 * m->ffn_l) % TC_BN == 0 && a.dim % TC_KT == 0 && m->q_l % TC_KT == 0 && m->ffn_l % TC_KT == 0;
}
static int enqueue_forward_tc(lnb_session* s, int S, int logits_rows, bool gather_logits = true) {
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  Launcher L{s->stream, true, &s->launches};
  const int32_t* pos_ptr = &s->st->pos;
  const int Mpad = (S + TC_BM - 1) / TC_BM * TC_BM;
  const int qkv_n = m->q_l + 2 * m->kv_l;
  const size_t tc_cache_off = (size_t)s->active_seq * s->seq_len * m->kv_l;
  int rc;
  if ((rc = launch_simple(L, gather_rows_kernel, dim3(S), dim3(256), 0, (const uint16_t*)m->tok_embd, (const int32_t*)s->d_tokens,
                          (const LnbDevState*)s->st, s->x, a.dim)))
    return rc;
  const int n_layers = (s->layer_limit > 0 && s->layer_limit < a.n_layers) ? s->layer_limit : a.n_layers;
  float scale = (float)sqrt((double)a.head_dim);
  {
    uint32_t u;
    memcpy(&u, &scale, 4);
    u &= 0xffff0000u;
    memcpy(&scale, &u, 4);
  }
  const size_t sdpa_smem = (size_t)s->seq_len * 12 + (size_t)a.head_dim * 4;
  const double* exp_tab = sdpa_tc_enabled() ? exp_table_device() : nullptr;
  const int ew_grid = m->sm_count * 8;
  for (int l = 0; l < n_layers; l++) {
    LayerW& W = m->layers[l];
    if ((rc = launch_simple(L, rmsnorm_x8_kernel, dim3(Mpad), dim3(256), 0, (const uint16_t*)s->x, a.dim, (const uint16_t*)W.attn_norm,
                            s->xn8, S, a.dim, a.norm_eps)))
      return rc;
    {
      GemmTcParams g{};
      g.X8 = s->xn8; g.W = W.wqkv; g.M = S; g.N = qkv_n; g.K = a.dim; g.out_bf16 = s->qkv_raw; g.ldo = qkv_n;
      if ((rc = launch_gemm_tc<TC_EPI_BF16>(L, g))) return rc;
    }
    if ((rc = launch_simple(L, rope_kv_kernel, dim3(ew_grid), dim3(256), 0, (const uint16_t*)s->qkv_raw, qkv_n, s->q, m->q_l, m->kv_l,
                            a.head_dim, s->ck[l] + tc_cache_off, s->cv[l] + tc_cache_off, (const float*)m->cis, pos_ptr, S)))
      return rc;
    if (a.head_dim == SP_HD && exp_tab && sdpa_tc_pipelined()) {
      if ((rc = launch_simple(L, sdpa_tc2_kernel, dim3(m->q_l / a.head_dim, (S + 127) / 128), dim3(256), (size_t)SX_SMEM,
                              (const uint16_t*)s->q, m->q_l, (const uint16_t*)(s->ck[l] + tc_cache_off),
                              (const uint16_t*)(s->cv[l] + tc_cache_off), m->kv_l, a.n_heads / a.n_kv_heads, s->o8, m->q_l, pos_ptr, S, scale,
                              exp_tab)))
        return rc;
    } else if (a.head_dim == SP_HD && exp_tab) {
      if ((rc = launch_simple(L, sdpa_tc_kernel, dim3(m->q_l / a.head_dim, (S + 127) / 128), dim3(256), (size_t)ST_SMEM,
                              (const uint16_t*)s->q, m->q_l, (const uint16_t*)(s->ck[l] + tc_cache_off),
                              (const uint16_t*)(s->cv[l] + tc_cache_off), m->kv_l, a.n_heads / a.n_kv_heads, s->o8, m->q_l, pos_ptr, S, scale,
                              exp_tab)))
        return rc;
    } else if (a.head_dim == SP_HD) {
      if ((rc = launch_simple(L, sdpa_prefill_kernel, dim3(m->q_l / a.head_dim, (S + SP_QB - 1) / SP_QB), dim3(256), (size_t)SP_SMEM,
                              (const uint16_t*)s->q, m->q_l, (const uint16_t*)(s->ck[l] + tc_cache_off),
                              (const uint16_t*)(s->cv[l] + tc_cache_off), m->kv_l, a.n_heads / a.n_kv_heads, s->o8, m->q_l, pos_ptr, S, scale)))
        return rc;
    } else if ((rc = launch_simple(L, sdpa_kernel, dim3(m->q_l / a.head_dim, S), dim3(128), sdpa_smem, (const uint16_t*)s->q, m->q_l,
                                   (const uint16_t*)(s->ck[l] + tc_cache_off), (const uint16_t*)(s->cv[l] + tc_cache_off), m->kv_l,
                                   a.n_heads / a.n_kv_heads, a.head_dim, s->o8, m->q_l, pos_ptr, 0, S, 1, 0, scale, 1)))
      return rc;
    {
      GemmTcParams g{};
      g.X8 = s->o8; g.W = W.wo; g.M = S; g.N = a.dim; g.K = m->q_l; g.ldo = a.dim;
      if (m->tp_size == 1) {
        g.out_bf16 = s->h1; g.res = s->x;
        if ((rc = launch_gemm_tc<TC_EPI_RESID>(L, g))) return rc;
      } else {
        g.out_f32 = s->part;
        if ((rc = launch_gemm_tc<TC_EPI_F32RAW>(L, g))) return rc;
        NC(g_nccl.AllReduce(s->part, s->part, (size_t)S * a.dim, ncclFloat32_, ncclSum_, m->comm, s->stream));
        if ((rc = launch_simple(L, resid_from_f32_kernel, dim3(ew_grid), dim3(256), 0, (const float*)s->part, (const uint16_t*)s->x, s->h1,
                                (int64_t)S * a.dim)))
          return rc;
      }
    }
    if ((rc = launch_simple(L, rmsnorm_x8_kernel, dim3(Mpad), dim3(256), 0, (const uint16_t*)s->h1, a.dim, (const uint16_t*)W.ffn_norm,
                            s->xn8, S, a.dim, a.norm_eps)))
      return rc;
    {
      GemmTcParams g{};
      // SwiGLU in the epilogue: the hidden activations go straight into the X8 operand of the w2 GEMM
      g.X8 = s->xn8; g.W = W.w13; g.M = S; g.N = 2 * m->ffn_l; g.K = a.dim; g.out_bf16 = s->m8; g.ldo = m->ffn_l; g.silu_tab = m->silu_tab;
      if ((rc = launch_gemm_tc<TC_EPI_SWIGLU>(L, g))) return rc;
    }
    {
      GemmTcParams g{};
      g.X8 = s->m8; g.W = W.w2; g.M = S; g.N = a.dim; g.K = m->ffn_l; g.ldo = a.dim;
      if (m->tp_size == 1) {
        g.out_bf16 = s->x; g.res = s->h1;
        if ((rc = launch_gemm_tc<TC_EPI_RESID>(L, g))) return rc;
      } else {
        g.out_f32 = s->part;
        if ((rc = launch_gemm_tc<TC_EPI_F32RAW>(L, g))) return rc;
        NC(g_nccl.AllReduce(s->part, s->part, (size_t)S * a.dim, ncclFloat32_, ncclSum_, m->comm, s->stream));
        if ((rc = launch_simple(L, resid_from_f32_kernel, dim3(ew_grid), dim3(256), 0, (const float*)s->part, (const uint16_t*)s->h1, s->x,
                                (int64_t)S * a.dim)))
          return rc;
      }
    }
  }
  // LM head: all rows on the tensor cores when asked for (and the vocabulary shard tiles), then the
  // last row through the GEMV, which also owns the greedy argmax (identical to the decode path)
  if (logits_rows > 1 && m->vocab_l % TC_BN == 0) {
    if ((rc = launch_simple(L, rmsnorm_x8_kernel, dim3(Mpad), dim3(256), 0, (const uint16_t*)s->x, a.dim, (const uint16_t*)m->norm, s->xn8, S,
                            a.dim, a.norm_eps)))
      return rc;
    GemmTcParams g{};
    g.X8 = s->xn8; g.W = m->output; g.M = S - 1; g.N = m->vocab_l; g.K = a.dim; g.out_f32 = s->logits; g.ldo = m->vocab_l;
    if (S > 1 && (rc = launch_gemm_tc<TC_EPI_F32TRUNC>(L, g))) return rc;
  } else if (logits_rows > 1 && S > 1) {
    GemvParams p{};
    p.W = m->output; p.N = m->vocab_l; p.K = a.dim; p.x = s->x; p.ldx = a.dim; p.norm_w = m->norm; p.eps = a.norm_eps;
    p.out_f32 = s->logits; p.ldo = m->vocab_l; p.st = nullptr; p.argmax_row = -1;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(L, s->mode, p, S - 1))) return rc;
  }
  {
    GemvParams p{};
    p.W = m->output; p.N = m->vocab_l; p.K = a.dim;
    p.x = s->x + (size_t)(S - 1) * a.dim; p.ldx = a.dim; p.norm_w = m->norm; p.eps = a.norm_eps;
    p.out_f32 = logits_rows > 0 ? s->logits + (logits_rows > 1 ? (size_t)(S - 1) * m->vocab_l : 0) : nullptr;
    p.ldo = m->vocab_l; p.n_offset = m->tp_rank * m->vocab_l;
    p.st = s->st; p.argmax_row = 0; p.publish = (m->tp_size == 1) ? 1 : 0; p.advance = 0; p.tok_out = s->d_tok_out;
    if ((rc = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(L, s->mode, p, 1))) return rc;
    if (m->tp_size > 1) {
      NC(g_nccl.AllReduce(&s->st->amax_key, &s->st->amax_key, 1, ncclUint64_, ncclMax_, m->comm, s->stream));
      if ((rc = launch_simple(L, publish_kernel, dim3(1), dim3(1), 0, s->st, 0, s->d_tok_out))) return rc;
      if (logits_rows > 0 && gather_logits) {
        const int rows = logits_rows > 1 ? S : 1;
        for (int r = 0; r < rows; r++)
          NC(g_nccl.AllGather(s->logits + (size_t)r * m->vocab_l, s->logits_full + (size_t)r * a.vocab_size, m->vocab_l, ncclFloat32_,
                              m->comm, s->stream));
      }
    }
  }
  s->last_rows = S;
  return 0;
}

extern "C" int lnb_forward(lnb_session* s, const int32_t* tokens, int S, int start_pos, float* logits, int all_rows,
                           int32_t* argmax_last) {
  if (!s || !tokens) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0) return fail(LNB_EINVAL, "empty token array");  // llamatransformer.go:146-148
  if (S > s->max_rows) return fail(LNB_EINVAL, "S %d exceeds the session's max_rows %d", S, s->max_rows);
  if (start_pos < 0 || start_pos + S > s->seq_len)
    return fail(LNB_EINVAL, "positions [%d, %d) exceed SequenceLength %d", start_pos, start_pos + S, s->seq_len);
  if (S > 1 && start_pos != 0 && !s->allow_chunked)
    return fail(LNB_EINVAL, "S>1 requires startPos 0 (the reference's [S,S] mask does not broadcast to [S,T]); "
                "lnb_session_set_chunked_prefill enables the [S,T]-mask extension");
  for (int i = 0; i < S; i++)
    if (tokens[i] < 0 || tokens[i] >= s->m->a.vocab_size) return fail(LNB_EINVAL, "token id %d out of range", tokens[i]);
  std::lock_guard<std::mutex> lk(s->mu);
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  const int lrows = logits ? (all_rows ? S : 1) : 0;
  if (lrows) {
    int rc = ensure_logits(s, (size_t)(all_rows ? s->max_rows : 1));
    if (rc) return rc;
  }
  int rc;
  if (S == 1 && engine_ok(s)) {
    // decode step: ONE persistent kernel; the token id rides in the launch of set_state_kernel (4 bytes of kernel arguments)
    if ((rc = enqueue_decode_engine(s, tokens[0], start_pos, 1, false, false, lrows ? s->logits : nullptr))) return rc;
    if (lrows && m->tp_size > 1)
      NC(g_nccl.AllGather(s->logits, s->logits_full, m->vocab_l, ncclFloat32_, m->comm, s->stream));
  } else {
    memcpy(s->h_pin, tokens, (size_t)S * 4);
    CU(cudaMemcpyAsync(s->d_tokens, s->h_pin, (size_t)S * 4, cudaMemcpyHostToDevice, s->stream));
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, start_pos, S, -1, 0);
    s->launches++;
    rc = forward_tc_ok(s, S) ? enqueue_forward_tc(s, S, lrows) : enqueue_forward(s, S, false, lrows, false, true);
    if (rc) return rc;
  }
  if (lrows) {
    const float* src = (m->tp_size > 1) ? s->logits_full : s->logits;
    CU(cudaMemcpyAsync(logits, src, (size_t)lrows * m->a.vocab_size * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  int32_t* tokpin = s->h_pin + s->max_rows;
  if (argmax_last) CU(cudaMemcpyAsync(tokpin, &s->st->next_token, 4, cudaMemcpyDeviceToHost, s->stream));
  if ((rc = p2p_err_enqueue(s))) return rc;
  if ((rc = sync_stream(s))) return rc;
  if ((rc = p2p_err_check(s))) return rc;
  if (argmax_last) *argmax_last = *tokpin;
  return 0;
}

// Forward whose logits stay on the device: the host mirror of LlamaTransformer.Forward hands the caller a tensor
// that carries a handle to them (ml.Tensor with a device view) and only lnb_session_logits_read copies rows out.
// inference.generateTokensInternal slices the last row and argmaxes it (inference.go:207-216): with the handle
// that is lnb_session_logits_argmax -- 4 bytes cross PCIe per token instead of S x 513 KB down and 513 KB up again.
// rows_kept: 1 = last row, S = all rows.  generation_out identifies the call (handles of older calls are stale).
extern "C" int lnb_forward_device(lnb_session* s, const int32_t* tokens, int S, int start_pos, int rows_kept,
                                  int32_t* argmax_last, int64_t* generation_out) {
  if (!s || !tokens) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0) return fail(LNB_EINVAL, "empty token array");
  if (rows_kept != 1 && rows_kept != S) return fail(LNB_EINVAL, "rows_kept must be 1 or S");
  // same checks and the same enqueue as lnb_forward; only the D2H of the logits is left out
  if (S > s->max_rows) return fail(LNB_EINVAL, "S %d exceeds the session's max_rows %d", S, s->max_rows);
  if (start_pos < 0 || start_pos + S > s->seq_len)
    return fail(LNB_EINVAL, "positions [%d, %d) exceed SequenceLength %d", start_pos, start_pos + S, s->seq_len);
  if (S > 1 && start_pos != 0 && !s->allow_chunked)
    return fail(LNB_EINVAL, "S>1 requires startPos 0 (the reference's [S,S] mask does not broadcast to [S,T]); "
                "lnb_session_set_chunked_prefill enables the [S,T]-mask extension");
  for (int i = 0; i < S; i++)
    if (tokens[i] < 0 || tokens[i] >= s->m->a.vocab_size) return fail(LNB_EINVAL, "token id %d out of range", tokens[i]);
  std::lock_guard<std::mutex> lk(s->mu);
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  int rc = ensure_logits(s, (size_t)(rows_kept > 1 ? s->max_rows : 1));
  if (rc) return rc;
  s->kept_rows = 0;
  if (S == 1 && engine_ok(s)) {
    if ((rc = enqueue_decode_engine(s, tokens[0], start_pos, 1, false, false, s->logits))) return rc;
  } else {
    memcpy(s->h_pin, tokens, (size_t)S * 4);
    CU(cudaMemcpyAsync(s->d_tokens, s->h_pin, (size_t)S * 4, cudaMemcpyHostToDevice, s->stream));
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, start_pos, S, -1, 0);
    s->launches++;
    rc = forward_tc_ok(s, S) ? enqueue_forward_tc(s, S, rows_kept, false) : enqueue_forward(s, S, false, rows_kept, false, true, false, false);
    if (rc) return rc;
  }
  int32_t* tokpin = s->h_pin + s->max_rows;
  CU(cudaMemcpyAsync(tokpin, &s->st->next_token, 4, cudaMemcpyDeviceToHost, s->stream));
  if ((rc = p2p_err_enqueue(s))) return rc;
  if ((rc = sync_stream(s))) return rc;
  if ((rc = p2p_err_check(s))) return rc;
  if (argmax_last) *argmax_last = *tokpin;
  s->kept_last_token = *tokpin;
  s->kept_rows = rows_kept;
  s->generation++;
  if (generation_out) *generation_out = s->generation;
  return 0;
}
static int logits_handle_check(lnb_session* s, int64_t generation, int row0, int rows) {
  if (generation != s->generation || s->kept_rows == 0)
    return fail(LNB_ESTATE, "logits handle of forward #%lld is stale (the session is at #%lld)", (long long)generation, (long long)s->generation);
  if (row0 < 0 || rows <= 0 || row0 + rows > s->kept_rows) return fail(LNB_EINVAL, "rows [%d, %d) outside the %d kept rows", row0, row0 + rows, s->kept_rows);
  return 0;
}
// rows [row0, row0+rows) of the kept logits -> host [rows, vocab] f32 (pinned staging; tensor-parallel: all-gather first,
// a collective -- every rank must make the same call)
extern "C" int lnb_session_logits_read(lnb_session* s, int64_t generation, int row0, int rows, float* host) {
  if (!s || !host) return fail(LNB_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lk(s->mu);
  int rc = logits_handle_check(s, generation, row0, rows);
  if (rc) return rc;
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  const size_t V = (size_t)m->a.vocab_size, bytes = (size_t)rows * V * 4;
  const float* src = s->logits + (size_t)row0 * m->vocab_l;
  if (m->tp_size > 1) {
    for (int r = 0; r < rows; r++)
      NC(g_nccl.AllGather(s->logits + (size_t)(row0 + r) * m->vocab_l, s->logits_full + (size_t)(row0 + r) * V, m->vocab_l, ncclFloat32_,
                          m->comm, s->stream));
    src = s->logits_full + (size_t)row0 * V;
  }
  if (bytes <= ((size_t)16 << 20)) {
    if (s->h_logits_bytes < bytes) {
      if (s->h_logits_pin) cudaFreeHost(s->h_logits_pin);
      s->h_logits_pin = nullptr; s->h_logits_bytes = 0;
      CU(cudaMallocHost((void**)&s->h_logits_pin, bytes));
      s->h_logits_bytes = bytes;
    }
    CU(cudaMemcpyAsync(s->h_logits_pin, src, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    memcpy(host, s->h_logits_pin, bytes);
  } else {
    CU(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  return 0;
}
// ml.Argmax (operations_impl.go:513-548) of kept rows, on the device: first maximum wins, NaN never.  The last row's
// greedy token was already produced by the LM-head kernel of the forward (fused argmax) and is returned as is; other
// rows run argmax_f32_kernel on the kept logits (tensor-parallel: keys are max-reduced over the ranks, a collective).
__global__ void __launch_bounds__(256) argmax_key_rows_kernel(const float* __restrict__ x, int cols, int ld, int n_offset,
                                                              unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long best[8];
  const float* xr = x + (size_t)blockIdx.x * ld;
  unsigned long long key = LNB_ARGMAX_EMPTY;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float v = xr[c];
    if (v > -3.402823466e+38f) {
      const unsigned long long k2 = argmax_key(v, (uint32_t)(c + n_offset));
      key = k2 > key ? k2 : key;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) key = best[i] > key ? best[i] : key;
    keys[blockIdx.x] = key;
  }
}
extern "C" int lnb_session_logits_argmax(lnb_session* s, int64_t generation, int row0, int rows, int32_t* out) {
  if (!s || !out) return fail(LNB_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lk(s->mu);
  int rc = logits_handle_check(s, generation, row0, rows);
  if (rc) return rc;
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  if (rows == 1 && row0 + 1 == s->kept_rows) {   // the row the generate loop asks for: argmaxed by the forward itself
    out[0] = s->kept_last_token;
    return 0;
  }
  if (!s->d_keys) CU(cudaMalloc((void**)&s->d_keys, (size_t)s->max_rows * 8));
  unsigned long long* hk = reinterpret_cast<unsigned long long*>(s->h_err + 2);  // pinned, 8-byte aligned, 7 words free
  for (int r0 = 0; r0 < rows; r0 += 7) {
    const int n = rows - r0 < 7 ? rows - r0 : 7;
    argmax_key_rows_kernel<<<n, 256, 0, s->stream>>>(s->logits + (size_t)(row0 + r0) * m->vocab_l, m->vocab_l, m->vocab_l,
                                                      m->tp_rank * m->vocab_l, s->d_keys);
    s->launches++;
    if (m->tp_size > 1) NC(g_nccl.AllReduce(s->d_keys, s->d_keys, (size_t)n, ncclUint64_, ncclMax_, m->comm, s->stream));
    CU(cudaMemcpyAsync(hk, s->d_keys, (size_t)n * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    for (int i = 0; i < n; i++)
      out[r0 + i] = (hk[i] == LNB_ARGMAX_EMPTY) ? -1 : (int32_t)(0xffffffffu - (uint32_t)(hk[i] & 0xffffffffull));
  }
  return 0;
}

// ---- batched decode (BASELINE config 5: concurrent prompts) --------------------------------------
extern "C" int lnb_session_set_active_sequence(lnb_session* s, int seq) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  if (seq < 0 || seq >= s->n_seq) return fail(LNB_EINVAL, "sequence %d out of range (session has %d)", seq, s->n_seq);
  std::lock_guard<std::mutex> lk(s->mu);
  if (seq != s->active_seq) drop_graph(s);
  s->active_seq = seq;
  return 0;
}
// One decode step for n independent sequences (rows 0..n-1 = sequences 0..n-1): tokens[i] is fed at
// positions[i] of sequence i.  The weights are streamed ONCE for all rows (the reference would run n
// separate Forward calls on n InferenceContexts; every row sees exactly the arithmetic of its
// own S=1 Forward).  logits: NULL or host [n, vocab] f32; argmax_out[n] = greedy token per sequence.
extern "C" int lnb_forward_batch(lnb_session* s, const int32_t* tokens, const int32_t* positions, int n, float* logits,
                                 int32_t* argmax_out) {
  if (!s || !tokens || !positions) return fail(LNB_EINVAL, "NULL argument");
  if (n < 1 || n > s->n_seq) return fail(LNB_EINVAL, "n %d exceeds the session's %d sequences", n, s->n_seq);
  for (int i = 0; i < n; i++) {
    if (tokens[i] < 0 || tokens[i] >= s->m->a.vocab_size) return fail(LNB_EINVAL, "token id %d out of range", tokens[i]);
    if (positions[i] < 0 || positions[i] >= s->seq_len) return fail(LNB_EINVAL, "position %d exceeds SequenceLength %d", positions[i], s->seq_len);
  }
  std::lock_guard<std::mutex> lk(s->mu);
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  int rc = ensure_logits(s, (size_t)s->max_rows);
  if (rc) return rc;
  memcpy(s->h_pin, tokens, (size_t)n * 4);
  int32_t* ppin = s->h_pin + s->max_rows;
  memcpy(ppin, positions, (size_t)n * 4);
  CU(cudaMemcpyAsync(s->d_tokens, s->h_pin, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_pos_arr, ppin, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
  if (batch_engine_ok(s)) {
    // ONE launch of the batch engine: every weight matrix streamed once for all n sequences
    if ((rc = batch_engine_build(s, logits ? s->logits : nullptr))) return rc;
    if ((rc = batch_engine_launch(s, n))) return rc;
    if (logits && m->tp_size > 1)
      for (int r = 0; r < n; r++)
        NC(g_nccl.AllGather(s->logits + (size_t)r * m->vocab_l, s->logits_full + (size_t)r * m->a.vocab_size, m->vocab_l, ncclFloat32_,
                            m->comm, s->stream));
  } else {
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, 0, n, -1, 0);
    s->launches++;
    rc = enqueue_forward(s, n, false, n > 1 ? n : 2, false, true, true);
    if (rc) return rc;
  }
  if (logits) {
    const float* src = (m->tp_size > 1) ? s->logits_full : s->logits;
    CU(cudaMemcpyAsync(logits, src, (size_t)n * m->a.vocab_size * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  int32_t* npin = s->h_pin + s->max_rows + 16;
  if (argmax_out) CU(cudaMemcpyAsync(npin, s->d_next_arr, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
  if ((rc = p2p_err_enqueue(s))) return rc;
  if ((rc = sync_stream(s))) return rc;
  if ((rc = p2p_err_check(s))) return rc;
  if (argmax_out) memcpy(argmax_out, npin, (size_t)n * 4);
  return 0;
}

extern "C" int lnb_decode_run(lnb_session* s, int32_t first_token, int start_pos, int n_steps, int use_graph,
                              int32_t* tokens_out, float* ms_out) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  if (n_steps <= 0 || start_pos < 0 || start_pos + n_steps > s->seq_len)
    return fail(LNB_EINVAL, "positions [%d, %d) exceed SequenceLength %d", start_pos, start_pos + n_steps, s->seq_len);
  if (first_token < 0 || first_token >= s->m->a.vocab_size) return fail(LNB_EINVAL, "token id %d out of range", first_token);
  std::lock_guard<std::mutex> lk(s->mu);
  CU(cudaSetDevice(s->m->device));
  if (engine_ok(s)) {
    // the whole run is ONE launch of the persistent engine: n_steps forwards, the greedy token fed back on the device
    int rc = engine_build(s, -1, nullptr);
    if (rc) return rc;
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, start_pos, 1, first_token, 1);
    s->launches++;
    *s->h_err = 0;
    CU(cudaEventRecord(s->ev0, s->stream));
    if ((rc = engine_launch(s, n_steps, true))) return rc;
    CU(cudaEventRecord(s->ev1, s->stream));
    CU(cudaMemcpyAsync(s->h_pin + s->max_rows + 8, s->d_tok_out, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, s->stream));
    if ((rc = sync_stream(s))) return rc;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    if (ms_out) *ms_out = ms;
    if (tokens_out) memcpy(tokens_out, s->h_pin + s->max_rows + 8, (size_t)n_steps * 4);
    return use_graph ? 1 : 0;   // (no graph is needed: there is nothing left to replay)
  }
  if (use_graph && !s->graph && !s->graph_tried) {
    s->graph_tried = true;
    cudaGraph_t g = nullptr;
    const int64_t saved = s->launches;
    cudaError_t e = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (e == cudaSuccess) {
      rc = enqueue_forward(s, 1, true, 0, true, true);
      e = cudaStreamEndCapture(s->stream, &g);
    }
    if (e == cudaSuccess && rc == 0 && g) {
      e = cudaGraphInstantiate(&s->graph, g, 0);
      if (e != cudaSuccess) s->graph = nullptr;
    }
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    s->launches = saved;
  }
  const bool graph = use_graph && s->graph;
  set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, start_pos, 1, first_token, 1);
  CU(cudaEventRecord(s->ev0, s->stream));
  const int64_t per_step_before = s->launches;
  for (int i = 0; i < n_steps; i++) {
    if (graph) {
      CU(cudaGraphLaunch(s->graph, s->stream));
    } else {
      int rc = enqueue_forward(s, 1, true, 0, true, true);
      if (rc) return rc;
    }
  }
  if (graph) {
    // count what the graph replays launch: one forward's worth per step
    int64_t per_step = 0;
    {
      const lnb_model_args& a = s->m->a;
      per_step = 2 + (int64_t)a.n_layers * (s->m->tp_size == 1 ? 5 : 7) + (s->m->tp_size == 1 ? 0 : 1) +
                 (s->mode == LNB_ACC_STRICT ? 2 * (int64_t)a.n_layers + 1 : 0);
    }
    s->launches = per_step_before + per_step * n_steps;
  }
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaMemcpyAsync(s->h_pin + s->max_rows + 8, s->d_tok_out, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, s->stream));
  { int rce = p2p_err_enqueue(s); if (rce) return rce; }
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  { int rce = p2p_err_check(s); if (rce) return rce; }
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  if (ms_out) *ms_out = ms;
  if (tokens_out) memcpy(tokens_out, s->h_pin + s->max_rows + 8, (size_t)n_steps * 4);
  return graph ? 1 : 0;
}

// Times one projection kernel of the decode step in isolation: launches it for every layer's
// weights back to back (n_layers * 100+ MB >> L2, so every launch streams from HBM), `reps`
// sweeps, CUDA events on the session stream.  kind: 0 = wq|wk|wv, 1 = wo, 2 = w1|w3, 3 = w2,
// 4 = LM head.  Outputs go to the session's scratch activations (results are discarded).
extern "C" int lnb_session_bench_kernel(lnb_session* s, int kind, int reps, float* ms_per_launch, int64_t* bytes_per_launch,
                                        int* launches) {
  if (!s || reps <= 0 || kind < 0 || kind > 4) return fail(LNB_EINVAL, "bad argument");
  std::lock_guard<std::mutex> lk(s->mu);
  lnb_model* m = s->m;
  const lnb_model_args& a = m->a;
  CU(cudaSetDevice(m->device));
  if (engine_ok(s)) {
    // the projection as the decode step runs it: a phase of the persistent engine (balanced static row split, weights
    // prefetched across the grid barrier).  One launch walks all layers' copies `reps` times; time per phase.
    int rc = engine_build(s, kind, nullptr);
    if (rc) return rc;
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, 0, 1, 0, 0);
    *s->h_err = 0;
    if ((rc = engine_launch(s, 1, false))) return rc;     // warm-up sweep
    set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, 0, 1, 0, 0);
    CU(cudaEventRecord(s->ev0, s->stream));
    if ((rc = engine_launch(s, reps, false))) return rc;
    CU(cudaEventRecord(s->ev1, s->stream));
    if ((rc = sync_stream(s))) return rc;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    const int n = s->n_phases * reps;
    int64_t bytes = 0;
    switch (kind) {
      case 0: bytes = (int64_t)(m->q_l + 2 * m->kv_l) * a.dim * 2; break;
      case 1: bytes = (int64_t)a.dim * m->q_l * 2; break;
      case 2: bytes = (int64_t)2 * m->ffn_l * a.dim * 2; break;
      case 3: bytes = (int64_t)a.dim * m->ffn_l * 2; break;
      default: bytes = (int64_t)m->vocab_l * a.dim * 2; break;
    }
    if (ms_per_launch) *ms_per_launch = ms / (float)n;
    if (bytes_per_launch) *bytes_per_launch = bytes;
    if (launches) *launches = n;
    s->eng_key_kind = -2;   // the next decode call rebuilds the step's phase list
    return 0;
  }
  Launcher L{s->stream, true, &s->launches};
  set_state_kernel<<<1, 1, 0, s->stream>>>(s->st, 0, 1, 0, 0);
  int n_launch = 0;
  int64_t bytes = 0;
  auto sweep = [&]() -> int {
    int rc = 0;
    const int nl = (kind == 4) ? 4 : a.n_layers;  // the LM head is one 1 GB matrix: repeat it
    for (int l = 0; l < nl && !rc; l++) {
      LayerW& W = m->layers[kind == 4 ? 0 : l];
      GemvParams p{};
      p.eps = a.norm_eps;
      p.rscale = (s->mode == LNB_ACC_STRICT) ? s->rs : nullptr;
      switch (kind) {
        case 0:
          p.W = W.wqkv; p.N = m->q_l + 2 * m->kv_l; p.K = a.dim; p.x = s->x; p.ldx = a.dim; p.norm_w = W.attn_norm;
          p.out_bf16 = s->q; p.ldo = m->q_l; p.q_dim = m->q_l; p.kv_dim = m->kv_l; p.head_dim = a.head_dim;
          p.cache_k = s->ck[l]; p.cache_v = s->cv[l]; p.cis = m->cis; p.pos_ptr = &s->st->pos;
          rc = launch_gemv<PRO_RMSNORM, EPI_QKV_ROPE>(L, s->mode, p, 1);
          break;
        case 1:
          p.W = W.wo; p.N = a.dim; p.K = m->q_l; p.x = s->o; p.ldx = m->q_l; p.ldo = a.dim; p.out_bf16 = s->h1; p.res = s->x;
          rc = launch_gemv<PRO_PLAIN, EPI_RESID>(L, s->mode, p, 1);
          break;
        case 2:
          p.W = W.w13; p.N = 2 * m->ffn_l; p.K = a.dim; p.x = s->h1; p.ldx = a.dim; p.norm_w = W.ffn_norm;
          p.out_bf16 = s->mbuf; p.ldo = m->ffn_l; p.silu_tab = m->silu_tab;
          rc = launch_gemv<PRO_RMSNORM, EPI_SWIGLU>(L, s->mode, p, 1);
          break;
        case 3:
          p.W = W.w2; p.N = a.dim; p.K = m->ffn_l; p.x = s->mbuf; p.ldx = m->ffn_l; p.ldo = a.dim; p.out_bf16 = s->h1; p.res = s->x;
          rc = launch_gemv<PRO_PLAIN, EPI_RESID>(L, s->mode, p, 1);
          break;
        case 4:
          p.W = m->output; p.N = m->vocab_l; p.K = a.dim; p.x = s->x; p.ldx = a.dim; p.norm_w = m->norm; p.ldo = m->vocab_l;
          p.st = nullptr; p.argmax_row = -1;
          rc = launch_gemv<PRO_RMSNORM, EPI_LOGITS>(L, s->mode, p, 1);
          break;
      }
      bytes = (int64_t)p.N * p.K * 2;
      n_launch++;
    }
    return rc;
  };
  if (s->mode == LNB_ACC_STRICT) {
    int rc0 = launch_rms_scale(L, (const uint16_t*)s->x, a.dim, s->rs, 1, a.dim, a.norm_eps);
    if (rc0) return rc0;
  }
  int rc = sweep();  // warm-up sweep
  if (rc) return rc;
  n_launch = 0;
  CU(cudaEventRecord(s->ev0, s->stream));
  for (int r = 0; r < reps; r++)
    if ((rc = sweep())) return rc;
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  if (ms_per_launch) *ms_per_launch = ms / (float)n_launch;
  if (bytes_per_launch) *bytes_per_launch = bytes;
  if (launches) *launches = n_launch;
  return 0;
}

// LNB_ENGINE_PROF=1: cycle sums of the engine's consumer thread 0, averaged and maximised over the CTAs, since the last
// call; out[2 * 8] = {mean, max} x {grid barrier, prologue, main loop, combine + epilogue, attention, peer reduce, -, -}
// 1: this session's S=1 steps run as the persistent decode engine; 0: as the kernel chain (lnb_last_error says why)
extern "C" int lnb_session_decode_engine(lnb_session* s) {
  if (!s) return fail(LNB_EINVAL, "session is NULL");
  std::lock_guard<std::mutex> lk(s->mu);
  CU(cudaSetDevice(s->m->device));
  if (engine_ok(s)) return 1;
  g_err = s->eng_state < 0 ? s->eng_why : !s->eng_single_default ? std::string("single-GPU FAST: the kernel chain is the default") : std::string("tensor-parallel session without the peer all-reduce (NCCL collectives cannot run inside a kernel)");
  return 0;
}
extern "C" int lnb_session_engine_profile(lnb_session* s, double* out16) {   // (out: 2 * ENG_NPROF doubles)
  if (!s || !out16) return fail(LNB_EINVAL, "NULL argument");
  if (!s->d_prof) return fail(LNB_ESTATE, "no engine profile (set LNB_ENGINE_PROF=1 before the session's first decode)");
  CU(cudaSetDevice(s->m->device));
  CU(cudaStreamSynchronize(s->stream));
  const int G = s->m->sm_count;
  std::vector<unsigned long long> h((size_t)G * ENG_NPROF);
  CU(cudaMemcpy(h.data(), s->d_prof, h.size() * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemset(s->d_prof, 0, h.size() * 8));
  for (int k = 0; k < ENG_NPROF; k++) {
    double sum = 0, mx = 0;
    for (int b = 0; b < G; b++) {
      const double v = (double)h[(size_t)b * ENG_NPROF + k];
      sum += v;
      mx = std::max(mx, v);
    }
    out16[2 * k] = sum / G;
    out16[2 * k + 1] = mx;
  }
  return 0;
}

extern "C" int lnb_session_read(lnb_session* s, int which, int layer, void* host, int64_t nbytes) {
  if (!s || !host) return fail(LNB_EINVAL, "NULL argument");
  lnb_model* m = s->m;
  CU(cudaSetDevice(m->device));
  const void* src = nullptr;
  int64_t avail = 0;
  switch (which) {
    case LNB_BUF_RESIDUAL:
      if (s->last_was_engine) {   // the engine keeps the residual stream as tagged words: strip the tags
        if (nbytes > (int64_t)m->a.dim * 2) return fail(LNB_EINVAL, "the decode engine keeps one row");
        std::vector<uint32_t> t((size_t)m->a.dim);
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaMemcpy(t.data(), s->x_t, t.size() * 4, cudaMemcpyDeviceToHost));
        uint16_t* o = (uint16_t*)host;
        for (int64_t i = 0; i < nbytes / 2; i++) o[i] = (uint16_t)(t[(size_t)i] & 0xffffu);
        return 0;
      }
      src = s->x; avail = (int64_t)s->max_rows * m->a.dim * 2; break;
    case LNB_BUF_CACHE_K:
    case LNB_BUF_CACHE_V:
      if (layer < 0 || layer >= m->a.n_layers) return fail(LNB_EINVAL, "bad layer %d", layer);
      src = (which == LNB_BUF_CACHE_K ? s->ck[layer] : s->cv[layer]) + (size_t)s->active_seq * s->seq_len * m->kv_l;
      avail = (int64_t)s->seq_len * m->kv_l * 2;
      break;
    case LNB_BUF_LOGITS: src = s->logits; avail = (int64_t)s->logits_rows * m->vocab_l * 4; break;
    default: return fail(LNB_EINVAL, "bad buffer id %d", which);
  }
  if (!src || nbytes > avail) return fail(LNB_EINVAL, "buffer %d holds %lld bytes, %lld requested", which, (long long)avail, (long long)nbytes);
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaMemcpy(host, src, (size_t)nbytes, cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------
// op-level API: host pointers in / out, staged through HBM, same kernels as the model path
// Device scratch for the op-level calls: a grow-only per-thread arena (cudaMalloc / cudaFree per
// call cost more than the kernels they serve).  Buffers are carved with 256-byte alignment and
// the arena is rewound at the start of every op.
struct OpArena {
  uint8_t* base = nullptr;
  size_t cap = 0, used = 0;
  int dev = -1;
  std::vector<void*> overflow;  // blocks allocated after the arena filled up in this op
  void rewind() {
    for (void* q : overflow) cudaFree(q);
    overflow.clear();
    used = 0;
  }
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    int d = 0;
    cudaGetDevice(&d);
    if (d != dev) {  // the calling thread switched devices: start over
      if (base) cudaFree(base);
      base = nullptr; cap = used = 0; dev = d;
    }
    if (used + bytes > cap) {
      if (used == 0) {  // nothing handed out yet: regrow the arena itself
        if (base) cudaFree(base);
        base = nullptr;
        const size_t want = bytes * 2 > ((size_t)8 << 20) ? bytes * 2 : ((size_t)8 << 20);
        if (cudaMalloc((void**)&base, want) != cudaSuccess) { cap = 0; return nullptr; }
        cap = want;
      } else {
        void* q = nullptr;
        if (cudaMalloc(&q, bytes) != cudaSuccess) return nullptr;
        overflow.push_back(q);
        return q;
      }
    }
    void* r = base + used;
    used += bytes;
    return r;
  }
};
static thread_local OpArena g_arena;
struct DevBuf {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) {
    p = g_arena.take(bytes ? bytes : 16);
    return p ? cudaSuccess : cudaErrorMemoryAllocation;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
struct OpScope {  // rewinds the arena when an op-level entry point begins
  OpScope() { g_arena.rewind(); }
};
#define OPBUF(buf, bytes)                                                                     \
  DevBuf buf;                                                                                 \
  { cudaError_t e_ = buf.alloc(bytes); if (e_ != cudaSuccess) return fail(LNB_ENOMEM, "cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(e_)); }
#define H2D(buf, host, bytes) CU(cudaMemcpy(buf.p, host, bytes, cudaMemcpyHostToDevice))
#define D2H(host, buf, bytes) CU(cudaMemcpy(host, buf.p, bytes, cudaMemcpyDeviceToHost))
static int op_finish() {
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return 0;
}
static int grid_for(int64_t n) { return (int)std::min<int64_t>(148 * 8, (n + 255) / 256); }

extern "C" int lnb_op_linear_bf16(const uint16_t* x, const uint16_t* w, uint16_t* out, int S, int K, int N, int acc_mode) {
  OpScope scope_;
  if (!x || !w || !out) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0 || K <= 0 || N <= 0) return fail(LNB_EINVAL, "non-positive shape");
  if (acc_mode != LNB_ACC_STRICT && acc_mode != LNB_ACC_FAST) return fail(LNB_EINVAL, "bad acc_mode %d", acc_mode);
  const size_t xb = (size_t)S * K * 2, wb = (size_t)N * K * 2, ob = (size_t)S * N * 2;
  OPBUF(dx, xb); OPBUF(dw, wb); OPBUF(dout, ob);
  H2D(dx, x, xb); H2D(dw, w, wb);
  const bool tileable = (N % 16 == 0) && (K % 8 == 0) && (CfgF1::smem_bytes(K) <= kMaxSmem);
  if (acc_mode == LNB_ACC_FAST && tc_eligible(S, N, K)) {
    // prompt-sized inputs: tcgen05 tensor-core tiles (hardware accumulation order)
    const int Mpad = (S + TC_BM - 1) / TC_BM * TC_BM;
    OPBUF(dwp, wb); OPBUF(dx8, (size_t)Mpad * K * 2);
    retile_kernel<<<grid_for((int64_t)N * K / 8), 256>>>(dw.as<uint16_t>(), K, 0, 0, N, K, dwp.as<uint16_t>(), 0, 1);
    pack_x8_kernel<<<grid_for((int64_t)Mpad * K / 8), 256>>>(dx.as<uint16_t>(), K, S, Mpad, K, dx8.as<uint16_t>());
    GemmTcParams g{};
    g.X8 = dx8.as<uint16_t>(); g.W = dwp.as<uint16_t>(); g.M = S; g.N = N; g.K = K; g.out_bf16 = dout.as<uint16_t>(); g.ldo = N;
    Launcher L{nullptr, false, nullptr};
    int rc = launch_gemm_tc<TC_EPI_BF16>(L, g);
    if (rc) return rc;
    rc = op_finish();
    if (rc) return rc;
    D2H(out, dout, ob);
    return 0;
  }
  if (!tileable) {
    // shapes outside the panel layout (e.g. the reference's 2x3 . 4x3^T test): reference order, one thread per output
    linear_naive_kernel<<<(int)(((int64_t)S * N + 255) / 256), 256>>>(dx.as<uint16_t>(), dw.as<uint16_t>(), dout.as<uint16_t>(), S, K, N);
  } else {
    OPBUF(dwp, wb);
    retile_kernel<<<grid_for((int64_t)N * K / 8), 256>>>(dw.as<uint16_t>(), K, 0, 0, N, K, dwp.as<uint16_t>(), 0, 1);
    GemvParams p{};
    p.W = dwp.as<uint16_t>(); p.N = N; p.K = K; p.x = dx.as<uint16_t>(); p.ldx = K; p.out_bf16 = dout.as<uint16_t>(); p.ldo = N;
    Launcher L{nullptr, false, nullptr};
    int rc = launch_gemv<PRO_PLAIN, EPI_BF16>(L, acc_mode, p, S);
    if (rc) return rc;
    rc = op_finish();
    if (rc) return rc;
    D2H(out, dout, ob);
    return 0;
  }
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, ob);
  return 0;
}

extern "C" int lnb_op_matmul_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int B, int M, int K, int N) {
  OpScope scope_;
  if (!a || !b || !out) return fail(LNB_EINVAL, "NULL argument");
  if (B <= 0 || M <= 0 || K <= 0 || N <= 0) return fail(LNB_EINVAL, "non-positive shape");
  const size_t ab = (size_t)B * M * K * 2, bb = (size_t)B * K * N * 2, ob = (size_t)B * M * N * 2;
  OPBUF(da, ab); OPBUF(db, bb); OPBUF(dout, ob);
  H2D(da, a, ab); H2D(db, b, bb);
  matmul_naive_kernel<<<(int)(((int64_t)B * M * N + 255) / 256), 256>>>(da.as<uint16_t>(), db.as<uint16_t>(), dout.as<uint16_t>(), B, M, K, N);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, ob);
  return 0;
}

extern "C" int lnb_op_rmsnorm_bf16(const uint16_t* x, const uint16_t* w, uint16_t* out, int S, int D, float eps, int acc_mode) {
  OpScope scope_;
  if (!x || !w || !out) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0 || D <= 0) return fail(LNB_EINVAL, "non-positive shape");
  const size_t xb = (size_t)S * D * 2;
  OPBUF(dx, xb); OPBUF(dw, (size_t)D * 2); OPBUF(dout, xb);
  H2D(dx, x, xb); H2D(dw, w, (size_t)D * 2);
  const bool strict = (acc_mode == LNB_ACC_STRICT);
  const float* rs = nullptr;
  OPBUF(drs, (size_t)S * 4);
  if (strict && D % 2 == 0 && (size_t)D * 4 <= 48 * 1024) {  // the model path's own scale kernel
    const Launcher L{nullptr, false, nullptr};
    int rc0 = launch_rms_scale(L, dx.as<uint16_t>(), D, drs.as<float>(), S, D, eps);
    if (rc0) return rc0;
    rs = drs.as<float>();
  }
  rmsnorm_kernel<<<S, 256>>>(dx.as<uint16_t>(), dw.as<uint16_t>(), dout.as<uint16_t>(), D, eps, strict ? 1 : 0, rs);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, xb);
  return 0;
}

extern "C" int lnb_op_rms_scale_f32(const uint16_t* x, float* r, int S, int D, float eps, int algo) {
  OpScope scope_;
  if (!x || !r) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0 || D <= 0 || (D & 1)) return fail(LNB_EINVAL, "bad shape");
  if (algo < RMS_AUTO || algo > RMS_ENGINE) return fail(LNB_EINVAL, "bad algo %d", algo);
  if (algo == RMS_ENGINE) {
    int ech = 0, ent = 0;
    if (!eng_scan_shape(D, &ech, &ent) || (size_t)D * 4 + 4096 > 200 * 1024) return fail(LNB_EINVAL, "row length %d does not fit the engine's scan", D);
    const size_t xb2 = (size_t)S * D * 2;
    OPBUF(dx2, xb2); OPBUF(dr2, (size_t)S * 4);
    H2D(dx2, x, xb2);
    const size_t smem = (size_t)D * 4 + 4096;
    switch (ech) {
      case 16: CU(cudaFuncSetAttribute(eng_rms_scale_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
               eng_rms_scale_kernel<16><<<S, ENG_NCONS, smem>>>(dx2.as<uint16_t>(), D, dr2.as<float>(), D, eps); break;
      case 8: CU(cudaFuncSetAttribute(eng_rms_scale_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              eng_rms_scale_kernel<8><<<S, ENG_NCONS, smem>>>(dx2.as<uint16_t>(), D, dr2.as<float>(), D, eps); break;
      case 4: CU(cudaFuncSetAttribute(eng_rms_scale_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              eng_rms_scale_kernel<4><<<S, ENG_NCONS, smem>>>(dx2.as<uint16_t>(), D, dr2.as<float>(), D, eps); break;
      default: CU(cudaFuncSetAttribute(eng_rms_scale_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
               eng_rms_scale_kernel<2><<<S, ENG_NCONS, smem>>>(dx2.as<uint16_t>(), D, dr2.as<float>(), D, eps); break;
    }
    int rc2 = op_finish();
    if (rc2) return rc2;
    D2H(r, dr2, (size_t)S * 4);
    return 0;
  }
  int ch, nt;
  const bool fits = seq_scan_shape(D, &ch, &nt);
  if ((algo == RMS_SCAN || algo == RMS_SEG) && !fits) return fail(LNB_EINVAL, "row length %d does not fit the scan kernels", D);
  if ((size_t)D * 4 > 48 * 1024 && (algo == RMS_CHAIN || !fits)) return fail(LNB_EINVAL, "row length %d too long", D);
  const size_t xb = (size_t)S * D * 2;
  OPBUF(dx, xb); OPBUF(dr, (size_t)S * 4);
  H2D(dx, x, xb);
  const Launcher L{nullptr, false, nullptr};
  int rc = launch_rms_scale(L, dx.as<uint16_t>(), D, dr.as<float>(), S, D, eps, algo);
  if (rc) return rc;
  rc = op_finish();
  if (rc) return rc;
  D2H(r, dr, (size_t)S * 4);
  return 0;
}

extern "C" int lnb_op_rope_bf16(const uint16_t* x, const float* cis, uint16_t* out, int S, int H, int hd, int start_pos) {
  OpScope scope_;
  if (!x || !cis || !out) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0 || H <= 0 || hd <= 0 || (hd & 1) || start_pos < 0) return fail(LNB_EINVAL, "bad shape");
  const size_t xb = (size_t)S * H * hd * 2, cb = (size_t)(start_pos + S) * (hd / 2) * 2 * 4;
  OPBUF(dx, xb); OPBUF(dc, cb); OPBUF(dout, xb);
  H2D(dx, x, xb); H2D(dc, cis, cb);
  rope_kernel<<<grid_for((int64_t)S * H * hd / 2), 256>>>(dx.as<uint16_t>(), dc.as<float>(), dout.as<uint16_t>(), S, H, hd, start_pos);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, xb);
  return 0;
}

extern "C" int lnb_op_attention_bf16(const uint16_t* q, const uint16_t* cache_k, const uint16_t* cache_v, uint16_t* out,
                                     int S, int T, int n_heads, int n_kv, int hd, int causal_mask, int acc_mode) {
  OpScope scope_;
  if (!q || !cache_k || !cache_v || !out) return fail(LNB_EINVAL, "NULL argument");
  if (S <= 0 || T < S || n_heads <= 0 || n_kv <= 0 || n_heads % n_kv || hd <= 0 || hd % 8) return fail(LNB_EINVAL, "bad shape");
  if (causal_mask && T != S) return fail(LNB_EINVAL, "causal mask needs T == S (reference mask is [S,S])");
  const size_t smem = (size_t)T * 12 + (size_t)hd * 4;
  if (smem > 200 * 1024) return fail(LNB_EINVAL, "T %d too long", T);
  const size_t qb = (size_t)S * n_heads * hd * 2, cb = (size_t)T * n_kv * hd * 2;
  OPBUF(dq, qb); OPBUF(dk, cb); OPBUF(dv, cb); OPBUF(dout, qb);
  H2D(dq, q, qb); H2D(dk, cache_k, cb); H2D(dv, cache_v, cb);
  CU(cudaFuncSetAttribute(sdpa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  float f = (float)sqrt((double)hd);
  uint32_t u;
  memcpy(&u, &f, 4);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  if (acc_mode == LNB_ACC_FAST && S > 1 && causal_mask && hd == SP_HD && (n_heads * hd) % 128 == 0) {
    // the prefill path's tiled two-pass kernel (writes the X8 operand of the Wo GEMM; unpacked here)
    const int Mpad = (S + 127) / 128 * 128;
    OPBUF(d8, (size_t)Mpad * n_heads * hd * 2); OPBUF(dpos, 16);
    CU(cudaMemset(dpos.p, 0, 16));
    const double* exp_tab = (sdpa_tc_enabled() && S >= 32) ? exp_table_device() : nullptr;
    if (exp_tab && sdpa_tc_pipelined()) {
      CU(cudaFuncSetAttribute(sdpa_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SX_SMEM));
      sdpa_tc2_kernel<<<dim3(n_heads, (S + 127) / 128), 256, SX_SMEM>>>(dq.as<uint16_t>(), n_heads * hd, dk.as<uint16_t>(), dv.as<uint16_t>(),
                                                                        n_kv * hd, n_heads / n_kv, d8.as<uint16_t>(), n_heads * hd,
                                                                        dpos.as<int32_t>(), S, f, exp_tab);
    } else if (exp_tab) {
      CU(cudaFuncSetAttribute(sdpa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
      sdpa_tc_kernel<<<dim3(n_heads, (S + 127) / 128), 256, ST_SMEM>>>(dq.as<uint16_t>(), n_heads * hd, dk.as<uint16_t>(), dv.as<uint16_t>(),
                                                                       n_kv * hd, n_heads / n_kv, d8.as<uint16_t>(), n_heads * hd,
                                                                       dpos.as<int32_t>(), S, f, exp_tab);
    } else {
      CU(cudaFuncSetAttribute(sdpa_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
      sdpa_prefill_kernel<<<dim3(n_heads, (S + SP_QB - 1) / SP_QB), 256, SP_SMEM>>>(dq.as<uint16_t>(), n_heads * hd, dk.as<uint16_t>(),
                                                                                    dv.as<uint16_t>(), n_kv * hd, n_heads / n_kv,
                                                                                    d8.as<uint16_t>(), n_heads * hd, dpos.as<int32_t>(), S, f);
    }
    unpack_x8_kernel<<<grid_for((int64_t)S * n_heads * hd / 8), 256>>>(d8.as<uint16_t>(), S, n_heads * hd, dout.as<uint16_t>(), n_heads * hd);
    int rc2 = op_finish();
    if (rc2) return rc2;
    D2H(out, dout, qb);
    return 0;
  }
  sdpa_kernel<<<dim3(n_heads, S), 128, smem>>>(dq.as<uint16_t>(), n_heads * hd, dk.as<uint16_t>(), dv.as<uint16_t>(), n_kv * hd,
                                                n_heads / n_kv, hd, dout.as<uint16_t>(), n_heads * hd, nullptr, T - S, S,
                                                causal_mask ? 1 : 0, acc_mode == LNB_ACC_STRICT ? 1 : 0, f, 0);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, qb);
  return 0;
}

static int silu_table_device(uint16_t** out) {
  static uint16_t* tab = nullptr;  // per process, device 0 of the calling thread's current device
  static int tab_dev = -1;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (!tab || tab_dev != dev) {
    std::vector<uint16_t> h;
    build_silu_table(h);
    CU(cudaMalloc((void**)&tab, 65536 * 2));
    CU(cudaMemcpy(tab, h.data(), 65536 * 2, cudaMemcpyHostToDevice));
    tab_dev = dev;
  }
  *out = tab;
  return 0;
}

extern "C" int lnb_op_silu_bf16(const uint16_t* x, uint16_t* out, int64_t n) {
  OpScope scope_;
  if (!x || !out || n < 0) return fail(LNB_EINVAL, "bad argument");
  if (n == 0) return 0;
  uint16_t* tab;
  int rc = silu_table_device(&tab);
  if (rc) return rc;
  OPBUF(dx, (size_t)n * 2); OPBUF(dout, (size_t)n * 2);
  H2D(dx, x, (size_t)n * 2);
  silu_kernel<<<grid_for(n), 256>>>(dx.as<uint16_t>(), tab, dout.as<uint16_t>(), n);
  rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, (size_t)n * 2);
  return 0;
}

static int binary_op(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n, bool is_add) {
  OpScope scope_;
  if (!a || !b || !out || n < 0) return fail(LNB_EINVAL, "bad argument");
  if (n == 0) return 0;
  OPBUF(da, (size_t)n * 2); OPBUF(db, (size_t)n * 2); OPBUF(dout, (size_t)n * 2);
  H2D(da, a, (size_t)n * 2); H2D(db, b, (size_t)n * 2);
  if (is_add) add_kernel<<<grid_for(n), 256>>>(da.as<uint16_t>(), db.as<uint16_t>(), dout.as<uint16_t>(), n);
  else mul_kernel<<<grid_for(n), 256>>>(da.as<uint16_t>(), db.as<uint16_t>(), dout.as<uint16_t>(), n);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, (size_t)n * 2);
  return 0;
}
extern "C" int lnb_op_add_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n) { return binary_op(a, b, out, n, true); }
extern "C" int lnb_op_mul_bf16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n) { return binary_op(a, b, out, n, false); }

extern "C" int lnb_op_softmax_f32(const float* x, float* out, int rows, int cols) {
  OpScope scope_;
  if (!x || !out || rows <= 0 || cols <= 0) return fail(LNB_EINVAL, "bad argument");
  const size_t b = (size_t)rows * cols * 4;
  OPBUF(dx, b); OPBUF(dout, b);
  H2D(dx, x, b);
  softmax_f32_kernel<<<rows, 256>>>(dx.as<float>(), dout.as<float>(), cols);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, b);
  return 0;
}

extern "C" int lnb_op_argmax_f32(const float* x, int rows, int cols, int32_t* out) {
  OpScope scope_;
  if (!x || !out || rows <= 0 || cols <= 0) return fail(LNB_EINVAL, "bad argument");
  const size_t b = (size_t)rows * cols * 4;
  OPBUF(dx, b); OPBUF(dout, (size_t)rows * 4);
  H2D(dx, x, b);
  argmax_f32_kernel<<<rows, 256>>>(dx.as<float>(), cols, dout.as<int32_t>());
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, (size_t)rows * 4);
  return 0;
}

extern "C" int lnb_op_get_rows_bf16(const uint16_t* emb, const int32_t* tokens, uint16_t* out, int S, int vocab, int dim) {
  OpScope scope_;
  if (!emb || !tokens || !out || S <= 0 || vocab <= 0 || dim <= 0 || dim % 8) return fail(LNB_EINVAL, "bad argument");
  for (int i = 0; i < S; i++)
    if (tokens[i] < 0 || tokens[i] >= vocab) return fail(LNB_EINVAL, "token id %d out of range", tokens[i]);
  OPBUF(de, (size_t)vocab * dim * 2); OPBUF(dt, (size_t)S * 4); OPBUF(dout, (size_t)S * dim * 2);
  H2D(de, emb, (size_t)vocab * dim * 2); H2D(dt, tokens, (size_t)S * 4);
  gather_rows_kernel<<<S, 256>>>(de.as<uint16_t>(), dt.as<int32_t>(), nullptr, dout.as<uint16_t>(), dim);
  int rc = op_finish();
  if (rc) return rc;
  D2H(out, dout, (size_t)S * dim * 2);
  return 0;
}
