// pth.cpp -- see pth.hpp.  Host-only (no CUDA); part of liblnb.so.
#include "pth.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdarg>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

namespace lnb {

// ---------------------------------------------------------------------------------------------
// small helpers
static const int kItemSize[10] = {2, 2, 4, 8, 1, 1, 2, 4, 8, 1};
static const char* kDTypeName[10] = {"bfloat16", "float16", "float32", "float64", "int8", "uint8", "int16", "int32", "int64", "bool"};
static const char* kStorageClass[10] = {"BFloat16Storage", "HalfStorage", "FloatStorage", "DoubleStorage", "CharStorage",
                                        "ByteStorage",     "ShortStorage", "IntStorage",  "LongStorage",   "BoolStorage"};
int pth_item_size(int dtype) { return (dtype >= 0 && dtype < 10) ? kItemSize[dtype] : 0; }
const char* pth_dtype_name(int dtype) { return (dtype >= 0 && dtype < 10) ? kDTypeName[dtype] : "?"; }

static std::string fmt(const char* f, ...) __attribute__((format(printf, 1, 2)));
static std::string fmt(const char* f, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// CRC-32 (IEEE 802.3, reflected 0xEDB88320), slicing-by-8
static uint32_t g_crc_tab[8][256];
static bool g_crc_init = false;
static void crc_init() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
  g_crc_init = true;
}
uint32_t crc32_ieee(uint32_t crc, const void* data, size_t n) {
  if (!g_crc_init) crc_init();
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8); n--; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    const uint32_t lo = (uint32_t)v ^ c, hi = (uint32_t)(v >> 32);
    c = g_crc_tab[7][lo & 0xff] ^ g_crc_tab[6][(lo >> 8) & 0xff] ^ g_crc_tab[5][(lo >> 16) & 0xff] ^ g_crc_tab[4][lo >> 24] ^
        g_crc_tab[3][hi & 0xff] ^ g_crc_tab[2][(hi >> 8) & 0xff] ^ g_crc_tab[1][(hi >> 16) & 0xff] ^ g_crc_tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

// ---------------------------------------------------------------------------------------------
// reader: mmap + zip central directory
PthFile::~PthFile() {
  if (base_) munmap((void*)base_, (size_t)size_);
  if (fd_ >= 0) close(fd_);
}

const PthFile::ZipEntry* PthFile::entry(const std::string& name) const {
  for (const auto& e : entries_)
    if (e.name == name) return &e;
  return nullptr;
}

int PthFile::find(const std::string& name) const {
  for (size_t i = 0; i < tensors_.size(); i++)
    if (tensors_[i].name == name) return (int)i;
  return -1;
}

bool PthFile::read_zip_index(std::string& err) {
  const uint64_t n = (uint64_t)size_;
  if (n < 22) { err = "not a zip archive (too short)"; return false; }
  // end of central directory record: last 22 + up to 65535 comment bytes
  int64_t eocd = -1;
  const int64_t lo = (int64_t)n - 22 - 65535;
  for (int64_t p = (int64_t)n - 22; p >= 0 && p >= lo; p--)
    if (rd32(base_ + p) == 0x06054b50u) { eocd = p; break; }
  if (eocd < 0) { err = "not a zip archive (no end-of-central-directory record)"; return false; }
  uint64_t n_entries = rd16(base_ + eocd + 10), cd_size = rd32(base_ + eocd + 12), cd_off = rd32(base_ + eocd + 16);
  if (n_entries == 0xffff || cd_size == 0xffffffffu || cd_off == 0xffffffffu) {
    if (eocd < 20 || rd32(base_ + eocd - 20) != 0x07064b50u) { err = "zip64 locator missing"; return false; }
    const uint64_t z = rd64(base_ + eocd - 20 + 8);
    if (n < 56 || z > n - 56 || rd32(base_ + z) != 0x06064b50u) { err = "zip64 end-of-central-directory record missing"; return false; }
    n_entries = rd64(base_ + z + 32);
    cd_size = rd64(base_ + z + 40);
    cd_off = rd64(base_ + z + 48);
  }
  if (cd_off > n || cd_size > n - cd_off) { err = "central directory out of bounds"; return false; }
  uint64_t p = cd_off;
  const uint64_t cd_end = cd_off + cd_size;
  entries_.clear();
  entries_.reserve((size_t)std::min<uint64_t>(n_entries, 1u << 20));
  for (uint64_t i = 0; i < n_entries; i++) {
    if (p + 46 > cd_end || rd32(base_ + p) != 0x02014b50u) { err = fmt("central directory entry %llu malformed", (unsigned long long)i); return false; }
    ZipEntry e;
    e.method = rd16(base_ + p + 10);
    e.comp_size = rd32(base_ + p + 20);
    e.size = rd32(base_ + p + 24);
    const uint32_t nlen = rd16(base_ + p + 28), elen = rd16(base_ + p + 30), clen = rd16(base_ + p + 32);
    e.local_offset = rd32(base_ + p + 42);
    if (p + 46 + nlen + elen + clen > cd_end) { err = "central directory entry overruns"; return false; }
    e.name.assign((const char*)base_ + p + 46, nlen);
    // zip64 extended information: only the fields whose 32-bit value is saturated, in this order
    const uint8_t* x = base_ + p + 46 + nlen;
    uint32_t xo = 0;
    while (xo + 4 <= elen) {
      const uint16_t id = rd16(x + xo), sz = rd16(x + xo + 2);
      if (xo + 4 + sz > elen) break;
      if (id == 0x0001) {
        uint32_t q = xo + 4;
        const uint32_t qe = xo + 4 + sz;
        if (e.size == 0xffffffffu && q + 8 <= qe) { e.size = rd64(x + q); q += 8; }
        if (e.comp_size == 0xffffffffu && q + 8 <= qe) { e.comp_size = rd64(x + q); q += 8; }
        if (e.local_offset == 0xffffffffu && q + 8 <= qe) { e.local_offset = rd64(x + q); q += 8; }
      }
      xo += 4 + sz;
    }
    if (n < 30 || e.local_offset > n - 30 || rd32(base_ + e.local_offset) != 0x04034b50u) { err = fmt("local header of \"%s\" malformed", e.name.c_str()); return false; }
    e.data_offset = e.local_offset + 30 + rd16(base_ + e.local_offset + 26) + rd16(base_ + e.local_offset + 28);
    if (e.data_offset > n || e.comp_size > n - e.data_offset) { err = fmt("data of \"%s\" out of bounds", e.name.c_str()); return false; }
    if (e.method == 0 && e.size != e.comp_size) { err = fmt("stored entry \"%s\": size fields disagree", e.name.c_str()); return false; }
    entries_.push_back(std::move(e));
    p += 46 + nlen + elen + clen;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// reader: unpickler (protocol <= 5 subset that torch.save emits for dict[str, Tensor])
namespace {
struct Val;
using VP = std::shared_ptr<Val>;
struct Val {
  enum Kind { NONE, BOOL, INT, FLOAT, STR, TUPLE, LIST, DICT, GLOBAL, STORAGE, TENSOR, MARK, OPAQUE } k = NONE;
  int64_t i = 0;        // BOOL / INT; STORAGE: dtype
  double f = 0;
  std::string s;        // STR; GLOBAL: "module.name"; STORAGE: key
  int64_t numel = 0;    // STORAGE
  std::vector<VP> items;  // TUPLE / LIST; DICT: key, value, key, value ...
  PthTensor t;          // TENSOR
};
static VP mk(Val::Kind k) { auto v = std::make_shared<Val>(); v->k = k; return v; }

struct Cursor {
  const uint8_t* p;
  size_t n, o = 0;
  bool need(size_t k) const { return o + k <= n && o + k >= o; }
};

static int storage_dtype(const std::string& g) {
  if (g.compare(0, 6, "torch.") != 0) return -1;
  for (int d = 0; d < 10; d++)
    if (g.compare(6, std::string::npos, kStorageClass[d]) == 0) return d;
  return -1;
}
static bool as_int(const VP& v, int64_t* out) {
  if (!v || (v->k != Val::INT && v->k != Val::BOOL)) return false;
  *out = v->i;
  return true;
}
static bool int_list(const VP& v, std::vector<int64_t>* out) {
  if (!v || (v->k != Val::TUPLE && v->k != Val::LIST)) return false;
  out->clear();
  for (auto& it : v->items) {
    int64_t x;
    if (!as_int(it, &x)) return false;
    out->push_back(x);
  }
  return true;
}
}  // namespace

bool PthFile::unpickle(const ZipEntry& pkl, std::string& err) {
  Cursor c{base_ + pkl.data_offset, (size_t)pkl.size};
  std::vector<VP> stack;
  std::map<uint32_t, VP> memo;
#define NEED(k) do { if (!c.need(k)) { err = "pickle truncated"; return false; } } while (0)
#define POP(var) VP var; do { if (stack.empty() || stack.back()->k == Val::MARK) { err = "pickle stack underflow"; return false; } var = stack.back(); stack.pop_back(); } while (0)
  auto pop_mark = [&](std::vector<VP>* items) -> bool {
    size_t m = stack.size();
    while (m > 0 && stack[m - 1]->k != Val::MARK) m--;
    if (m == 0) return false;
    items->assign(stack.begin() + m, stack.end());
    stack.resize(m - 1);
    return true;
  };
  auto read_line = [&](std::string* out) -> bool {
    size_t e = c.o;
    while (e < c.n && c.p[e] != '\n') e++;
    if (e >= c.n) return false;
    out->assign((const char*)c.p + c.o, e - c.o);
    c.o = e + 1;
    return true;
  };
  auto push_str = [&](size_t len) -> bool {
    if (!c.need(len)) return false;
    VP v = mk(Val::STR);
    v->s.assign((const char*)c.p + c.o, len);
    c.o += len;
    stack.push_back(v);
    return true;
  };
  auto reduce = [&](const VP& fn, const VP& args, VP* out) -> bool {
    if (fn->k != Val::GLOBAL) { err = "REDUCE on a non-callable"; return false; }
    const std::string& g = fn->s;
    if (g == "collections.OrderedDict") {
      VP d = mk(Val::DICT);
      if (args && !args->items.empty()) {   // OrderedDict([(k, v), ...])
        const VP& lst = args->items[0];
        if (lst->k != Val::LIST && lst->k != Val::TUPLE) { err = "OrderedDict(arg): unsupported argument"; return false; }
        for (auto& kv : lst->items) {
          if ((kv->k != Val::TUPLE && kv->k != Val::LIST) || kv->items.size() != 2) { err = "OrderedDict(arg): items must be pairs"; return false; }
          d->items.push_back(kv->items[0]);
          d->items.push_back(kv->items[1]);
        }
      }
      *out = d;
      return true;
    }
    if (g == "torch._utils._rebuild_tensor_v2" || g == "torch._utils._rebuild_tensor") {
      // (storage, storage_offset, size, stride[, requires_grad, backward_hooks, metadata])   src/torch/types.go:23-36
      if (!args || args->items.size() < 4 || args->items[0]->k != Val::STORAGE) { err = "_rebuild_tensor_v2: bad arguments"; return false; }
      VP t = mk(Val::TENSOR);
      const VP& st = args->items[0];
      t->t.dtype = (int)st->i;
      t->t.storage_key = st->s;
      t->t.storage_numel = st->numel;
      if (!as_int(args->items[1], &t->t.storage_offset) || !int_list(args->items[2], &t->t.shape) || !int_list(args->items[3], &t->t.stride) ||
          t->t.shape.size() != t->t.stride.size()) {
        err = "_rebuild_tensor_v2: bad offset / size / stride";
        return false;
      }
      *out = t;
      return true;
    }
    if (g == "torch._utils._rebuild_parameter") {
      if (!args || args->items.empty() || args->items[0]->k != Val::TENSOR) { err = "_rebuild_parameter: bad arguments"; return false; }
      *out = args->items[0];
      return true;
    }
    err = fmt("unknown class \"%s\" not found", g.c_str());   // same text as findClassTorch (torchmodelreader.go:103-113)
    return false;
  };

  for (;;) {
    NEED(1);
    const uint8_t op = c.p[c.o++];
    switch (op) {
      case 0x80: NEED(1); if (c.p[c.o] > 5) { err = fmt("unsupported pickle protocol: %d", c.p[c.o]); return false; } c.o++; break;
      case 0x95: NEED(8); c.o += 8; break;                                   // FRAME
      case '}': stack.push_back(mk(Val::DICT)); break;
      case ']': stack.push_back(mk(Val::LIST)); break;
      case ')': stack.push_back(mk(Val::TUPLE)); break;
      case '(': stack.push_back(mk(Val::MARK)); break;
      case 'N': stack.push_back(mk(Val::NONE)); break;
      case 0x88: case 0x89: { VP v = mk(Val::BOOL); v->i = (op == 0x88); stack.push_back(v); break; }
      case 'q': NEED(1); if (stack.empty()) { err = "BINPUT on empty stack"; return false; } memo[c.p[c.o++]] = stack.back(); break;
      case 'r': NEED(4); if (stack.empty()) { err = "LONG_BINPUT on empty stack"; return false; } memo[rd32(c.p + c.o)] = stack.back(); c.o += 4; break;
      case 0x94: if (stack.empty()) { err = "MEMOIZE on empty stack"; return false; } { const uint32_t idx = (uint32_t)memo.size(); memo[idx] = stack.back(); } break;
      case 'h': case 'j': {
        uint32_t idx;
        if (op == 'h') { NEED(1); idx = c.p[c.o++]; } else { NEED(4); idx = rd32(c.p + c.o); c.o += 4; }
        auto it = memo.find(idx);
        if (it == memo.end()) { err = fmt("memo value not found at index %u", idx); return false; }
        stack.push_back(it->second);
        break;
      }
      case 'X': case 'T': case 'B': { NEED(4); const uint32_t len = rd32(c.p + c.o); c.o += 4; if (!push_str(len)) { err = "pickle truncated"; return false; } break; }
      case 0x8c: case 'U': case 'C': { NEED(1); const uint32_t len = c.p[c.o++]; if (!push_str(len)) { err = "pickle truncated"; return false; } break; }
      case 0x8d: { NEED(8); const uint64_t len = rd64(c.p + c.o); c.o += 8; if (len > c.n || !push_str((size_t)len)) { err = "pickle truncated"; return false; } break; }
      case 'c': {
        std::string mod, name;
        if (!read_line(&mod) || !read_line(&name)) { err = "pickle truncated"; return false; }
        VP v = mk(Val::GLOBAL);
        v->s = mod + "." + name;
        stack.push_back(v);
        break;
      }
      case 0x93: {
        POP(name); POP(mod);
        if (name->k != Val::STR || mod->k != Val::STR) { err = "STACK_GLOBAL needs two strings"; return false; }
        VP v = mk(Val::GLOBAL);
        v->s = mod->s + "." + name->s;
        stack.push_back(v);
        break;
      }
      case 'J': { NEED(4); VP v = mk(Val::INT); v->i = (int32_t)rd32(c.p + c.o); c.o += 4; stack.push_back(v); break; }
      case 'K': { NEED(1); VP v = mk(Val::INT); v->i = c.p[c.o++]; stack.push_back(v); break; }
      case 'M': { NEED(2); VP v = mk(Val::INT); v->i = rd16(c.p + c.o); c.o += 2; stack.push_back(v); break; }
      case 0x8a: {
        NEED(1); const uint32_t len = c.p[c.o++]; NEED(len);
        if (len > 8) { err = "LONG1 wider than 64 bits"; return false; }
        uint64_t u = 0;
        for (uint32_t b = 0; b < len; b++) u |= (uint64_t)c.p[c.o + b] << (8 * b);
        if (len && len < 8 && (c.p[c.o + len - 1] & 0x80)) u |= ~0ull << (8 * len);   // sign-extend
        c.o += len;
        VP v = mk(Val::INT); v->i = (int64_t)u; stack.push_back(v);
        break;
      }
      case 'G': {
        NEED(8);
        uint64_t u = 0;
        for (int b = 0; b < 8; b++) u = (u << 8) | c.p[c.o + b];
        c.o += 8;
        VP v = mk(Val::FLOAT); memcpy(&v->f, &u, 8); stack.push_back(v);
        break;
      }
      case 't': { VP v = mk(Val::TUPLE); if (!pop_mark(&v->items)) { err = "TUPLE without MARK"; return false; } stack.push_back(v); break; }
      case 0x85: case 0x86: case 0x87: {
        const size_t k = op - 0x84;
        if (stack.size() < k) { err = "pickle stack underflow"; return false; }
        VP v = mk(Val::TUPLE);
        v->items.assign(stack.end() - k, stack.end());
        for (auto& it : v->items) if (it->k == Val::MARK) { err = "pickle stack underflow"; return false; }
        stack.resize(stack.size() - k);
        stack.push_back(v);
        break;
      }
      case 'l': { VP v = mk(Val::LIST); if (!pop_mark(&v->items)) { err = "LIST without MARK"; return false; } stack.push_back(v); break; }
      case 'a': { POP(x); if (stack.empty() || stack.back()->k != Val::LIST) { err = "APPEND to a non-list"; return false; } stack.back()->items.push_back(x); break; }
      case 'e': {
        std::vector<VP> items;
        if (!pop_mark(&items) || stack.empty() || stack.back()->k != Val::LIST) { err = "APPENDS to a non-list"; return false; }
        for (auto& it : items) stack.back()->items.push_back(it);
        break;
      }
      case 's': {
        POP(v); POP(k);
        if (stack.empty() || stack.back()->k != Val::DICT) { err = "SETITEM on a non-dict"; return false; }
        stack.back()->items.push_back(k); stack.back()->items.push_back(v);
        break;
      }
      case 'u': {
        std::vector<VP> items;
        if (!pop_mark(&items) || stack.empty() || stack.back()->k != Val::DICT || (items.size() & 1)) { err = "SETITEMS on a non-dict"; return false; }
        for (auto& it : items) stack.back()->items.push_back(it);
        break;
      }
      case 'R': {
        POP(args); POP(fn);
        if (args->k != Val::TUPLE) { err = "REDUCE arguments must be a tuple"; return false; }
        VP out;
        if (!reduce(fn, args, &out)) return false;
        stack.push_back(out);
        break;
      }
      case 0x81: {  // NEWOBJ: cls.__new__(cls, *args)
        POP(args); POP(cls);
        if (cls->k == Val::GLOBAL && cls->s == "collections.OrderedDict") stack.push_back(mk(Val::DICT));
        else { err = fmt("unknown class \"%s\" not found", cls->k == Val::GLOBAL ? cls->s.c_str() : "?"); return false; }
        break;
      }
      case 'b': { POP(state); if (stack.empty()) { err = "BUILD on empty stack"; return false; } break; }  // e.g. OrderedDict._metadata: ignored
      case 'Q': {
        // persistent id ('storage', storage class, key, location, numel)   torchmodelreader.go:115-145
        POP(pid);
        if (pid->k != Val::TUPLE || pid->items.size() < 5 || pid->items[0]->k != Val::STR || pid->items[0]->s != "storage") {
          err = "pid[0] must have value \"storage\"";
          return false;
        }
        if (pid->items[1]->k != Val::GLOBAL) { err = "pid[1] must be a storage class"; return false; }
        const int dt = storage_dtype(pid->items[1]->s);
        if (dt < 0) { err = fmt("unknown class \"%s\" not found", pid->items[1]->s.c_str()); return false; }
        VP v = mk(Val::STORAGE);
        v->i = dt;
        if (pid->items[2]->k != Val::STR) { err = "pid[2] must be the storage key"; return false; }
        v->s = pid->items[2]->s;
        if (!as_int(pid->items[4], &v->numel) || v->numel < 0) { err = "pid[4] must be the element count"; return false; }
        stack.push_back(v);
        break;
      }
      case '.': {
        POP(top);
        if (top->k != Val::DICT) { err = "the pickled object is not a dict of tensors"; return false; }
        const std::string base = pkl.name.substr(0, pkl.name.size() - 4);   // "<archive>/data" (torchmodelreader.go:94)
        tensors_.clear();
        for (size_t i = 0; i + 1 < top->items.size(); i += 2) {
          const VP& k = top->items[i];
          const VP& v = top->items[i + 1];
          if (k->k != Val::STR || v->k != Val::TENSOR) continue;   // non-tensor entries are not ours to interpret
          PthTensor t = v->t;
          t.name = k->s;
          int found = find(t.name);                                // later keys replace earlier ones (PickleDict.Set)
          const ZipEntry* e = entry(base + "/" + t.storage_key);
          if (!e) { err = fmt("file \"%s/%s\" not found in Torch model file", base.c_str(), t.storage_key.c_str()); return false; }
          if (e->method != 0) { err = fmt("storage \"%s\" is compressed; only stored entries can be mapped", e->name.c_str()); return false; }
          const int64_t isz = pth_item_size(t.dtype);
          int64_t numel = 1;
          for (int64_t d : t.shape) {
            if (d < 0 || (d && numel > INT64_MAX / (d ? d : 1))) { err = fmt("tensor \"%s\": bad shape", t.name.c_str()); return false; }
            numel *= d;
          }
          int64_t expect = 1;
          t.contiguous = true;
          for (int d = (int)t.shape.size() - 1; d >= 0; d--) {
            if (t.shape[d] != 1 && t.stride[d] != expect) t.contiguous = false;
            expect *= t.shape[d];
          }
          // elements of the storage the tensor can touch: [offset, offset + sum((shape[d]-1)*stride[d]) + 1), every
          // product checked against the storage size (a view such as zeros(1).expand(65536, 4096) has numel >> span)
          int64_t span = numel ? 1 : 0;
          bool bad = t.storage_offset < 0 || t.storage_offset > t.storage_numel || isz <= 0 ||
                     (uint64_t)t.storage_numel > (uint64_t)e->size / (uint64_t)isz;          // by division: no wrap-around
          for (size_t d = 0; d < t.shape.size() && !bad && numel; d++) {
            const int64_t st = t.stride[d], ext = t.shape[d] - 1;
            if (st < 0) bad = true;                                                          // torch never pickles negative strides
            else if (ext > 0 && st > 0) {
              if (st > t.storage_numel / ext) bad = true;
              else span += st * ext;
            }
            if (span > t.storage_numel) bad = true;
          }
          if (!bad && span > t.storage_numel - t.storage_offset) bad = true;
          if (bad) {
            err = fmt("tensor \"%s\" exceeds its storage", t.name.c_str());
            return false;
          }
          t.file_offset = (int64_t)e->data_offset + t.storage_offset * isz;
          // bytes reachable from file_offset: the dense size for contiguous tensors, the touched span of the storage
          // for strided views (callers that want a dense copy must gather through shape/stride themselves)
          t.nbytes = (t.contiguous ? numel : span) * isz;
          if (found >= 0) tensors_.erase(tensors_.begin() + found);
          tensors_.push_back(std::move(t));
        }
        return true;
      }
      default: err = fmt("unsupported Pickle op code: 0x%X '%c'", op, (op >= 32 && op < 127) ? op : '?'); return false;
    }
  }
#undef NEED
#undef POP
}

bool PthFile::open(const std::string& path, std::string& err) {
  fd_ = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
  if (fd_ < 0) { err = fmt("open %s: %s", path.c_str(), strerror(errno)); return false; }
  struct stat st;
  if (fstat(fd_, &st) != 0) { err = fmt("error getting file info: %s", strerror(errno)); return false; }
  size_ = (int64_t)st.st_size;
  if (size_ <= 0) { err = "not a zip archive (empty file)"; return false; }
  void* p = mmap(nullptr, (size_t)size_, PROT_READ, MAP_SHARED, fd_, 0);   // memorymapper_unix.go:34
  if (p == MAP_FAILED) { err = fmt("error mapping file into memory: %s", strerror(errno)); return false; }
  base_ = (const uint8_t*)p;
  if (!read_zip_index(err)) return false;
  const ZipEntry* pkl = nullptr;
  int n_pkl = 0;
  for (const auto& e : entries_)
    if (e.name.size() >= 4 && e.name.compare(e.name.size() - 4, 4, ".pkl") == 0) { pkl = &e; n_pkl++; }
  if (n_pkl != 1) { err = fmt("no .pkl file found in Torch model file \"%s\"", path.c_str()); return false; }   // torchmodelreader.go:48-50
  if (pkl->method != 0) { err = "data.pkl is compressed; torch.save stores it uncompressed"; return false; }
  return unpickle(*pkl, err);
}

// ---------------------------------------------------------------------------------------------
// writer
PthWriter::~PthWriter() {
  if (f_) fclose(f_);
}
bool PthWriter::open(const std::string& path, std::string& err) {
  f_ = fopen(path.c_str(), "wb");
  if (!f_) { err = fmt("create %s: %s", path.c_str(), strerror(errno)); return false; }
  setvbuf(f_, nullptr, _IOFBF, 1 << 20);
  return true;
}
bool PthWriter::put(const void* p, size_t n) {
  if (n && fwrite(p, 1, n, f_) != n) { failed_ = true; return false; }
  pos_ += n;
  return true;
}
namespace {
struct Buf {
  std::string b;
  void u8(uint8_t v) { b.push_back((char)v); }
  void u16(uint16_t v) { u8(v & 0xff); u8(v >> 8); }
  void u32(uint32_t v) { u16(v & 0xffff); u16(v >> 16); }
  void u64(uint64_t v) { u32((uint32_t)v); u32((uint32_t)(v >> 32)); }
  void str(const std::string& s) { b += s; }
};
}  // namespace

bool PthWriter::add_entry(const std::string& name, const void* data, uint64_t size, bool align64, std::string& err) {
  if (size >= 0xffffffffull) { err = fmt("entry \"%s\" is 4 GiB or larger", name.c_str()); return false; }
  Rec r;
  r.name = name;
  r.size = size;
  r.local_offset = pos_;
  r.crc = crc32_ieee(0, data, (size_t)size);
  // torch pads the local extra field ("FB" record) so that the payload starts on a 64-byte boundary
  uint32_t pad = 0;
  if (align64) pad = (uint32_t)((64 - (pos_ + 30 + name.size() + 4) % 64) % 64);
  Buf h;
  h.u32(0x04034b50u); h.u16(20); h.u16(0); h.u16(0); h.u16(0); h.u16(0x0021);
  h.u32(r.crc); h.u32((uint32_t)size); h.u32((uint32_t)size);
  h.u16((uint16_t)name.size()); h.u16(align64 ? (uint16_t)(4 + pad) : 0);
  h.str(name);
  if (align64) { h.u16(0x4246); h.u16((uint16_t)pad); h.b.append(pad, 'Z'); }
  if (!put(h.b.data(), h.b.size()) || !put(data, (size_t)size)) { err = fmt("write failed: %s", strerror(errno)); return false; }
  recs_.push_back(r);
  return true;
}

bool PthWriter::add(const std::string& name, int dtype, const void* data, const std::vector<int64_t>& shape, std::string& err) {
  if (!f_ || failed_) { err = "writer is not open"; return false; }
  const int isz = pth_item_size(dtype);
  if (!isz) { err = "bad dtype"; return false; }
  int64_t numel = 1;
  for (int64_t d : shape) {
    if (d < 0) { err = "negative dimension"; return false; }
    numel *= d;
  }
  if (numel > 0x7fffffffll) { err = fmt("tensor \"%s\": element count needs a LONG1 pickle op the reference cannot read", name.c_str()); return false; }
  for (int64_t d : shape)
    if (d > 0x7fffffffll) { err = "dimension too large"; return false; }
  const std::string key = std::to_string(metas_.size());
  if (!add_entry("archive/data/" + key, data, (uint64_t)numel * isz, true, err)) return false;
  metas_.push_back(Meta{name, dtype, shape});
  return true;
}

static void pk_int(Buf& p, int64_t v) {
  if (v >= 0 && v < 256) { p.u8('K'); p.u8((uint8_t)v); }
  else if (v >= 0 && v < 65536) { p.u8('M'); p.u16((uint16_t)v); }
  else { p.u8('J'); p.u32((uint32_t)(int32_t)v); }
}
static void pk_str(Buf& p, const std::string& s) { p.u8('X'); p.u32((uint32_t)s.size()); p.str(s); }
static void pk_int_tuple(Buf& p, const std::vector<int64_t>& v) {
  // TUPLE3 is avoided on purpose: the reference's load_tuple3 duplicates the middle element
  // (src/pickle/pickledispatch.go:236-240); MARK ... TUPLE is read correctly by every unpickler.
  if (v.empty()) { p.u8(')'); return; }
  if (v.size() >= 3) p.u8('(');
  for (int64_t x : v) pk_int(p, x);
  p.u8(v.size() == 1 ? 0x85 : v.size() == 2 ? 0x86 : 't');
}

bool PthWriter::finish(std::string& err) {
  if (!f_ || failed_) { err = "writer is not open"; return false; }
  // data.pkl -- protocol 2, only opcodes of src/pickle/pickledispatch.go:52-77
  Buf p;
  p.u8(0x80); p.u8(2);
  p.u8('}'); p.u8('q'); p.u8(0);
  p.u8('(');
  int memo_dtype[10];
  for (int& m : memo_dtype) m = -1;
  int next_memo = 5;
  bool first = true;
  for (size_t i = 0; i < metas_.size(); i++) {
    const Meta& m = metas_[i];
    pk_str(p, m.name);
    if (first) { p.u8('c'); p.str("torch._utils\n_rebuild_tensor_v2\n"); p.u8('q'); p.u8(1); } else { p.u8('h'); p.u8(1); }
    p.u8('(');      // argument tuple of _rebuild_tensor_v2
    p.u8('(');      // persistent id
    if (first) { pk_str(p, "storage"); p.u8('q'); p.u8(2); } else { p.u8('h'); p.u8(2); }
    if (memo_dtype[m.dtype] < 0) {
      p.u8('c'); p.str(std::string("torch\n") + kStorageClass[m.dtype] + "\n");
      memo_dtype[m.dtype] = next_memo++;
      p.u8('q'); p.u8((uint8_t)memo_dtype[m.dtype]);
    } else { p.u8('h'); p.u8((uint8_t)memo_dtype[m.dtype]); }
    pk_str(p, std::to_string(i));
    if (first) { pk_str(p, "cpu"); p.u8('q'); p.u8(3); } else { p.u8('h'); p.u8(3); }
    int64_t numel = 1;
    for (int64_t d : m.shape) numel *= d;
    pk_int(p, numel);
    p.u8('t'); p.u8('Q');
    pk_int(p, 0);                         // storage offset
    pk_int_tuple(p, m.shape);
    std::vector<int64_t> stride(m.shape.size());
    int64_t s = 1;
    for (int d = (int)m.shape.size() - 1; d >= 0; d--) { stride[d] = s; s *= m.shape[d]; }
    pk_int_tuple(p, stride);
    p.u8(0x89);                           // requires_grad = False
    if (first) { p.u8('c'); p.str("collections\nOrderedDict\n"); p.u8('q'); p.u8(4); } else { p.u8('h'); p.u8(4); }
    p.u8(')'); p.u8('R');                 // backward_hooks = OrderedDict()
    p.u8('t'); p.u8('R');
    first = false;
  }
  p.u8('u'); p.u8('.');
  if (!add_entry("archive/data.pkl", p.b.data(), p.b.size(), false, err)) return false;
  if (!add_entry("archive/byteorder", "little", 6, false, err)) return false;
  if (!add_entry("archive/version", "3\n", 2, false, err)) return false;

  // central directory (+ zip64 records when offsets or counts do not fit)
  // LNB_PTH_FORCE_ZIP64=1: emit the zip64 records regardless of size (they are legal for any archive) so that
  // the >4 GiB code path of the 8B checkpoint can be exercised with a tiny file
  const char* fz = getenv("LNB_PTH_FORCE_ZIP64");
  const bool force64 = fz && *fz && *fz != '0';
  const uint64_t cd_off = pos_;
  for (const Rec& r : recs_) {
    const bool z = force64 || r.local_offset >= 0xffffffffull;
    Buf h;
    h.u32(0x02014b50u); h.u16(z ? 45 : 20); h.u16(z ? 45 : 20); h.u16(0); h.u16(0); h.u16(0); h.u16(0x0021);
    h.u32(r.crc); h.u32((uint32_t)r.size); h.u32((uint32_t)r.size);
    h.u16((uint16_t)r.name.size()); h.u16(z ? 12 : 0); h.u16(0); h.u16(0); h.u16(0); h.u32(0);
    h.u32(z ? 0xffffffffu : (uint32_t)r.local_offset);
    h.str(r.name);
    if (z) { h.u16(0x0001); h.u16(8); h.u64(r.local_offset); }
    if (!put(h.b.data(), h.b.size())) { err = fmt("write failed: %s", strerror(errno)); return false; }
  }
  const uint64_t cd_size = pos_ - cd_off, n = recs_.size();
  const bool z64 = force64 || n >= 0xffff || cd_off >= 0xffffffffull || cd_size >= 0xffffffffull;
  Buf t;
  if (z64) {
    const uint64_t z64_off = pos_;
    t.u32(0x06064b50u); t.u64(44); t.u16(45); t.u16(45); t.u32(0); t.u32(0); t.u64(n); t.u64(n); t.u64(cd_size); t.u64(cd_off);
    t.u32(0x07064b50u); t.u32(0); t.u64(z64_off); t.u32(1);
  }
  t.u32(0x06054b50u); t.u16(0); t.u16(0);
  t.u16(z64 ? 0xffff : (uint16_t)n); t.u16(z64 ? 0xffff : (uint16_t)n);
  t.u32(z64 ? 0xffffffffu : (uint32_t)cd_size); t.u32(z64 ? 0xffffffffu : (uint32_t)cd_off); t.u16(0);
  if (!put(t.b.data(), t.b.size())) { err = fmt("write failed: %s", strerror(errno)); return false; }
  const int rc = fclose(f_);
  f_ = nullptr;
  if (rc != 0) { err = fmt("close failed: %s", strerror(errno)); return false; }
  return true;
}

// ---------------------------------------------------------------------------------------------
// params.json (flat object of numbers / booleans / strings / null; nested values are skipped)
namespace {
struct J {
  const char* p;
  const char* e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool str(std::string* out) {
    if (p >= e || *p != '"') return false;
    p++;
    out->clear();
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) { p++; out->push_back(*p == 'n' ? '\n' : *p == 't' ? '\t' : *p); p++; }
      else out->push_back(*p++);
    }
    if (p >= e) return false;
    p++;
    return true;
  }
  bool skip_value() {   // arrays / objects / anything we do not map
    int depth = 0;
    while (p < e) {
      if (*p == '"') { std::string s; if (!str(&s)) return false; continue; }
      if (*p == '[' || *p == '{') depth++;
      else if (*p == ']' || *p == '}') { if (depth == 0) return true; depth--; if (depth == 0) { p++; return true; } }
      else if (*p == ',' && depth == 0) return true;
      p++;
    }
    return depth == 0;
  }
};
}  // namespace

bool parse_params_json(const std::string& text, ParamsJson& out, std::string& err) {
  J j{text.data(), text.data() + text.size()};
  j.ws();
  if (j.p >= j.e || *j.p != '{') { err = "params.json: expected an object"; return false; }
  j.p++;
  for (;;) {
    j.ws();
    if (j.p < j.e && *j.p == '}') return true;
    std::string key;
    if (!j.str(&key)) { err = "params.json: expected a key"; return false; }
    j.ws();
    if (j.p >= j.e || *j.p != ':') { err = "params.json: expected ':'"; return false; }
    j.p++;
    j.ws();
    const char* v0 = j.p;
    bool is_null = false, is_bool = false, bval = false, is_num = false;
    double num = 0;
    if (j.e - j.p >= 4 && !strncmp(j.p, "null", 4)) { is_null = true; j.p += 4; }
    else if (j.e - j.p >= 4 && !strncmp(j.p, "true", 4)) { is_bool = true; bval = true; j.p += 4; }
    else if (j.e - j.p >= 5 && !strncmp(j.p, "false", 5)) { is_bool = true; j.p += 5; }
    else if (j.p < j.e && (*j.p == '-' || (*j.p >= '0' && *j.p <= '9'))) {
      char* end = nullptr;
      std::string tmp(j.p, (size_t)std::min<ptrdiff_t>(j.e - j.p, 64));
      num = strtod(tmp.c_str(), &end);
      if (end == tmp.c_str()) { err = "params.json: bad number"; return false; }
      j.p += end - tmp.c_str();
      is_num = true;
    } else if (!j.skip_value()) { err = "params.json: malformed value"; return false; }
    (void)v0;
    if (!is_null) {   // encoding/json leaves the default in place for null
      if (is_num) {
        if (key == "dim") out.dim = (int)num;
        else if (key == "n_layers") out.n_layers = (int)num;
        else if (key == "n_heads") out.n_heads = (int)num;
        else if (key == "n_kv_heads") out.n_kv_heads = (int)num;
        else if (key == "vocab_size") out.vocab_size = (int)num;
        else if (key == "multiple_of") out.multiple_of = (int)num;
        else if (key == "ffn_dim_multiplier") out.ffn_dim_multiplier = num;
        else if (key == "norm_eps") out.norm_eps = (float)num;
        else if (key == "rope_theta") out.rope_theta = num;
      } else if (is_bool) {
        if (key == "use_scaled_rope") out.use_scaled_rope = bval;
      }
    }
    j.ws();
    if (j.p < j.e && *j.p == ',') { j.p++; continue; }
    if (j.p < j.e && *j.p == '}') return true;
    err = "params.json: expected ',' or '}'";
    return false;
  }
}

}  // namespace lnb
