// tokenizer.cpp -- see tokenizer.hpp.  Host-only (no CUDA); part of liblnb.so.
#include "tokenizer.hpp"

#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstring>
#include <queue>

#include "unicode_tables.hpp"

namespace lnb {

// ---------------------------------------------------------------------------------------------
// character classes of the split regexp (Go RE2 semantics)
static bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    if (cp < r[mid][0]) hi = mid - 1;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
static inline bool is_L(uint32_t cp) { return in_ranges(kUnicodeL, kUnicodeL_count, cp); }
static inline bool is_N(uint32_t cp) { return in_ranges(kUnicodeN, kUnicodeN_count, cp); }
// RE2's \s is ASCII only: [\t\n\f\r ]
static inline bool is_space(uint32_t cp) { return cp == ' ' || cp == '\t' || cp == '\n' || cp == '\f' || cp == '\r'; }
static inline bool is_newline(uint32_t cp) { return cp == '\r' || cp == '\n'; }
static inline bool is_other(uint32_t cp) { return !is_space(cp) && !is_L(cp) && !is_N(cp); }   // [^\s\p{L}\p{N}]

// one code point at byte offset p (p < s.size()); malformed bytes decode as U+FFFD of width 1, as in Go
static uint32_t decode(const std::string& s, size_t p, size_t* len) {
  const unsigned char* b = (const unsigned char*)s.data() + p;
  const size_t n = s.size() - p;
  const unsigned char c = b[0];
  *len = 1;
  if (c < 0x80) return c;
  int need;
  uint32_t cp, min;
  if ((c & 0xe0) == 0xc0) { need = 1; cp = c & 0x1f; min = 0x80; }
  else if ((c & 0xf0) == 0xe0) { need = 2; cp = c & 0x0f; min = 0x800; }
  else if ((c & 0xf8) == 0xf0) { need = 3; cp = c & 0x07; min = 0x10000; }
  else return 0xfffd;
  if ((size_t)need >= n) return 0xfffd;
  for (int i = 1; i <= need; i++) {
    if ((b[i] & 0xc0) != 0x80) return 0xfffd;
    cp = (cp << 6) | (b[i] & 0x3f);
  }
  if (cp < min || cp > 0x10ffff || (cp >= 0xd800 && cp <= 0xdfff)) return 0xfffd;
  *len = (size_t)need + 1;
  return cp;
}

// (?i:x) for the letters of the contraction alternative; Go folds with unicode.SimpleFold orbits, so 's' also
// matches U+017F (LATIN SMALL LETTER LONG S)
static bool fold_eq(uint32_t cp, char lower) {
  if (cp == (uint32_t)lower || cp == (uint32_t)(lower - 32)) return true;
  return lower == 's' && cp == 0x17f;
}

// (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+
// leftmost-first: the alternatives are tried in order at `pos`, each with greedy quantifiers and backtracking.
size_t Vocab::next_piece(const std::string& text, size_t pos) {
  const size_t n = text.size();
  size_t l0;
  const uint32_t c0 = decode(text, pos, &l0);
  // 1. contractions
  if (c0 == '\'' && pos + 1 < n) {
    size_t l1, l2 = 0;
    const uint32_t c1 = decode(text, pos + 1, &l1);
    const uint32_t c2 = (pos + 1 + l1 < n) ? decode(text, pos + 1 + l1, &l2) : 0;
    if (fold_eq(c1, 's') || fold_eq(c1, 't')) return pos + 1 + l1;
    if (fold_eq(c1, 'r') && l2 && fold_eq(c2, 'e')) return pos + 1 + l1 + l2;
    if (fold_eq(c1, 'v') && l2 && fold_eq(c2, 'e')) return pos + 1 + l1 + l2;
    if (fold_eq(c1, 'm')) return pos + 1 + l1;
    if (fold_eq(c1, 'l') && l2 && fold_eq(c2, 'l')) return pos + 1 + l1 + l2;
    if (fold_eq(c1, 'd')) return pos + 1 + l1;
  }
  // 2. [^\r\n\p{L}\p{N}]?\p{L}+
  {
    size_t q = pos;
    if (!is_L(c0) && !is_N(c0) && !is_newline(c0)) q = pos + l0;   // the optional prefix character
    size_t e = q;
    while (e < n) {
      size_t l;
      if (!is_L(decode(text, e, &l))) break;
      e += l;
    }
    if (e > q) return e;   // (if the prefix was taken but no letter follows, dropping it cannot help: c0 is no letter)
  }
  // 3. \p{N}{1,3}
  if (is_N(c0)) {
    size_t e = pos + l0;
    for (int k = 1; k < 3 && e < n; k++) {
      size_t l;
      if (!is_N(decode(text, e, &l))) break;
      e += l;
    }
    return e;
  }
  // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
  {
    size_t q = (c0 == ' ') ? pos + 1 : pos;
    size_t e = q;
    while (e < n) {
      size_t l;
      if (!is_other(decode(text, e, &l))) break;
      e += l;
    }
    if (e > q) {
      while (e < n && (text[e] == '\r' || text[e] == '\n')) e++;
      return e;
    }
  }
  // 5. \s*[\r\n]+   6. \s+
  if (is_space(c0)) {
    size_t e = pos, last_nl = std::string::npos;
    while (e < n && is_space((unsigned char)text[e])) {
      if (text[e] == '\r' || text[e] == '\n') last_nl = e;
      e++;
    }
    return last_nl != std::string::npos ? last_nl + 1 : e;
  }
  return pos + l0;   // unreachable: every code point is a letter, a number, white space or "other"
}

// ---------------------------------------------------------------------------------------------
// vocabulary
static int b64val(unsigned char c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}
// base64.StdEncoding: padded, no line breaks
static bool b64decode(const char* s, size_t n, std::string* out) {
  out->clear();
  if (n % 4) return false;
  for (size_t i = 0; i < n; i += 4) {
    int v[4];
    int pad = 0;
    for (int k = 0; k < 4; k++) {
      if (s[i + k] == '=') {
        if (i + 4 != n || k < 2) return false;
        v[k] = 0;
        pad++;
      } else {
        if (pad) return false;
        v[k] = b64val((unsigned char)s[i + k]);
        if (v[k] < 0) return false;
      }
    }
    const uint32_t w = ((uint32_t)v[0] << 18) | ((uint32_t)v[1] << 12) | ((uint32_t)v[2] << 6) | (uint32_t)v[3];
    out->push_back((char)(w >> 16));
    if (pad < 2) out->push_back((char)((w >> 8) & 0xff));
    if (pad < 1) out->push_back((char)(w & 0xff));
  }
  return true;
}

bool Vocab::load(const std::string& path, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "open " + path + ": " + strerror(errno); return false; }
  std::string text;
  char buf[65536];
  size_t k;
  while ((k = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, k);
  fclose(f);
  return load_from_text(text, err);
}

bool Vocab::load_from_text(const std::string& text, std::string& err) {
  token_to_id.clear();
  id_to_token.clear();
  std::vector<std::pair<std::string, int>> entries;
  size_t p = 0;
  int line_no = 0;
  while (p < text.size()) {
    size_t e = text.find('\n', p);
    if (e == std::string::npos) e = text.size();
    size_t le = e;
    if (le > p && text[le - 1] == '\r') le--;   // bufio.ScanLines drops a trailing \r
    line_no++;
    const size_t sp = text.find(' ', p);
    if (sp == std::string::npos || sp >= le) { err = "tokenizer model line " + std::to_string(line_no) + ": expected \"<base64> <rank>\""; return false; }
    std::string tok;
    if (!b64decode(text.data() + p, sp - p, &tok)) { err = "tokenizer model line " + std::to_string(line_no) + ": illegal base64 data"; return false; }
    size_t re = text.find(' ', sp + 1);
    if (re == std::string::npos || re > le) re = le;
    if (re == sp + 1) { err = "tokenizer model line " + std::to_string(line_no) + ": missing rank"; return false; }
    long rank = 0;
    for (size_t i = sp + 1; i < re; i++) {
      if (text[i] < '0' || text[i] > '9' || rank > INT_MAX / 10 - 1) { err = "tokenizer model line " + std::to_string(line_no) + ": bad rank"; return false; }
      rank = rank * 10 + (text[i] - '0');
    }
    entries.emplace_back(std::move(tok), (int)rank);
    p = e + 1;
  }
  if (entries.empty()) { err = "tokenizer model is empty"; return false; }
  n_mergeable = (int)entries.size();
  // the special tokens follow the mergeable ranks (tiktokenreader.go:46-72)
  static const char* named[11] = {"<|begin_of_text|>", "<|end_of_text|>", "<|reserved_special_token_0|>", "<|reserved_special_token_1|>",
                                  "<|finetune_right_pad_id|>", "<|step_id|>", "<|start_header_id|>", "<|end_header_id|>",
                                  "<|eom_id|>", "<|eot_id|>", "<|python_tag|>"};
  std::vector<std::string> special(named, named + 11);
  for (int i = 0; i < 256 - 11; i++) special.push_back("<|reserved_special_token_" + std::to_string(2 + i) + "|>");
  id_to_token.assign((size_t)n_mergeable + special.size(), std::string());
  std::vector<char> seen(id_to_token.size(), 0);
  for (auto& e : entries) {
    if (e.second < 0 || e.second >= n_mergeable) { err = "tokenizer model: rank " + std::to_string(e.second) + " outside 0.." + std::to_string(n_mergeable - 1); return false; }
    if (seen[e.second]) { err = "tokenizer model: duplicate rank " + std::to_string(e.second); return false; }
    seen[e.second] = 1;
    token_to_id[e.first] = e.second;
    id_to_token[e.second] = e.first;
  }
  for (size_t i = 0; i < special.size(); i++) {
    token_to_id[special[i]] = n_mergeable + (int)i;
    id_to_token[n_mergeable + i] = special[i];
  }
  bos_id = id_of("<|begin_of_text|>");
  eos_id = id_of("<|end_of_text|>");
  pad_id = -1;
  unknown_id = -1;
  stop_ids = {id_of("<|eom_id|>"), id_of("<|eot_id|>")};
  return true;
}

int32_t Vocab::id_of(const std::string& token) const {
  auto it = token_to_id.find(token);
  return it == token_to_id.end() ? -1 : it->second;
}

// ---------------------------------------------------------------------------------------------
// byte pair merge: repeatedly fuse the adjacent pair of parts whose concatenation has the lowest rank
// (leftmost on ties) until no adjacent pair is a token.  Works on BYTES like tiktoken and the reference.
// Same result as the reference's quadratic scan (tokenize.go:109-173), found with a heap in O(n log n) so that a
// pathological piece (a 100 KB "word") cannot stall the caller.
void Vocab::byte_pair_merge(const std::string& piece, std::vector<int32_t>& out) const {
  const int n = (int)piece.size();
  auto rank_of = [&](int a, int b) -> int {   // rank of bytes [a, b) or INT_MAX
    auto it = token_to_id.find(piece.substr((size_t)a, (size_t)(b - a)));
    return it == token_to_id.end() ? INT_MAX : it->second;
  };
  // parts as a linked list over their start offsets: part i = [i, end[i]); nxt/prv = neighbouring starts (-1 = none)
  std::vector<int> end(n), nxt(n), prv(n);
  std::vector<char> alive(n, 1);
  for (int i = 0; i < n; i++) { end[i] = i + 1; nxt[i] = i + 1 < n ? i + 1 : -1; prv[i] = i - 1; }
  struct Cand {
    int rank, start, span;                     // span = end of the right part at push time: stale entries are skipped
    bool operator>(const Cand& o) const { return rank != o.rank ? rank > o.rank : start > o.start; }
  };
  std::priority_queue<Cand, std::vector<Cand>, std::greater<Cand>> heap;
  auto push = [&](int i) {
    if (i < 0 || nxt[i] < 0) return;
    const int r = rank_of(i, end[nxt[i]]);
    if (r != INT_MAX) heap.push(Cand{r, i, end[nxt[i]]});
  };
  for (int i = 0; i + 1 < n; i++) push(i);
  while (!heap.empty()) {
    const Cand c = heap.top();
    heap.pop();
    const int i = c.start;
    if (!alive[i] || nxt[i] < 0 || end[nxt[i]] != c.span || end[i] >= c.span) continue;   // one side changed since the push
    const int j = nxt[i];
    if (rank_of(i, end[j]) != c.rank) continue;
    alive[j] = 0;                                // fuse part j into part i
    end[i] = end[j];
    nxt[i] = nxt[j];
    if (nxt[i] >= 0) prv[nxt[i]] = i;
    push(prv[i]);
    push(i);
  }
  for (int i = 0; i >= 0 && i < n; i = nxt[i]) {
    auto it = token_to_id.find(piece.substr((size_t)i, (size_t)(end[i] - i)));
    out.push_back(it == token_to_id.end() ? 0 : it->second);   // Go's map zero value (tokenize.go:169)
  }
}

void Vocab::tokenize_string(const std::string& text, std::vector<int32_t>& out) const {
  size_t p = 0;
  while (p < text.size()) {
    const size_t e = next_piece(text, p);
    const std::string piece = text.substr(p, e - p);
    auto it = token_to_id.find(piece);
    if (it != token_to_id.end()) out.push_back(it->second);
    else byte_pair_merge(piece, out);
    p = e;
  }
}

bool Vocab::tokenize_prompt(const std::vector<PromptPart>& parts, std::vector<int32_t>& out, std::string& err) const {
  const int32_t b_txt = id_of("<|begin_of_text|>"), b_hdr = id_of("<|start_header_id|>"), e_hdr = id_of("<|end_header_id|>"),
                e_turn = id_of("<|eot_id|>");
  if (b_txt < 0 || b_hdr < 0 || e_hdr < 0 || e_turn < 0) { err = "vocabulary has no special tokens"; return false; }
  out.push_back(b_txt);
  for (size_t i = 0; i <= parts.size(); i++) {
    const bool last_assistant = (i == parts.size());           // appended by the reference (tokenize.go:36-40)
    const std::string header = last_assistant ? "assistant" : parts[i].header;
    const std::string content = last_assistant ? "" : parts[i].content;
    if (!last_assistant && content.empty()) continue;
    out.push_back(b_hdr);
    tokenize_string(header, out);
    out.push_back(e_hdr);
    tokenize_string("\n\n", out);
    tokenize_string(content, out);
    if (!last_assistant) out.push_back(e_turn);
  }
  return true;
}

bool Vocab::detokenize(const int32_t* ids, int n, std::string& out) const {
  out.clear();
  for (int i = 0; i < n; i++) {
    if (ids[i] == pad_id) break;                               // tokenize.go:246-248
    if (ids[i] < 0 || ids[i] >= (int)id_to_token.size()) return false;
    out += id_to_token[(size_t)ids[i]];
  }
  return true;
}

static void b64encode(const unsigned char* p, size_t n, std::string* out) {
  static const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  for (size_t i = 0; i < n; i += 3) {
    const uint32_t b0 = p[i], b1 = i + 1 < n ? p[i + 1] : 0, b2 = i + 2 < n ? p[i + 2] : 0;
    const uint32_t w = (b0 << 16) | (b1 << 8) | b2;
    out->push_back(T[w >> 18]);
    out->push_back(T[(w >> 12) & 63]);
    out->push_back(i + 1 < n ? T[(w >> 6) & 63] : '=');
    out->push_back(i + 2 < n ? T[w & 63] : '=');
  }
}

bool write_synthetic_vocab(const std::string& path, int n_mergeable, std::string& err) {
  if (n_mergeable < 256 || n_mergeable > 256 + (1 << 23)) { err = "synthetic vocabulary needs 256 <= n_mergeable <= 256 + 2^23"; return false; }
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "create " + path + ": " + strerror(errno); return false; }
  std::string line;
  for (int r = 0; r < n_mergeable; r++) {
    unsigned char tok[3];
    size_t len = 1;
    if (r < 256) tok[0] = (unsigned char)r;
    else {
      const uint32_t j = (uint32_t)(r - 256);
      tok[0] = (unsigned char)(j & 0xff);
      tok[1] = (unsigned char)((j >> 8) & 0xff);
      tok[2] = (unsigned char)(0x80 | ((j >> 16) & 0x7f));
      len = 3;
    }
    line.clear();
    b64encode(tok, len, &line);
    line += " " + std::to_string(r) + "\n";
    if (fwrite(line.data(), 1, line.size(), f) != line.size()) { fclose(f); err = "write failed"; return false; }
  }
  if (fclose(f) != 0) { err = "close failed"; return false; }
  return true;
}

}  // namespace lnb
