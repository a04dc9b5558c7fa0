// tokenizer.hpp -- tiktoken vocabulary + Llama-3 BPE tokenizer, host-only C++.
//
// What it replaces: tiktoken.Load (src/tiktoken/tiktokenreader.go:12-85), model.NewVocabulary
// (src/model/vocabulary.go:23-50, incl. the split regexp at :36), InferenceEngine.TokenizeString /
// bytePairMerge / Tokenize (src/inference/tokenize.go:27-193) and the byte concatenation under
// TokenBatchToString (:239-258).  The emoji alias annotation of the console UI (src/inference/emoji.go) is
// out of scope: it needs github.com/enescakir/emoji's alias table and x/text's rune names, neither of which
// is in the reference tree.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace lnb {

struct PromptPart {   // inference.PromptPart (tokenize.go:21-25)
  std::string header, content;
};

class Vocab {
 public:
  // tiktoken.Load + NewVocabulary: `<base64 token> <rank>` lines, then the 256 special tokens appended
  bool load(const std::string& path, std::string& err);
  // same from memory (tests build synthetic vocabularies); ranks must be 0..n-1 without gaps
  bool load_from_text(const std::string& text, std::string& err);

  int size() const { return (int)id_to_token.size(); }
  int32_t id_of(const std::string& token) const;   // -1 when absent
  // InferenceEngine.TokenizeString (tokenize.go:175-193)
  void tokenize_string(const std::string& text, std::vector<int32_t>& out) const;
  // InferenceEngine.Tokenize (tokenize.go:27-95): chat template around the parts + the trailing assistant header
  bool tokenize_prompt(const std::vector<PromptPart>& parts, std::vector<int32_t>& out, std::string& err) const;
  // bytes of the tokens up to the first PadId; false when an id is out of range
  bool detokenize(const int32_t* ids, int n, std::string& out) const;

  // split regexp of vocabulary.go:36 as a hand-written matcher: byte offset one past the piece starting at pos
  static size_t next_piece(const std::string& text, size_t pos);

  std::unordered_map<std::string, int32_t> token_to_id;
  std::vector<std::string> id_to_token;
  int32_t bos_id = -1, eos_id = -1, pad_id = -1, unknown_id = -1;
  std::vector<int32_t> stop_ids;
  int n_mergeable = 0;

 private:
  void byte_pair_merge(const std::string& piece, std::vector<int32_t>& out) const;   // tokenize.go:109-173
};

// A stand-in tokenizer.model with exactly n_mergeable ranks: the 256 single bytes followed by unique 3-byte filler
// strings.  No 2-byte token exists, so the byte-pair merge never fires: text tokenizes byte by byte (a whole 3-byte
// piece may coincide with a filler and become that one id; it detokenizes to the same bytes).  Lets a synthetic model directory satisfy checkModelArgs (VocabSize == vocabulary length,
// src/model/loader.go:98-120) so that the unmodified reference -- or lnb_generate prompt=... -- runs end to end.
bool write_synthetic_vocab(const std::string& path, int n_mergeable, std::string& err);

}  // namespace lnb
