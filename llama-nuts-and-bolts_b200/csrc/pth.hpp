// pth.hpp -- PyTorch checkpoint (.pth = zip + pickle) reader and writer, host-only C++.
//
// What it replaces: src/torch/torchmodelreader.go:39-145 (zip lookup, persistent_load,
// storage -> mmap slice), src/torch/types.go:23-56 (rebuild_tensor_v2, TorchStorage.Load),
// src/pickle/* (the unpickler) and src/common/memorymapper_unix.go:21-45 (read-only mmap of the file).
// The tensors stay where the kernel page cache has them: PthFile hands out pointers into the mapping and
// lnb_model_load_pth feeds them straight to the (sharding) upload path -- no intermediate host copy.
//
// The writer emits the subset of pickle protocol 2 the reference's unpickler dispatches
// (src/pickle/pickledispatch.go:52-77), one storage per tensor at offset 0 (its rebuild_tensor_v2
// ignores the storage offset, src/torch/types.go:23-36), so the unmodified Go loader can read the
// synthetic checkpoint; tests also round-trip both directions through torch.load / torch.save.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace lnb {

enum PthDType { PTH_BF16 = 0, PTH_F16 = 1, PTH_F32 = 2, PTH_F64 = 3, PTH_I8 = 4, PTH_U8 = 5, PTH_I16 = 6, PTH_I32 = 7, PTH_I64 = 8, PTH_BOOL = 9 };
int pth_item_size(int dtype);
const char* pth_dtype_name(int dtype);

struct PthTensor {
  std::string name;
  int dtype = PTH_BF16;
  std::vector<int64_t> shape, stride;
  std::string storage_key;      // file name stem under <archive>/data/
  int64_t storage_offset = 0;   // in elements
  int64_t storage_numel = 0;    // elements in the whole storage (persistent id field 4)
  int64_t file_offset = 0;      // byte offset of element 0 of THIS tensor inside the .pth file
  int64_t nbytes = 0;           // numel * item size
  bool contiguous = true;
};

class PthFile {
 public:
  PthFile() = default;
  ~PthFile();
  PthFile(const PthFile&) = delete;
  PthFile& operator=(const PthFile&) = delete;
  // returns false and fills err on any malformed input (never throws across this interface)
  bool open(const std::string& path, std::string& err);
  const std::vector<PthTensor>& tensors() const { return tensors_; }
  const uint8_t* data(const PthTensor& t) const { return base_ + t.file_offset; }
  int find(const std::string& name) const;
  int64_t file_size() const { return size_; }

 private:
  struct ZipEntry {
    std::string name;
    uint16_t method = 0;
    uint64_t comp_size = 0, size = 0, local_offset = 0, data_offset = 0;
  };
  bool read_zip_index(std::string& err);
  bool unpickle(const ZipEntry& pkl, std::string& err);
  const ZipEntry* entry(const std::string& name) const;
  const uint8_t* base_ = nullptr;
  int64_t size_ = 0;
  int fd_ = -1;
  std::vector<ZipEntry> entries_;
  std::vector<PthTensor> tensors_;
};

class PthWriter {
 public:
  PthWriter() = default;
  ~PthWriter();
  bool open(const std::string& path, std::string& err);
  // one storage per tensor; data is copied to the file immediately
  bool add(const std::string& name, int dtype, const void* data, const std::vector<int64_t>& shape, std::string& err);
  bool finish(std::string& err);  // data.pkl, version, byteorder, central directory

 private:
  struct Rec {
    std::string name;
    uint32_t crc = 0;
    uint64_t size = 0, local_offset = 0;
  };
  struct Meta {
    std::string name;
    int dtype;
    std::vector<int64_t> shape;
  };
  bool put(const void* p, size_t n);
  bool add_entry(const std::string& name, const void* data, uint64_t size, bool align64, std::string& err);
  FILE* f_ = nullptr;
  uint64_t pos_ = 0;
  std::vector<Rec> recs_;
  std::vector<Meta> metas_;
  bool failed_ = false;
};

uint32_t crc32_ieee(uint32_t crc, const void* data, size_t n);

// params.json -> the fields of model.ModelArgs (src/model/modelargs.go:13-64); absent keys keep the
// reference's defaults (NewModelArgs :29-44).  Returns false with err on malformed JSON.
struct ParamsJson {
  int dim = 4096, n_layers = 32, n_heads = 32, n_kv_heads = -1, vocab_size = -1, multiple_of = 256;
  double ffn_dim_multiplier = -1;
  float norm_eps = 1e-5f;
  double rope_theta = 500000;
  bool use_scaled_rope = false;
};
bool parse_params_json(const std::string& text, ParamsJson& out, std::string& err);

}  // namespace lnb
