// Mutation fuzzer for the .pth reader (llama-nuts-and-bolts_b200/csrc/pth.cpp), built with
// -fsanitize=address,undefined by tests/test_pth_cpu.py.  Every mutant must either parse (then every tensor's
// byte range must be readable) or fail with a message -- never crash.  Test infrastructure.
//   usage: pth_fuzz <valid.pth> <scratch path> <iterations> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "pth.hpp"

static uint64_t s_rng;
static uint32_t rnd() {
  s_rng = s_rng * 6364136223846793005ULL + 1442695040888963407ULL;
  return (uint32_t)(s_rng >> 33);
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint8_t> good;
  uint8_t buf[65536];
  size_t k;
  while ((k = fread(buf, 1, sizeof(buf), f)) > 0) good.insert(good.end(), buf, buf + k);
  fclose(f);
  const int iters = atoi(argv[3]);
  s_rng = strtoull(argv[4], nullptr, 10) * 2654435761u + 1;
  // offsets of interesting regions: the pickle and the tail (central directory, end records)
  size_t pkl = 0;
  for (size_t i = 0; i + 8 < good.size(); i++)
    if (!memcmp(&good[i], "data.pkl", 8)) { pkl = i; break; }
  long n_ok = 0, n_err = 0;
  volatile uint8_t sink = 0;
  for (int it = 0; it < iters; it++) {
    std::vector<uint8_t> b = good;
    switch (it % 4) {
      case 0: b.resize(rnd() % good.size()); break;
      case 1: for (uint32_t j = 0, n = 1 + rnd() % 6; j < n; j++) b[rnd() % b.size()] = (uint8_t)rnd(); break;
      case 2: for (int j = 0; j < 3; j++) b[pkl + rnd() % std::min<size_t>(600, b.size() - pkl)] = (uint8_t)rnd(); break;
      default: for (int j = 0; j < 3; j++) b[b.size() - 1 - rnd() % std::min<size_t>(700, b.size())] = (uint8_t)rnd(); break;
    }
    f = fopen(argv[2], "wb");
    if (!f) return 2;
    if (!b.empty()) fwrite(b.data(), 1, b.size(), f);
    fclose(f);
    lnb::PthFile pf;
    std::string err;
    if (pf.open(argv[2], err)) {
      for (const auto& t : pf.tensors())
        if (t.nbytes > 0) { sink ^= pf.data(t)[0]; sink ^= pf.data(t)[t.nbytes - 1]; }   // strided views too: nbytes = touched span
      n_ok++;
    } else {
      if (err.empty()) { fprintf(stderr, "failure without a message at iteration %d\n", it); return 1; }
      n_err++;
    }
  }
  printf("parsed %ld rejected %ld\n", n_ok, n_err);
  return 0;
}
