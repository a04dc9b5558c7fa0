// lnb_generate -- the cmd/main.go of this repository for the synthetic checkpoint: loads (random-inits)
// Llama-3.1-8B on device 0 through the C++ host mirror, runs the reference generate loop on the fixed
// 8-token prompt and prints the generated token ids and the decode rate.
// Build: make -C host      Run: host/lnb_generate [seq_len=136] [strict|fast] [tiny | <modelDir>] [batch=N]
// batch=N: N prompts (prompt b = ids[1:] + 977 b) generated together, one pass over the weights per step.
// prompt=TEXT (with <modelDir>/tokenizer.model): TEXT goes through the chat template and the BPE tokenizer, the
// generated ids come back as text (cmd/main.go's flow without the console UI).
// <modelDir> holds params.json + consolidated.00.pth (model.LoadModel, src/model/loader.go:18-70);
// `lnb_generate --write-synthetic <modelDir> [tiny]` writes such a directory for the synthetic weights (host-only).
// `lnb_generate --tokenize <tokenizer.model> <text>` prints the chat-template token ids of <text> (host-only).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "lnb_host.hpp"

using namespace lnb_host;

static int write_synthetic_dir(const char* dir, bool tiny) {
  model::ModelArgs args = model::ModelArgs::Llama31_8B();
  const char* params = "{\"dim\": 4096, \"n_layers\": 32, \"n_heads\": 32, \"n_kv_heads\": 8, \"vocab_size\": 128256, "
                       "\"ffn_dim_multiplier\": 1.3, \"multiple_of\": 1024, \"norm_eps\": 1e-05, \"rope_theta\": 500000.0, "
                       "\"use_scaled_rope\": true}\n";
  if (tiny) {
    args.dim = 256; args.n_layers = 2; args.n_heads = 8; args.n_kv_heads = 2; args.head_dim = 32; args.ffn_dim = 512;
    args.vocab_size = 1024; args.max_seq_len = 64;
    params = "{\"dim\": 256, \"n_layers\": 2, \"n_heads\": 8, \"n_kv_heads\": 2, \"vocab_size\": 1024, \"ffn_dim_multiplier\": 0.75, "
             "\"multiple_of\": 256, \"norm_eps\": 1e-05, \"rope_theta\": 500000.0, \"use_scaled_rope\": true}\n";
  }
  const std::string d(dir);
  FILE* f = fopen((d + "/params.json").c_str(), "w");
  if (!f) { fprintf(stderr, "error: cannot write %s/params.json\n", dir); return 1; }
  fputs(params, f);
  fclose(f);
  check(lnb_pth_write_synthetic((d + "/consolidated.00.pth").c_str(), &args, tiny ? 7 : 0x4C4E42));
  check(lnb_vocab_write_synthetic((d + "/tokenizer.model").c_str(), args.vocab_size - 256));   // + 256 special tokens
  printf("wrote %s/params.json, %s/consolidated.00.pth and %s/tokenizer.model\n", dir, dir, dir);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 2 && !strcmp(argv[1], "--write-synthetic")) {
    try {
      return write_synthetic_dir(argv[2], argc > 3 && !strcmp(argv[3], "tiny"));
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  if (argc > 3 && !strcmp(argv[1], "--tokenize")) {   // host-only: chat template + BPE + round trip
    try {
      model::Tokenizer tok(argv[2]);
      const std::vector<int32_t> ids = tok.Tokenize({{"user", argv[3]}});
      printf("ids:");
      for (int32_t t : ids) printf(" %d", t);
      printf("\ntext: %s\n", tok.TokenBatchToString(ids).c_str());
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  if (argc > 3 && !strcmp(argv[1], "--detok-stream")) {   // host-only: streaming TokenToString over the given ids
    try {
      model::Tokenizer tok(argv[2]);
      std::string waiting, text;
      for (int i = 3; i < argc; i++) {
        const bool added = tok.TokenToString(atoi(argv[i]), waiting, text);
        printf("%d %s [%s]\n", atoi(argv[i]), added ? "waiting" : "text", text.c_str());
      }
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  const int seq_len = argc > 1 ? atoi(argv[1]) : 136;
  const int acc = (argc > 2 && !strcmp(argv[2], "fast")) ? LNB_ACC_FAST : LNB_ACC_STRICT;
  const bool tiny = argc > 3 && !strcmp(argv[3], "tiny");
  const bool has_dir = argc > 3 && !tiny && strncmp(argv[3], "batch=", 6) != 0;
  try {
    model::ModelArgs args = model::ModelArgs::Llama31_8B();
    std::vector<int32_t> prompt{128000, 9906, 11, 856, 836, 374, 220, 16};
    if (tiny) {
      args.dim = 256; args.n_layers = 2; args.n_heads = 8; args.n_kv_heads = 2; args.head_dim = 32; args.ffn_dim = 512;
      args.vocab_size = 1024; args.max_seq_len = 64;
      prompt = {1, 50, 999, 7, 300, 12, 64, 2};
    }
    const bool from_dir = has_dir;
    std::unique_ptr<model::LlamaTransformer> loaded;
    if (from_dir) {
      loaded = model::LoadModel(argv[3], 0, 2048);
      if (loaded->args.vocab_size < 128256) prompt = {1, 50, 999, 7, 300, 12, 64, 2};
    } else {
      loaded = std::make_unique<model::LlamaTransformer>(args, 0);
      loaded->InitSynthetic(tiny ? 7 : 0x4C4E42);
      loaded->Finalize();
    }
    model::LlamaTransformer& transformer = *loaded;
    model::Vocabulary vocab;
    if (tiny || transformer.args.vocab_size < 128256) vocab.StopTokenIds = {1000000000};
    int batch = 0;
    const char* prompt_text = nullptr;
    for (int i = 1; i < argc; i++) {
      if (!strncmp(argv[i], "batch=", 6)) batch = atoi(argv[i] + 6);
      if (!strncmp(argv[i], "prompt=", 7)) prompt_text = argv[i] + 7;
    }
    std::unique_ptr<model::Tokenizer> tokenizer;
    if (prompt_text) {
      if (!from_dir) throw Error("prompt= needs a model directory with tokenizer.model");
      tokenizer = std::make_unique<model::Tokenizer>(std::string(argv[3]) + "/tokenizer.model");
      if (tokenizer->Size() != transformer.args.vocab_size)   // checkModelArgs (loader.go:98-120)
        throw Error("VocabSize=" + std::to_string(transformer.args.vocab_size) + " and vocabulary model length=" +
                    std::to_string(tokenizer->Size()) + " aren't equal");
      vocab = tokenizer->GetVocabulary();
      prompt = tokenizer->Tokenize({{"user", prompt_text}});
    }
    if (batch > 0) {
      const int mod = transformer.args.vocab_size < 128256 ? 1000 : 128000;
      std::vector<std::vector<int32_t>> prompts(batch, prompt);
      for (int b = 0; b < batch; b++)
        for (size_t j = 1; j < prompt.size(); j++) prompts[b][j] = (prompt[j] + 977 * b) % mod;
      std::vector<std::vector<int32_t>> outs(batch);
      auto b0 = std::chrono::steady_clock::now();
      inference::GenerateTokensBatch(transformer, vocab, seq_len, acc, prompts,
                                     [&](int seq, inference::GenerationState, int32_t tok) { outs[seq].push_back(tok); });
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
      size_t total = 0;
      for (int b = 0; b < batch; b++) {
        printf("tokens[%d]:", b);
        for (int32_t t : outs[b]) printf(" %d", t);
        printf("\n");
        total += outs[b].size();
      }
      printf("generated %zu tokens for %d prompts in %.1f ms = %.1f tokens/s aggregate (prefill included)\n", total, batch, sec * 1e3, total / sec);
      return 0;
    }
    std::vector<int32_t> out;
    auto t0 = std::chrono::steady_clock::now();
    std::chrono::steady_clock::time_point t_first;
    inference::GenerateTokens(transformer, vocab, seq_len, acc, prompt, [&](inference::GenerationState, int32_t tok) {
      if (out.empty()) t_first = std::chrono::steady_clock::now();
      out.push_back(tok);
    });
    auto t1 = std::chrono::steady_clock::now();
    if (tokenizer) printf("text: %s\n", tokenizer->TokenBatchToString(out).c_str());
    printf("tokens:");
    for (int32_t t : out) printf(" %d", t);
    const double dec_s = std::chrono::duration<double>(t1 - t_first).count();
    printf("\ngenerated %zu tokens (%s); prefill+first token %.1f ms; decode %.2f tokens/s through the reference-shaped host loop\n",
           out.size(), acc == LNB_ACC_STRICT ? "strict" : "fast", std::chrono::duration<double>(t_first - t0).count() * 1e3,
           out.size() > 1 ? (out.size() - 1) / dec_s : 0.0);
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
