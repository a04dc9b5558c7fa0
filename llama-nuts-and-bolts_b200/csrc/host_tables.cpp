// host_tables.cpp -- init-time tables the library builds when the host does not supply them.
// Product code (not the oracle): the Go host normally computes these itself
// (model.precomputeFreqsCis, ml.TABLE_SILU) and may upload them through
// lnb_model_set_rope_table / lnb_model_set_silu_table instead.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace lnb {

static inline float bf_to_f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f_to_bf(float f) {  // truncation, src/dtype/bfloat16.go:59-61
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);
}

// model.precomputeFreqsCis + applyScaling (src/model/llamatransformer.go:662-751).
// The table is built in bf16 like the reference: inverse frequencies, positions and their
// products are all truncated to bf16 before cos/sin are taken in float64.
// out: [end][dim/2][2] float (cos, sin) == the complex64 tensor.
void build_rope_table(int dim, int end, double theta, bool use_scaled, std::vector<float>& out) {
  const int half = dim / 2;
  std::vector<uint16_t> freqs(half);
  const float dimf = (float)dim;
  for (int i = 0; i < half; i++) {
    const float val = bf_to_f(f_to_bf((float)(2 * i)));              // ARange(0, dim, 2, BF16)
    const float f = (float)(1.0 / std::pow(theta, (double)(val / dimf)));
    freqs[i] = f_to_bf(f);
  }
  if (use_scaled) {  // Llama-3.1 frequency remap, evaluated in float32 (applyScaling :662-692)
    const float scale_factor = 8.0f, low_freq_factor = 1.0f, high_freq_factor = 4.0f, old_context_len = 8192.0f;
    const float low_freq_wavelen = old_context_len / low_freq_factor;
    const float high_freq_wavelen = old_context_len / high_freq_factor;
    const float two_pi = (float)(2.0 * 3.14159265358979323846264338327950288);
    for (int i = 0; i < half; i++) {
      const float freq = bf_to_f(freqs[i]);
      const float wavelen = two_pi / freq;
      float nf;
      if (wavelen < high_freq_wavelen) {
        nf = freq;
      } else if (wavelen > low_freq_wavelen) {
        nf = freq / scale_factor;
      } else {
        const float smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor);
        volatile float t1 = (1.0f - smooth) * freq;  // volatile: keep every f32 rounding (no contraction)
        volatile float t1s = t1 / scale_factor;
        volatile float t2 = smooth * freq;
        nf = t1s + t2;
      }
      freqs[i] = f_to_bf(nf);
    }
  }
  out.resize((size_t)end * half * 2);
  for (int p = 0; p < end; p++) {
    const float tp = bf_to_f(f_to_bf((float)p));                     // ARange(0, end, 1, BF16)
    for (int i = 0; i < half; i++) {
      volatile float prod = tp * bf_to_f(freqs[i]);                  // ml.Outer, f32 product
      const double ang = (double)bf_to_f(f_to_bf(prod));             // ... stored as bf16
      out[((size_t)p * half + i) * 2 + 0] = (float)std::cos(ang);    // ml.Polar: complex64(complex(cos, sin))
      out[((size_t)p * half + i) * 2 + 1] = (float)std::sin(ang);
    }
  }
}

// ml.TABLE_SILU (src/ml/activations.go:11-25) already passed through the truncating store
// of ml.Silu's BF16 branch (:38): out[b] = t( f32( v / (1 + exp(-v)) ) ), v = f64(bf16 b).
void build_silu_table(std::vector<uint16_t>& out) {
  out.resize(65536);
  for (int i = 0; i < 65536; i++) {
    const double v = (double)bf_to_f((uint16_t)i);
    out[i] = f_to_bf((float)(v / (1.0 + std::exp(-v))));
  }
}

}  // namespace lnb
