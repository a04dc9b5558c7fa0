// Host model of rms_scale_scan_kernel's control flow (llama-nuts-and-bolts_b200/csrc/seqsum.cuh), built on the
// SAME seq_term / seq_compose / seq_anchor_ok code the kernel compiles, with the CTA-wide scan written as a
// plain loop.  Test infrastructure: lets the CPU suite prove the binade-scan algorithm bit-exact against the
// reference's one-accumulator loop on adversarial inputs without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>
#include "seqsum.cuh"

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

extern "C" float seqsum_reference(const float* terms, int n) {
  volatile float s = 0.f;
  for (int i = 0; i < n; i++) s = s + terms[i];
  return s;
}

// n = nt * ch terms; returns the modelled kernel's sum; *iters = trips through the scan loop
extern "C" float seqsum_scan_model(const float* terms, int nt, int ch, int* iters) {
  uint32_t s_sum = 0u;
  int c0 = 0, it = 0;
  std::vector<SeqInc> g(nt);
  std::vector<uint32_t> tot(nt);
  while (c0 < nt) {
    it++;
    const uint32_t sb = s_sum;
    if (!seq_anchor_ok(sb)) {
      volatile float s = u2f(sb);
      for (int k = 0; k < ch; k++) s = s + terms[c0 * ch + k];
      s_sum = f2u(s);
      c0++;
      continue;
    }
    const int E = (int)(sb >> 23);
    const uint32_t S = (sb & 0x7fffffu) | 0x800000u;
    SeqInc run; run.i0 = run.i1 = 0u;
    int first = nt;
    for (int t = 0; t < nt; t++) {
      SeqInc m; m.i0 = m.i1 = 0u;
      if (t >= c0) for (int k = 0; k < ch; k++) m = seq_compose(m, seq_term(f2u(terms[t * ch + k]), E));
      run = seq_compose(run, m);
      g[t] = run;
      tot[t] = S + seq_eval(run, S & 1u);
      if (t >= c0 && tot[t] >= (1u << 24) && first == nt) first = t;
    }
    if (first >= nt) {
      s_sum = ((uint32_t)E << 23) + (tot[nt - 1] - (1u << 23));
      c0 = nt;
    } else {
      const uint32_t prev = (first <= c0) ? S : tot[first - 1];
      volatile float s = u2f(((uint32_t)E << 23) + (prev - (1u << 23)));
      for (int k = 0; k < ch; k++) s = s + terms[first * ch + k];
      s_sum = f2u(s);
      c0 = first + 1;
    }
  }
  if (iters) *iters = it;
  return u2f(s_sum);
}

// Host model of rms_scale_seg_kernel (predict -> fold runs -> walk).  `sabotage` != 0 corrupts the binade
// predictions pseudo-randomly: the walk's validity test must keep the result exact whatever it is fed.
extern "C" float seqsum_seg_model(const float* terms, int nt, int ch, uint32_t sabotage, int* jumps, int* walked) {
  std::vector<int> code(nt);
  std::vector<SeqSeg> v(nt);
  std::vector<int> run_end(nt, 0);
  std::vector<SeqInc> run_map(nt);
  uint32_t rng = sabotage * 2654435761u + 12345u;
  float before = 0.f;
  for (int t = 0; t < nt; t++) {
    float cs = 0.f;
    for (int k = 0; k < ch; k++) cs += terms[t * ch + k];
    const float after = before + cs;
    code[t] = (t == 0) ? 0 : seq_predict(before, after);
    before = after;
    if (sabotage) {
      rng = rng * 1664525u + 1013904223u;
      const uint32_t roll = (rng >> 8) % 16u;
      if (roll == 0) code[t] = 0;
      else if (roll == 1 && t) code[t] = code[t] ? code[t] + 1 : 127;
      else if (roll == 2 && t) code[t] = code[t] > 1 ? code[t] - 1 : 120;
      else if (roll == 3 && t) code[t] = 1 + (int)((rng >> 12) % 254u);
      if (code[t] > 254) code[t] = 254;
    }
  }
  SeqSeg acc; acc.flag = 0; acc.head = 0; acc.m.i0 = acc.m.i1 = 0;
  for (int t = 0; t < nt; t++) {
    SeqSeg e; e.m.i0 = e.m.i1 = 0u;
    if (code[t]) for (int k = 0; k < ch; k++) e.m = seq_compose(e.m, seq_term(f2u(terms[t * ch + k]), code[t]));
    const int pc = t ? code[t - 1] : 0;
    e.flag = (code[t] == 0 || pc != code[t]) ? 1u : 0u;
    e.head = (uint32_t)t;
    acc = t ? seq_seg_op(acc, e) : e;
    v[t] = acc;
  }
  for (int t = 0; t < nt; t++) {
    const int nc = (t + 1 < nt) ? code[t + 1] : 0;
    if (code[t] && nc != code[t]) { run_end[v[t].head] = t + 1; run_map[v[t].head] = v[t].m; }
  }
  uint32_t sb = 0u;
  int c = 0, nj = 0, nwalk = 0;
  while (c < nt) {
    uint32_t nb;
    if (run_end[c] > c && seq_try_jump(sb, code[c], run_map[c], &nb)) { sb = nb; c = run_end[c]; nj++; continue; }
    volatile float s = u2f(sb);
    for (int k = 0; k < ch; k++) s = s + terms[c * ch + k];
    sb = f2u(s);
    c++; nwalk++;
  }
  if (jumps) *jumps = nj;
  if (walked) *walked = nwalk;
  return u2f(sb);
}
