// seqsum.cuh -- the reference's strictly sequential fp32 sum of NON-NEGATIVE terms, bit for bit, in
// O(log) depth instead of one dependent FADD per element.
//
// Where it is used: RMSNorm's mean of squares (src/model/llamatransformer.go:641-656 -> ml.Pow, ml.Mean:
// one float32 accumulator walking the row left to right).  A 4096-element row costs 4096 dependent FADDs
// = 18k cycles on one thread (13 us, 65 times per token = 17 % of the STRICT decode step).
//
// Why a parallel evaluation can be exact.  Let s be the running sum, E its binade, u = ulp(s) = 2^(E-150)
// (E = biased exponent), S = s/u an integer in [2^23, 2^24).  For a term x >= 0 and as long as the sum stays
// in the binade,  RN(s + x) = (S + inc) * u  with  x/u = q + f,  q integer, 0 <= f < 1:
//     f < 1/2 -> inc = q        f > 1/2 -> inc = q + 1        f = 1/2 -> inc = q or q + 1, whichever makes S + inc EVEN
// so a term acts on the state only through the PARITY of S: it is a map  parity -> increment, stored as the
// pair (i0, i1).  Such maps compose associatively ( (g1;g2)(p) = g1(p) + g2(p xor (g1(p)&1)) ), so the whole
// tail of the row is one prefix scan.  The sum is monotone (terms >= 0), so "the prefix up to chunk c stays
// below 2^24" proves that no intermediate sum left the binade and the integer model was exact up to there.
// The first chunk whose prefix reaches 2^24 is re-done with real FADDs by one thread (8 dependent adds),
// the binade is re-read from the result and the scan restarts behind it.  A row crosses about
// log2(n_chunks) binades, so the loop runs a dozen times instead of 4096 dependent adds.
// Anything irregular (zero / subnormal / non-finite running sum) takes the real-FADD chunk path, which is
// the reference arithmetic itself -- the result is bit-identical for every input, only the speed varies.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define SEQ_HD __host__ __device__ __forceinline__
#else
#define SEQ_HD inline
#endif

struct SeqInc {
  uint32_t i0, i1;  // increment in ulps when the running integer sum is even / odd; saturates at SEQ_SAT
};
#define SEQ_SAT (1u << 25)  // >= 2^24: any saturated prefix is a binade crossing; sums of two never overflow 32 bits

SEQ_HD uint32_t seq_sat(uint32_t v) { return v < SEQ_SAT ? v : SEQ_SAT; }

// map of one term x (fp32 bit pattern, finite, sign bit clear) against a running sum of biased exponent E (1..254)
SEQ_HD SeqInc seq_term(uint32_t xbits, int E) {
  SeqInc r;
  const int ex = (int)(xbits >> 23);
  const uint32_t m = ex ? ((xbits & 0x7fffffu) | 0x800000u) : (xbits & 0x7fffffu);  // x = m * 2^(max(ex,1) - 150)
  const int sh = E - (ex ? ex : 1);                                                  // x / u = m * 2^-sh
  if (m == 0u || sh >= 25) { r.i0 = r.i1 = 0u; return r; }                           // x/u < 1/2 (m < 2^24)
  if (sh <= 0) {                                                                     // x >= 2^(E-150+23): leaves the binade
    const uint32_t q = (-sh >= 2) ? SEQ_SAT : seq_sat(m << (-sh));
    r.i0 = r.i1 = q;
    return r;
  }
  const uint32_t q = m >> sh, rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
  if (rem < half) { r.i0 = r.i1 = q; }
  else if (rem > half) { r.i0 = r.i1 = q + 1u; }
  else { r.i0 = q + (q & 1u); r.i1 = q + 1u - (q & 1u); }                            // tie: land on an even sum
  return r;
}

// first g1, then g2
SEQ_HD SeqInc seq_compose(SeqInc g1, SeqInc g2) {
  SeqInc r;
  r.i0 = seq_sat(g1.i0 + ((g1.i0 & 1u) ? g2.i1 : g2.i0));
  r.i1 = seq_sat(g1.i1 + ((g1.i1 & 1u) ? g2.i0 : g2.i1));
  return r;
}
SEQ_HD uint32_t seq_eval(SeqInc g, uint32_t parity) { return parity ? g.i1 : g.i0; }

// running sum usable as a binade anchor: normal and finite
SEQ_HD bool seq_anchor_ok(uint32_t sbits) {
  const uint32_t e = (sbits >> 23) & 0xffu;
  return (sbits >> 31) == 0u && e >= 1u && e <= 254u;
}

// ---- one-pass variant: predict every chunk's binade first --------------------------------------------
// An approximate (parallel, differently rounded) prefix sum A tells in which binade the exact running sum
// will be while it walks over chunk c -- except near a power of two.  seq_predict() returns that binade when
// [A_c (1 - 2^-10), A_{c+1} (1 + 2^-10)] lies inside one binade (the sequential sum differs from the exact
// prefix by < n 2^-24 <= 2^-11 relative), else 0 = "uncertain".  Runs of chunks with the same prediction are
// folded into ONE parity map by a segmented scan; a single thread then walks the row segment by segment:
//   head of a run, running sum in the predicted binade, S + map(parity) < 2^24  ->  jump over the whole run
//   anything else (uncertain chunk, failed check)                               ->  the chunk's real FADDs
// The prediction only decides how fast the walk is: every jump is guarded by the exact validity test above.
SEQ_HD uint32_t seq_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
SEQ_HD float seq_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
SEQ_HD int seq_predict(float approx_before, float approx_after) {
  const uint32_t lo = seq_f2u(approx_before * 0.9990234375f), hi = seq_f2u(approx_after * 1.0009765625f);
  const uint32_t elo = lo >> 23, ehi = hi >> 23;   // non-negative inputs: no sign bit (NaN -> 255 -> rejected)
  return (elo == ehi && elo >= 1u && elo <= 254u) ? (int)elo : 0;
}
// jump over a run: running sum bits sb, predicted binade E, the run's map; returns false when the run must be walked
SEQ_HD bool seq_try_jump(uint32_t sb, int E, SeqInc run, uint32_t* sb_out) {
  if ((int)(sb >> 23) != E) return false;          // also rejects sign / inf / nan (E <= 254)
  const uint32_t S = (sb & 0x7fffffu) | 0x800000u;
  const uint32_t tot = S + seq_eval(run, S & 1u);
  if (tot >= (1u << 24)) return false;             // left the binade somewhere inside the run
  *sb_out = ((uint32_t)E << 23) + (tot - (1u << 23));
  return true;
}
// segmented-scan element: flag = a run starts here (or inside the aggregated range)
struct SeqSeg {
  uint32_t flag, head;
  SeqInc m;
};
SEQ_HD SeqSeg seq_seg_op(SeqSeg a, SeqSeg b) {     // a earlier, b later
  SeqSeg r;
  r.flag = a.flag | b.flag;
  r.head = b.flag ? b.head : a.head;
  r.m = b.flag ? b.m : seq_compose(a.m, b.m);
  return r;
}

#ifdef __CUDACC__
#include "common.cuh"

// LNB_ACC_STRICT RMSNorm scale  r[row] = f32(1 / sqrt(f64( (sum_seq_k x[row,k]^2) / D + eps )))
// with the sequential sum evaluated by the binade scan above.  grid = rows, block = NT = D / CH threads
// (a multiple of 32, <= SEQ_MAX_THREADS); thread t owns the CH consecutive squares of chunk t.
// seq_scan_shape() picks (CH, NT) or reports that the row length does not fit (-> rms_scale_kernel).
#define SEQ_MAX_THREADS 512
static inline bool seq_scan_shape(int D, int* ch, int* nt) {
  const int cand[4] = {8, 16, 4, 2};
  for (int i = 0; i < 4; i++) {
    const int c = cand[i];
    if (D % c == 0 && (D / c) % 32 == 0 && D / c <= SEQ_MAX_THREADS && D / c >= 32) { *ch = c; *nt = D / c; return true; }
  }
  return false;
}
template <int CH>
__global__ void __launch_bounds__(SEQ_MAX_THREADS) rms_scale_scan_kernel(const uint16_t* __restrict__ x, int ldx, float* __restrict__ r,
                                                                     int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ uint32_t s_i0[SEQ_MAX_THREADS / 32], s_i1[SEQ_MAX_THREADS / 32], s_wtot[SEQ_MAX_THREADS / 32];
  const int NT = (int)blockDim.x;
  __shared__ uint32_t s_sum, s_first;
  __shared__ int s_c0;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const uint16_t* xr = x + (size_t)blockIdx.x * ldx + (size_t)t * CH;
  float sq[CH];
#pragma unroll
  for (int k = 0; k < CH; k += 2) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(xr + k);
    const float a = bf_lo(w), b = bf_hi(w);
    sq[k] = __fmul_rn(a, a);
    sq[k + 1] = __fmul_rn(b, b);
  }
  if (t == 0) { s_sum = 0u; s_c0 = 0; }
  __syncthreads();

  for (;;) {
    const int c0 = s_c0;                 // chunks [0, c0) are folded into s_sum
    if (c0 >= NT) break;
    const uint32_t sb = s_sum;
    if (!seq_anchor_ok(sb)) {            // no binade to anchor on (0, subnormal, inf, nan): plain reference arithmetic
      __syncthreads();                   // everyone has read s_sum / s_c0
      if (t == c0) {
        float s = __uint_as_float(sb);
#pragma unroll
        for (int k = 0; k < CH; k++) s = __fadd_rn(s, sq[k]);
        s_sum = __float_as_uint(s);
        s_c0 = c0 + 1;
      }
      __syncthreads();
      continue;
    }
    const int E = (int)(sb >> 23);
    const uint32_t S = (sb & 0x7fffffu) | 0x800000u;
    // map of my chunk (identity for chunks already folded in)
    SeqInc g;
    g.i0 = g.i1 = 0u;
    if (t >= c0) {
#pragma unroll
      for (int k = 0; k < CH; k++) g = seq_compose(g, seq_term(__float_as_uint(sq[k]), E));
    }
    // inclusive scan over the CTA (maps compose left to right)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SeqInc o;
      o.i0 = __shfl_up_sync(0xffffffffu, g.i0, d);
      o.i1 = __shfl_up_sync(0xffffffffu, g.i1, d);
      if (lane >= d) g = seq_compose(o, g);
    }
    if (lane == 31) { s_i0[wid] = g.i0; s_i1[wid] = g.i1; }
    if (t == 0) s_first = (uint32_t)NT;
    __syncthreads();
    SeqInc pre;
    pre.i0 = pre.i1 = 0u;
    for (int w = 0; w < wid; w++) {
      SeqInc o;
      o.i0 = s_i0[w]; o.i1 = s_i1[w];
      pre = seq_compose(pre, o);
    }
    g = seq_compose(pre, g);
    const uint32_t tot = S + seq_eval(g, S & 1u);      // integer sum after my chunk, exact while < 2^24
    const bool cross = (t >= c0) && tot >= (1u << 24);
    const uint32_t bal = __ballot_sync(0xffffffffu, cross);
    if (bal && lane == (__ffs(bal) - 1)) atomicMin(&s_first, (uint32_t)t);
    // the sum before my chunk = the sum after the previous chunk
    uint32_t prev = __shfl_up_sync(0xffffffffu, tot, 1);
    if (lane == 31) s_wtot[wid] = tot;
    __syncthreads();
    const int first = (int)s_first;
    if (lane == 0) prev = (wid == 0) ? S : s_wtot[wid - 1];
    if (t <= c0) prev = S;
    if (first >= NT) {          // the rest of the row stays in this binade: done
      if (t == NT - 1) {
        s_sum = ((uint32_t)E << 23) + (tot - (1u << 23));   // tot < 2^24: mantissa = tot - 2^23
        s_c0 = NT;
      }
    } else if (t == first) {             // my chunk leaves the binade: redo it with the reference's own FADDs
      float s = __uint_as_float(((uint32_t)E << 23) + (prev - (1u << 23)));
#pragma unroll
      for (int k = 0; k < CH; k++) s = __fadd_rn(s, sq[k]);
      s_sum = __float_as_uint(s);
      s_c0 = first + 1;
    }
    __syncthreads();
  }
  if (t == 0) {
    const float sum = __uint_as_float(s_sum);
    const float me = __fadd_rn(__fdiv_rn(sum, (float)D), eps);
    r[blockIdx.x] = (float)(1.0 / sqrt((double)me));
  }
}
// One-pass variant of rms_scale_scan_kernel (same contract, same launch shape): predict, fold runs, walk.
template <int CH>
__global__ void __launch_bounds__(SEQ_MAX_THREADS) rms_scale_seg_kernel(const uint16_t* __restrict__ x, int ldx, float* __restrict__ r,
                                                                    int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float s_sq[SEQ_MAX_THREADS * CH];
  __shared__ uint32_t s_run_i0[SEQ_MAX_THREADS], s_run_i1[SEQ_MAX_THREADS];
  __shared__ uint16_t s_run_end[SEQ_MAX_THREADS];
  __shared__ uint8_t s_code[SEQ_MAX_THREADS];
  __shared__ float s_wsum[32];
  __shared__ uint32_t s_wflag[32], s_whead[32], s_wi0[32], s_wi1[32];
  const int NT = (int)blockDim.x, nw = NT >> 5;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const uint16_t* xr = x + (size_t)blockIdx.x * ldx + (size_t)t * CH;
  float sq[CH];
#pragma unroll
  for (int k = 0; k < CH; k += 2) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(xr + k);
    const float a = bf_lo(w), b = bf_hi(w);
    sq[k] = __fmul_rn(a, a);
    sq[k + 1] = __fmul_rn(b, b);
  }
#pragma unroll
  for (int k = 0; k < CH; k += 2) *reinterpret_cast<float2*>(&s_sq[t * CH + k]) = make_float2(sq[k], sq[k + 1]);
  // ---- approximate prefix sums (any rounding will do) ------------------------------------------------
  float cs = 0.f;
#pragma unroll
  for (int k = 0; k < CH; k++) cs = __fadd_rn(cs, sq[k]);
  float inc = cs;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc = __fadd_rn(inc, v);
  }
  if (lane == 31) s_wsum[wid] = inc;
  s_run_end[t] = 0;
  __syncthreads();
  float wpre = 0.f;
  {
    float wv = (lane < nw) ? s_wsum[lane] : 0.f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float v = __shfl_up_sync(0xffffffffu, wv, d);
      if (lane >= d) wv = __fadd_rn(wv, v);
    }
    const float p = __shfl_sync(0xffffffffu, wv, (wid + 31) & 31);
    if (wid > 0) wpre = p;
  }
  float before = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) before = 0.f;
  before = __fadd_rn(before, wpre);
  const float after = __fadd_rn(inc, wpre);
  const int code = (t == 0) ? 0 : seq_predict(before, after);
  s_code[t] = (uint8_t)code;
  // ---- my chunk's parity map under the predicted binade ------------------------------------------------
  SeqSeg v;
  v.m.i0 = v.m.i1 = 0u;
  if (code) {
#pragma unroll
    for (int k = 0; k < CH; k++) v.m = seq_compose(v.m, seq_term(__float_as_uint(sq[k]), code));
  }
  int pc = __shfl_up_sync(0xffffffffu, code, 1), nc = __shfl_down_sync(0xffffffffu, code, 1);
  __syncthreads();
  if (lane == 0) pc = t ? (int)s_code[t - 1] : 0;
  if (lane == 31) nc = (t + 1 < NT) ? (int)s_code[t + 1] : 0;
  v.flag = (code == 0 || pc != code) ? 1u : 0u;   // uncertain chunks are runs of their own (never jumped)
  v.head = (uint32_t)t;
  // ---- segmented inclusive scan: composition from the head of my run through my chunk -----------------
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    SeqSeg o;
    o.flag = __shfl_up_sync(0xffffffffu, v.flag, d);
    o.head = __shfl_up_sync(0xffffffffu, v.head, d);
    o.m.i0 = __shfl_up_sync(0xffffffffu, v.m.i0, d);
    o.m.i1 = __shfl_up_sync(0xffffffffu, v.m.i1, d);
    if (lane >= d) v = seq_seg_op(o, v);
  }
  if (lane == 31) { s_wflag[wid] = v.flag; s_whead[wid] = v.head; s_wi0[wid] = v.m.i0; s_wi1[wid] = v.m.i1; }
  __syncthreads();
  {
    SeqSeg a;
    a.flag = (lane < nw) ? s_wflag[lane] : 1u;
    a.head = (lane < nw) ? s_whead[lane] : 0u;
    a.m.i0 = (lane < nw) ? s_wi0[lane] : 0u;
    a.m.i1 = (lane < nw) ? s_wi1[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SeqSeg o;
      o.flag = __shfl_up_sync(0xffffffffu, a.flag, d);
      o.head = __shfl_up_sync(0xffffffffu, a.head, d);
      o.m.i0 = __shfl_up_sync(0xffffffffu, a.m.i0, d);
      o.m.i1 = __shfl_up_sync(0xffffffffu, a.m.i1, d);
      if (lane >= d) a = seq_seg_op(o, a);
    }
    SeqSeg p;
    const int src = (wid + 31) & 31;
    p.flag = __shfl_sync(0xffffffffu, a.flag, src);
    p.head = __shfl_sync(0xffffffffu, a.head, src);
    p.m.i0 = __shfl_sync(0xffffffffu, a.m.i0, src);
    p.m.i1 = __shfl_sync(0xffffffffu, a.m.i1, src);
    if (wid > 0) v = seq_seg_op(p, v);
  }
  if (code && nc != code) {                        // last chunk of a run: publish the run at its head
    s_run_end[v.head] = (uint16_t)(t + 1);
    s_run_i0[v.head] = v.m.i0;
    s_run_i1[v.head] = v.m.i1;
  }
  __syncthreads();
  // ---- the walk ---------------------------------------------------------------------------------------
  if (t == 0) {
    uint32_t sb = 0u;
    int c = 0;
    while (c < NT) {
      const int e = (int)s_run_end[c];
      const int E = (int)s_code[c];
      SeqInc run;
      run.i0 = s_run_i0[c]; run.i1 = s_run_i1[c];
      uint32_t nb;
      if (e > c && seq_try_jump(sb, E, run, &nb)) { sb = nb; c = e; continue; }
      float s = __uint_as_float(sb);
#pragma unroll
      for (int k = 0; k < CH; k++) s = __fadd_rn(s, s_sq[c * CH + k]);
      sb = __float_as_uint(s);
      c++;
    }
    const float me = __fadd_rn(__fdiv_rn(__uint_as_float(sb), (float)D), eps);
    r[blockIdx.x] = (float)(1.0 / sqrt((double)me));
  }
}
#endif  // __CUDACC__
