// gemm_tc.cuh -- the prompt-processing projection: C[M,N] = X[M,K] . W[N,K]^T on the 5th-generation
// tensor cores (tcgen05.mma, bf16 in, fp32 accumulate in TMEM), operands staged into shared memory by
// the TMA engine's bulk-async copies.  Replaces ml.LinearTransformation
// (src/ml/operations_impl.go:427-447 -> operations_lineartransform.go:145-207) when S is large
// (prefill); the decode path (S = 1..8) uses gemv.cuh.
//
// Numerics: products are exact, accumulation is fp32 in the tensor core's own order, the result is
// truncated to bf16 exactly like the reference's ToBFloat16 (:205).  The accumulation ORDER is the
// hardware's, so this path belongs to LNB_ACC_FAST; LNB_ACC_STRICT keeps the sequential GEMV.
//
// Operand layouts (both "K-major, no swizzle" in UMMA terms: 8-row x 16-byte core matrices):
//   W  : the same panel-major HBM layout the GEMV streams (8-row panels = core matrices, gemv.cuh): the 16
//        panels of a 128-row tile are copied panel by panel (KT*16 B each) and form ONE N=128 B operand:
//        LBO = 128 B (next k-chunk), SBO = KT*16 B (next panel).
//   X  : "X8" activations written by the preceding kernel, tile-major: [M/128][K/128] tiles of
//        [16 row-groups][16 k-chunks][8 rows][8 elems] = 32 KB contiguous -> ONE bulk copy per stage;
//        LBO = 128 B (next k-chunk), SBO = 2048 B (next 8 rows).
// One CTA computes a 128 x 128 tile: 8 accumulators of 128 lanes x 16 columns in TMEM (128 columns).
// Warp roles: 0, 2, 3 = bulk-copy producers (one thread each), 1 = MMA issuer (one elected lane), 2 also allocates TMEM,
// 4..7 = epilogue (tcgen05.ld -> truncate -> global).
#pragma once
#include "common.cuh"

namespace lnb {

enum { TC_EPI_BF16 = 0, TC_EPI_RESID = 1, TC_EPI_F32TRUNC = 2, TC_EPI_F32RAW = 3, TC_EPI_SWIGLU = 4 };

struct GemmTcParams {
  const uint16_t* X8;   // activations, X8 layout, M padded to a multiple of 128 (padding rows are zero)
  const uint16_t* W;    // weights, 16-row panel-major
  int M, N, K;          // M = valid rows; N % 128 == 0; K % 64 == 0
  uint16_t* out_bf16;   // [M, ldo] row-major (TC_EPI_BF16 / TC_EPI_RESID)
  float* out_f32;       // [M, ldo] (TC_EPI_F32TRUNC: f32(t(acc)); TC_EPI_F32RAW: acc)
  const uint16_t* res;  // [M, ldo] residual (TC_EPI_RESID)
  int ldo;
  // TC_EPI_SWIGLU (W = the stacked w1|w3 matrix, half-panel interleave: columns 8p..8p+3 = gate of hidden units 4p..4p+3,
  // 8p+4..8p+7 = their up values): out_bf16 = the X8 operand of the w2 GEMM, ldo = its K (ffn), m = t(t(TABLE_SILU[t(g)]) * t(u))
  // (llamatransformer.go:601-614); padding rows of the last M tile are written too (they come out as 0)
  const uint16_t* silu_tab;
};

constexpr int TC_BM = 128, TC_BN = 128, TC_KT = 128, TC_NS = 3;
constexpr int TC_A_STAGE = TC_BM * TC_KT * 2;  // 32 KB
constexpr int TC_B_STAGE = TC_BN * TC_KT * 2;  // 32 KB
constexpr int TC_SMEM = 1024 + TC_NS * (TC_A_STAGE + TC_B_STAGE);
constexpr int TC_THREADS = 256;

// ---- tcgen05 wrappers (PTX as in /opt/skills/guides/blackwell_cuda_programming.md) ---------------
LNB_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LNB_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
LNB_DEVINL void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LNB_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute/arch/mma_sm100_desc.hpp bit layout)
LNB_DEVINL uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);            // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;  // leading byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;  // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  return d;                                           // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M x N
LNB_DEVINL constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
LNB_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Optional A operand from TMEM (".ts" form, -DLNB_TC_A_IN_TMEM=1): the 128 x 16 activation slice of a k16
// step is copied shared -> tensor memory (tcgen05.cp, 128 lanes x 256 bit) and the MMA reads it from there.
// Verified correct on B200; not faster than the SS form for this tile shape, so it is off by default.
LNB_DEVINL void tmem_cp_128x256b(uint32_t taddr, uint64_t s_desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}
LNB_DEVINL void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// generic-proxy writes (st.shared / st.global by threads) before async-proxy reads (tcgen05.mma operands, bulk copies)
LNB_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
LNB_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of the (converged) warp, chosen by the hardware: ptxas then knows a single lane issues the tcgen05 instructions
// under this predicate and moves their operands to uniform registers once, without a loop over possibly distinct values
LNB_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0u;
}
#ifndef LNB_TC_A_IN_TMEM
#define LNB_TC_A_IN_TMEM 0
#endif
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
LNB_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LNB_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- layout helpers for the tensor-core path (X8 = tile-major activations, see the header comment)
// X8 address of element (row, col) of a [*, K] activation matrix
LNB_DEVINL size_t x8_index(int row, int col, int K) {
  const size_t tile = (size_t)(row >> 7) * (K >> 7) + (col >> 7);                 // [M/128][K/128]
  const int g = (row >> 3) & 15, ch = (col >> 3) & 15;                            // [16 groups][16 chunks]
  return (((tile * 16 + g) * 16 + ch) * 8 + (row & 7)) * 8 + (col & 7);
}

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [NS]
  uint64_t* empty_bar = full_bar + TC_NS;                  // [NS]
  uint64_t* acc_bar = empty_bar + TC_NS;                   // accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  uint8_t* sA = smem + 1024;
  uint8_t* sB = sA + TC_NS * TC_A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.x walks the M tiles: consecutive CTAs share one weight tile (read from HBM once, then L2 hits)
  const int m0 = blockIdx.x * TC_BM;   // first activation row
  const int n0 = blockIdx.y * TC_BN;   // first weight row of this tile
  const int n_kt = p.K / TC_KT;

  if (tid == 0) {
    for (int s = 0; s < TC_NS; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, LNB_TC_A_IN_TMEM ? 256 : 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0 || warp == 2 || warp == 3) {
    // Three producer threads (warps 0, 2, 3): a stage is 1 + 16 bulk copies, and ONE thread needs ~50 cycles per copy -- 850
    // cycles per stage against the 512 cycles its eight MMAs occupy the tensor pipe (ncu: tensor pipe 34 % active, L2 -> SM
    // traffic at a third of what the SM can take: the pipe was waiting for the copy ISSUE, not for bandwidth).  Warp 0 announces
    // the stage's bytes and copies the activation tile and 4 weight panels, warps 2 and 3 copy 6 panels each; a copy that
    // completes before the announcement only makes the barrier's pending-byte count transiently negative.
    if (lane == 0) {
      pdl_wait();  // the activations come from the previous kernel (weights alone could go earlier)
      const int part = (warp == 0) ? 0 : warp - 1;
      const int j0 = (part == 0) ? 0 : (part == 1 ? 4 : 10), j1 = (part == 0) ? 4 : (part == 1 ? 10 : 16);
      const uint64_t pol_w = l2_policy_evict_first();
      const uint64_t pol_x = l2_policy_evict_last();   // every N-tile re-reads the same activations
      const uint8_t* xb = reinterpret_cast<const uint8_t*>(p.X8);
      const uint8_t* wb = reinterpret_cast<const uint8_t*>(p.W);
      for (int t = 0; t < n_kt; t++) {
        const int s = t % TC_NS;
        mbar_wait(&empty_bar[s], (((uint32_t)(t / TC_NS)) & 1u) ^ 1u);
        const size_t k0 = (size_t)t * TC_KT;
        if (part == 0) {
          mbar_expect_tx(&full_bar[s], TC_A_STAGE + TC_B_STAGE);
          // A: one contiguous 128 x 128 tile of the tile-major X8 layout
          bulk_g2s(sA + (size_t)s * TC_A_STAGE, xb + ((size_t)(m0 / TC_BM) * (p.K / TC_KT) + t) * TC_A_STAGE, TC_A_STAGE,
                   &full_bar[s], pol_x);
        }
        // B: 16 panels of 8 rows, each (8 rows x KT) = KT*16 contiguous bytes
        for (int j = j0; j < j1; j++)
          bulk_g2s(sB + (size_t)s * TC_B_STAGE + (size_t)j * (TC_KT * 16),
                   wb + ((size_t)(n0 / 8 + j) * p.K + k0) * 16, TC_KT * 16, &full_bar[s], pol_w);
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop, one elected lane issues: the operands are then warp-uniform for ptxas and sit in uniform
    // registers (from a `lane == 0` branch every MMA was preceded by a loop of R2UR transfers, ~120 cycles -- twice the
    // 64 cycles a 128 x 128 x 16 MMA occupies the tensor pipe)
    const bool lead = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
    for (int t = 0; t < n_kt; t++) {
      const int s = t % TC_NS;
      mbar_wait(&full_bar[s], ((uint32_t)(t / TC_NS)) & 1u);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + (size_t)s * TC_A_STAGE);
      const uint32_t b_base = smem_u32(sB + (size_t)s * TC_B_STAGE);
      uint64_t a_desc = umma_desc(a_base, 128, TC_KT * 16);  // SBO = 16 chunks * 128 B; next k16 step = +256 B = +16 in the address field
      uint64_t b_desc = umma_desc(b_base, 128, TC_KT * 16);
#pragma unroll
      for (int k16 = 0; k16 < TC_KT / 16; k16++) {
#if LNB_TC_A_IN_TMEM
        const uint32_t a_tmem = tmem_base + 128 + (uint32_t)(k16 & 1) * 8;   // 8 columns = 16 bf16 per lane, double-buffered
        if (lead) {
          tmem_cp_128x256b(a_tmem, a_desc);
          umma_bf16_ts(tmem_base, a_tmem, b_desc, idesc, (t > 0 || k16 > 0) ? 1u : 0u);
        }
#else
        if (lead) umma_bf16(tmem_base, a_desc, b_desc, idesc, (t > 0 || k16 > 0) ? 1u : 0u);   // 128 x 128 x 16 per instruction
#endif
        a_desc += 16; b_desc += 16;
      }
      if (lead) umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
    }
    if (lead) umma_commit(acc_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global ================================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;     // activation row == TMEM lane
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    if (EPI == TC_EPI_RESID) pdl_wait();    // residual operand is written by an earlier kernel
#pragma unroll 1
    for (int j = 0; j < TC_BN / 16; j++) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 16), r);
      if (EPI == TC_EPI_SWIGLU) {
        uint16_t sg[8];
#pragma unroll
        for (int e = 0; e < 8; e++) sg[e] = p.silu_tab[r[(e >> 2) * 8 + (e & 3)] >> 16];          // t(TABLE_SILU[t(gate)])
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float lo = __fmul_rn(bf2f(sg[e]), __uint_as_float(r[(e >> 2) * 8 + 4 + (e & 3)] & 0xffff0000u));
          const float hi = __fmul_rn(bf2f(sg[e + 1]), __uint_as_float(r[((e + 1) >> 2) * 8 + 4 + ((e + 1) & 3)] & 0xffff0000u));
          o[e >> 1] = (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u);
        }
        *reinterpret_cast<uint4*>(p.out_bf16 + x8_index(row, (n0 + j * 16) >> 1, p.ldo)) = make_uint4(o[0], o[1], o[2], o[3]);
      } else if (row < p.M) {
        const size_t o = (size_t)row * p.ldo + n0 + j * 16;
        if (EPI == TC_EPI_BF16 || EPI == TC_EPI_RESID) {
          uint32_t packed[8];
#pragma unroll
          for (int e = 0; e < 8; e++) {
            float a = __uint_as_float(r[2 * e]), b = __uint_as_float(r[2 * e + 1]);
            if (EPI == TC_EPI_RESID) {
              const uint32_t rw = *reinterpret_cast<const uint32_t*>(p.res + o + 2 * e);
              a = __fadd_rn(bf_lo(rw), trunc_bf(a));   // ml.Add(x, t(linear))  llamatransformer.go:232,248
              b = __fadd_rn(bf_hi(rw), trunc_bf(b));
            }
            packed[e] = (uint32_t)f2bf(a) | ((uint32_t)f2bf(b) << 16);
          }
          uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + o);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        } else {
          float4* dst = reinterpret_cast<float4*>(p.out_f32 + o);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            float4 v;
            v.x = __uint_as_float(r[4 * e]); v.y = __uint_as_float(r[4 * e + 1]);
            v.z = __uint_as_float(r[4 * e + 2]); v.w = __uint_as_float(r[4 * e + 3]);
            if (EPI == TC_EPI_F32TRUNC) { v.x = trunc_bf(v.x); v.y = trunc_bf(v.y); v.z = trunc_bf(v.z); v.w = trunc_bf(v.w); }
            dst[e] = v;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, LNB_TC_A_IN_TMEM ? 256 : 128);
  }
}


// ---- -----------------------------------------------------
// row-major [M, K] bf16 -> X8 ([Mpad/8][K/8][8][8]); rows >= M are written as zeros
__global__ void pack_x8_kernel(const uint16_t* __restrict__ src, int ld, int M, int Mpad, int K, uint16_t* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t chunks = K / 8, total = (int64_t)Mpad * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / chunks, ch = i % chunks;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < M) v = *reinterpret_cast<const uint4*>(src + row * ld + ch * 8);
    *reinterpret_cast<uint4*>(dst + x8_index((int)row, (int)ch * 8, K)) = v;
  }
}

// X8 -> row-major [M, K] (tests / op-level API)
__global__ void unpack_x8_kernel(const uint16_t* __restrict__ src, int M, int K, uint16_t* __restrict__ dst, int ld) {
  const int64_t chunks = K / 8, total = (int64_t)M * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / chunks, ch = i % chunks;
    *reinterpret_cast<uint4*>(dst + row * ld + ch * 8) = *reinterpret_cast<const uint4*>(src + x8_index((int)row, (int)ch * 8, K));
  }
}

// RMSNorm.Forward (llamatransformer.go:633-660) for the prefill path: one CTA (256 threads) per row,
// LNB_ACC_FAST reduction order (256 interleaved partial sums, butterfly, then over warps -- the same
// order as gemv.cuh's prologue), output written straight into the X8 layout the tensor-core GEMM reads.
// Rows in [M, Mpad) are zero-filled (grid = Mpad).
__global__ void __launch_bounds__(256) rmsnorm_x8_kernel(const uint16_t* __restrict__ x, int ldx, const uint16_t* __restrict__ w,
                                                         uint16_t* __restrict__ out_x8, int M, int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float wsum[8];
  __shared__ float rs;
  const int row = blockIdx.x, tid = threadIdx.x;
  if (row >= M) {
    for (int k = tid; k < D; k += 256) out_x8[x8_index(row, k, D)] = 0;
    return;
  }
  const uint16_t* xr = x + (size_t)row * ldx;
  float sum = 0.f;
  for (int k = tid; k < D; k += 256) {
    const float v = bf2f(xr[k]);
    sum = __fmaf_rn(v, v, sum);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
  if ((tid & 31) == 0) wsum[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; i++) tot = __fadd_rn(tot, wsum[i]);
    const float me = __fadd_rn(__fdiv_rn(tot, (float)D), eps);
    rs = (float)(1.0 / sqrt((double)me));
  }
  __syncthreads();
  const float r = rs;
  for (int k = tid; k < D; k += 256) {
    const float n1 = trunc_bf(__fmul_rn(bf2f(xr[k]), r));
    out_x8[x8_index(row, k, D)] = f2bf(__fmul_rn(n1, bf2f(w[k])));
  }
}

// After the fused wq|wk|wv GEMM of the prefill path: RoPE on q and k (f64 intermediates, see
// gemv.cuh EPI_QKV_ROPE), q -> q_out [S, q_dim], k -> cacheK[pos0+s], v -> cacheV[pos0+s]
// (llamatransformer.go:374-403).  One thread per (row, pair of columns).
__global__ void rope_kv_kernel(const uint16_t* __restrict__ qkv, int ld, uint16_t* __restrict__ q_out, int q_dim, int kv_dim,
                               int head_dim, uint16_t* __restrict__ cache_k, uint16_t* __restrict__ cache_v,
                               const float* __restrict__ cis, const int32_t* __restrict__ pos_ptr, int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int pairs = (q_dim + 2 * kv_dim) / 2;
  const int64_t total = (int64_t)S * pairs;
  const int pos0 = *pos_ptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(i / pairs), n = (int)(i % pairs) * 2;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(qkv + (size_t)s * ld + n);
    const int pos = pos0 + s;
    if (n < q_dim + kv_dim) {
      const int nn = (n < q_dim) ? n : n - q_dim;
      const int ii = (nn % head_dim) >> 1;
      const float2 fc = *reinterpret_cast<const float2*>(cis + ((size_t)pos * (head_dim / 2) + ii) * 2);
      const double a = (double)bf_lo(w), b = (double)bf_hi(w), c = (double)fc.x, d = (double)fc.y;
      const uint32_t o = (uint32_t)f2bf((float)(a * c - b * d)) | ((uint32_t)f2bf((float)(a * d + b * c)) << 16);
      if (n < q_dim) *reinterpret_cast<uint32_t*>(q_out + (size_t)s * q_dim + n) = o;
      else *reinterpret_cast<uint32_t*>(cache_k + (size_t)pos * kv_dim + nn) = o;
    } else {
      *reinterpret_cast<uint32_t*>(cache_v + (size_t)pos * kv_dim + (n - q_dim - kv_dim)) = w;
    }
  }
}

// After the fused w1|w3 GEMM (columns follow the half-panel interleave of the stacked matrix:
// 4 gate columns, then the 4 up columns of the same hidden units): m = t( t(TABLE_SILU[g]) * u )
// (llamatransformer.go:601-614), written in X8 layout for the w2 GEMM.  Rows in [M, Mpad) -> 0.
__global__ void swiglu_x8_kernel(const uint16_t* __restrict__ gu, int ld, const uint16_t* __restrict__ silu_tab,
                                 uint16_t* __restrict__ out_x8, int M, int Mpad, int ffn) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)Mpad * ffn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / ffn), c = (int)(i % ffn);
    uint16_t o = 0;
    if (row < M) {
      const size_t base = (size_t)row * ld + (size_t)(c >> 2) * 8 + (c & 3);
      const uint16_t sg = silu_tab[gu[base]];
      o = f2bf(__fmul_rn(bf2f(sg), bf2f(gu[base + 4])));
    }
    out_x8[x8_index(row, c, ffn)] = o;
  }
}

}  // namespace lnb
