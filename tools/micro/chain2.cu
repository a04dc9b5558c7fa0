// variants of the real inner loop with explicit (rolled) software pipelining
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0u, 0x1044)); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
#define FMA8(a, wv, xa, xb)                                                                   \
  a = __fmaf_rn(xa.x, bf_lo(wv.x), a); a = __fmaf_rn(xa.y, bf_hi(wv.x), a);                    \
  a = __fmaf_rn(xa.z, bf_lo(wv.y), a); a = __fmaf_rn(xa.w, bf_hi(wv.y), a);                    \
  a = __fmaf_rn(xb.x, bf_lo(wv.z), a); a = __fmaf_rn(xb.y, bf_hi(wv.z), a);                    \
  a = __fmaf_rn(xb.z, bf_lo(wv.w), a); a = __fmaf_rn(xb.w, bf_hi(wv.w), a);

template <int G>
__global__ void chainP(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* ws = (uint4*)sm;
  float* xs = (float*)(sm + (size_t)nchunks * 512);
  for (int i = threadIdx.x; i < nchunks * 32; i += blockDim.x) ws[i] = make_uint4(0x3f803f80u, 0x3f003f00u, 0x3e803e80u, 0x3f803f80u);
  for (int i = threadIdx.x; i < nchunks * 8; i += blockDim.x) xs[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float a = 0.f;
  uint4 wA[G], wB[G];
  float4 xA0[G], xA1[G], xB0[G], xB1[G];
  long long t0 = clock64();
#pragma unroll
  for (int q = 0; q < G; q++) { wA[q] = ws[q * 32 + threadIdx.x]; xA0[q] = *(const float4*)(xs + q * 8); xA1[q] = *(const float4*)(xs + q * 8 + 4); }
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2 * G) {
#pragma unroll
    for (int q = 0; q < G; q++) { int cc = c + G + q; wB[q] = ws[cc * 32 + threadIdx.x]; xB0[q] = *(const float4*)(xs + cc * 8); xB1[q] = *(const float4*)(xs + cc * 8 + 4); }
#pragma unroll
    for (int q = 0; q < G; q++) { FMA8(a, wA[q], xA0[q], xA1[q]) }
    if (c + 2 * G < nchunks) {
#pragma unroll
      for (int q = 0; q < G; q++) { int cc = c + 2 * G + q; wA[q] = ws[cc * 32 + threadIdx.x]; xA0[q] = *(const float4*)(xs + cc * 8); xA1[q] = *(const float4*)(xs + cc * 8 + 4); }
    }
#pragma unroll
    for (int q = 0; q < G; q++) { FMA8(a, wB[q], xB0[q], xB1[q]) }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// x held as packed bf16x2 in smem? no: variant with x via one LDS.128 of 8 bf16 and conversion (fewer LDS, more ALU)
template <int G>
__global__ void chainQ(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* ws = (uint4*)sm;
  uint4* xs = (uint4*)(sm + (size_t)nchunks * 512);   // x as bf16: one uint4 per chunk
  for (int i = threadIdx.x; i < nchunks * 32; i += blockDim.x) ws[i] = make_uint4(0x3f803f80u, 0x3f003f00u, 0x3e803e80u, 0x3f803f80u);
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) xs[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  __syncthreads();
  float a = 0.f;
  uint4 wA[G], wB[G], xA[G], xB[G];
  long long t0 = clock64();
#define FMA8X(a, wv, xv)                                                                             \
  a = __fmaf_rn(bf_lo(xv.x), bf_lo(wv.x), a); a = __fmaf_rn(bf_hi(xv.x), bf_hi(wv.x), a);             \
  a = __fmaf_rn(bf_lo(xv.y), bf_lo(wv.y), a); a = __fmaf_rn(bf_hi(xv.y), bf_hi(wv.y), a);             \
  a = __fmaf_rn(bf_lo(xv.z), bf_lo(wv.z), a); a = __fmaf_rn(bf_hi(xv.z), bf_hi(wv.z), a);             \
  a = __fmaf_rn(bf_lo(xv.w), bf_lo(wv.w), a); a = __fmaf_rn(bf_hi(xv.w), bf_hi(wv.w), a);
#pragma unroll
  for (int q = 0; q < G; q++) { wA[q] = ws[q * 32 + threadIdx.x]; xA[q] = xs[q]; }
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2 * G) {
#pragma unroll
    for (int q = 0; q < G; q++) { int cc = c + G + q; wB[q] = ws[cc * 32 + threadIdx.x]; xB[q] = xs[cc]; }
#pragma unroll
    for (int q = 0; q < G; q++) { FMA8X(a, wA[q], xA[q]) }
    if (c + 2 * G < nchunks) {
#pragma unroll
      for (int q = 0; q < G; q++) { int cc = c + 2 * G + q; wA[q] = ws[cc * 32 + threadIdx.x]; xA[q] = xs[cc]; }
    }
#pragma unroll
    for (int q = 0; q < G; q++) { FMA8X(a, wB[q], xB[q]) }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// FADD-only chain over precomputed products, rolled pipeline
template <int G>
__global__ void chainR(float* out, long long* cyc, int n4) {
  extern __shared__ __align__(16) uint8_t sm[];
  float4* ps = (float4*)sm;
  for (int i = threadIdx.x; i < n4 * 32; i += blockDim.x) ps[i] = make_float4(1.f, .5f, .25f, 1.f);
  __syncthreads();
  float a = 0.f;
  float4 A[G], B[G];
  long long t0 = clock64();
#pragma unroll
  for (int q = 0; q < G; q++) A[q] = ps[q * 32 + threadIdx.x];
#pragma unroll 1
  for (int c = 0; c < n4; c += 2 * G) {
#pragma unroll
    for (int q = 0; q < G; q++) B[q] = ps[(c + G + q) * 32 + threadIdx.x];
#pragma unroll
    for (int q = 0; q < G; q++) { a = __fadd_rn(a, A[q].x); a = __fadd_rn(a, A[q].y); a = __fadd_rn(a, A[q].z); a = __fadd_rn(a, A[q].w); }
    if (c + 2 * G < n4) {
#pragma unroll
      for (int q = 0; q < G; q++) A[q] = ps[(c + 2 * G + q) * 32 + threadIdx.x];
    }
#pragma unroll
    for (int q = 0; q < G; q++) { a = __fadd_rn(a, B[q].x); a = __fadd_rn(a, B[q].y); a = __fadd_rn(a, B[q].z); a = __fadd_rn(a, B[q].w); }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class K> void run(const char* name, K k, int n, size_t smem, int per) {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h = 0;
  for (int r = 0; r < 2; r++) { k<<<1, 32, smem>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); }
  printf("%-34s %.2f cyc/k  (%s)\n", name, (double)h / ((double)n * per), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  int nch = 256;
  run("P<2> rolled 2-chunk groups", chainP<2>, nch, nch * 512 + nch * 32, 8);
  run("P<4> rolled 4-chunk groups", chainP<4>, nch, nch * 512 + nch * 32, 8);
  run("P<8> rolled 8-chunk groups", chainP<8>, nch, nch * 512 + nch * 32, 8);
  run("Q<4> x as bf16 (1 LDS), 4-chunk", chainQ<4>, nch, nch * 512 + nch * 16, 8);
  run("Q<8> x as bf16 (1 LDS), 8-chunk", chainQ<8>, nch, nch * 512 + nch * 16, 8);
  run("R<4> FADD products, 4 float4", chainR<4>, 256, 256 * 512, 4);
  run("R<8> FADD products, 8 float4", chainR<8>, 256, 256 * 512, 4);
  return 0;
}
