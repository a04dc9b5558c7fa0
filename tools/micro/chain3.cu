// can __syncwarp() pin the issue order of shared-memory loads ahead of a dependent FFMA chain?
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
#define LOADG(W, X0, X1, c0)                                                                  \
  _Pragma("unroll") for (int q = 0; q < G; q++) { int cc = (c0) + q; W[q] = ws[cc * 32 + threadIdx.x];  \
    X0[q] = *(const float4*)(xs + cc * 8); X1[q] = *(const float4*)(xs + cc * 8 + 4); }
#define FMAG(W, X0, X1) _Pragma("unroll") for (int q = 0; q < G; q++) { FMA8(a, W[q], X0[q], X1[q]) }

template <int G, int SYNC, int SETS>
__global__ void chain(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* ws = (uint4*)sm;
  float* xs = (float*)(sm + (size_t)nchunks * 512);
  for (int i = threadIdx.x; i < nchunks * 32; i += blockDim.x) ws[i] = make_uint4(0x3f803f80u, 0x3f003f00u, 0x3e803e80u, 0x3f803f80u);
  for (int i = threadIdx.x; i < nchunks * 8; i += blockDim.x) xs[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float a = 0.f;
  uint4 wA[G], wB[G], wC[G];
  float4 xA0[G], xA1[G], xB0[G], xB1[G], xC0[G], xC1[G];
  long long t0 = clock64();
  if (SETS == 2) {
    LOADG(wA, xA0, xA1, 0)
#pragma unroll 1
    for (int c = 0; c < nchunks; c += 2 * G) {
      LOADG(wB, xB0, xB1, c + G)
      if (SYNC) __syncwarp();
      FMAG(wA, xA0, xA1)
      if (c + 2 * G < nchunks) { LOADG(wA, xA0, xA1, c + 2 * G) }
      if (SYNC) __syncwarp();
      FMAG(wB, xB0, xB1)
    }
  } else {
    LOADG(wA, xA0, xA1, 0)
    LOADG(wB, xB0, xB1, G)
#pragma unroll 1
    for (int c = 0; c < nchunks; c += 3 * G) {
      if (c + 2 * G < nchunks) { LOADG(wC, xC0, xC1, c + 2 * G) }
      if (SYNC) __syncwarp();
      FMAG(wA, xA0, xA1)
      if (c + 3 * G < nchunks) { LOADG(wA, xA0, xA1, c + 3 * G) }
      if (SYNC) __syncwarp();
      if (c + G < nchunks) { FMAG(wB, xB0, xB1) }
      if (c + 4 * G < nchunks) { LOADG(wB, xB0, xB1, c + 4 * G) }
      if (SYNC) __syncwarp();
      if (c + 2 * G < nchunks) { FMAG(wC, xC0, xC1) }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class K> void run(const char* name, K k, int n) {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h = 0;
  for (int r = 0; r < 2; r++) { k<<<1, 32, n * 512 + n * 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); }
  printf("%-40s %.2f cyc/k  (%s)\n", name, (double)h / ((double)n * 8), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int n = 240;
  run("2 sets G=4 no sync", chain<4, 0, 2>, n);
  run("2 sets G=4 syncwarp", chain<4, 1, 2>, n);
  run("3 sets G=4 no sync", chain<4, 0, 3>, n);
  run("3 sets G=4 syncwarp", chain<4, 1, 3>, n);
  run("2 sets G=8 syncwarp", chain<8, 1, 2>, n);
  run("3 sets G=2 syncwarp", chain<2, 1, 3>, n);
  run("3 sets G=8 syncwarp", chain<8, 1, 3>, n);
  return 0;
}
