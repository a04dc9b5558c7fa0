// micro-benchmarks for the dependent-FFMA accumulation chain of LNB_ACC_STRICT
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0u, 0x1044)); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__global__ void chainA(float* out, long long* cyc, float m, int n) {
  float a = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) a = __fmaf_rn(a, m, 1.0f);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// accumulator chain: acc = x*w + acc, x,w from registers (independent of acc)
__global__ void chainB(float* out, long long* cyc, int n) {
  float a = 0.f, x = out[threadIdx.x], w = out[threadIdx.x + 32];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; i++) a = __fmaf_rn(x, w, a);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the real inner loop: w chunk (uint4) + x (2 float4) from shared memory, 8 FMAs per chunk
__global__ void chainC(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint4* ws = (uint4*)sm;                    // [nchunks][32 lanes]
  float* xs = (float*)(sm + (size_t)nchunks * 512);  // [nchunks*8]
  for (int i = threadIdx.x; i < nchunks * 32; i += blockDim.x) ws[i] = make_uint4(0x3f803f80u, 0x3f003f00u, 0x3e803e80u, 0x3f803f80u);
  for (int i = threadIdx.x; i < nchunks * 8; i += blockDim.x) xs[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float a = 0.f;
  long long t0 = clock64();
#pragma unroll 4
  for (int c = 0; c < nchunks; c++) {
    uint4 wv = ws[c * 32 + threadIdx.x];
    float4 xa = *(const float4*)(xs + c * 8), xb = *(const float4*)(xs + c * 8 + 4);
    a = __fmaf_rn(xa.x, bf_lo(wv.x), a); a = __fmaf_rn(xa.y, bf_hi(wv.x), a);
    a = __fmaf_rn(xa.z, bf_lo(wv.y), a); a = __fmaf_rn(xa.w, bf_hi(wv.y), a);
    a = __fmaf_rn(xb.x, bf_lo(wv.z), a); a = __fmaf_rn(xb.y, bf_hi(wv.z), a);
    a = __fmaf_rn(xb.z, bf_lo(wv.w), a); a = __fmaf_rn(xb.w, bf_hi(wv.w), a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// same, but weights pre-converted to f32 in shared memory (2 x float4 per chunk), no cvt in the chain warp
__global__ void chainD(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) uint8_t sm[];
  float4* ws = (float4*)sm;                  // [nchunks][2][32]
  float* xs = (float*)(sm + (size_t)nchunks * 1024);
  for (int i = threadIdx.x; i < nchunks * 64; i += blockDim.x) ws[i] = make_float4(1.f, .5f, .25f, 1.f);
  for (int i = threadIdx.x; i < nchunks * 8; i += blockDim.x) xs[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float a = 0.f;
  long long t0 = clock64();
#pragma unroll 4
  for (int c = 0; c < nchunks; c++) {
    float4 w0 = ws[(c * 2) * 32 + threadIdx.x], w1 = ws[(c * 2 + 1) * 32 + threadIdx.x];
    float4 xa = *(const float4*)(xs + c * 8), xb = *(const float4*)(xs + c * 8 + 4);
    a = __fmaf_rn(xa.x, w0.x, a); a = __fmaf_rn(xa.y, w0.y, a); a = __fmaf_rn(xa.z, w0.z, a); a = __fmaf_rn(xa.w, w0.w, a);
    a = __fmaf_rn(xb.x, w1.x, a); a = __fmaf_rn(xb.y, w1.y, a); a = __fmaf_rn(xb.z, w1.z, a); a = __fmaf_rn(xb.w, w1.w, a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// products precomputed: chain does only FADD of float4s from shared memory
__global__ void chainE(float* out, long long* cyc, int n4) {
  extern __shared__ __align__(16) uint8_t sm[];
  float4* ps = (float4*)sm;                  // [n4][32]
  for (int i = threadIdx.x; i < n4 * 32; i += blockDim.x) ps[i] = make_float4(1.f, .5f, .25f, 1.f);
  __syncthreads();
  float a = 0.f;
  long long t0 = clock64();
#pragma unroll 8
  for (int c = 0; c < n4; c++) {
    float4 p = ps[c * 32 + threadIdx.x];
    a = __fadd_rn(a, p.x); a = __fadd_rn(a, p.y); a = __fadd_rn(a, p.z); a = __fadd_rn(a, p.w);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 4096);
  long long h;
  int n = 4096;
  for (int rep = 0; rep < 2; rep++) {
    chainA<<<1, 32>>>(out, cyc, 0.999f, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("A dependent FFMA (a=a*m+1):      %.2f cyc/op\n", (double)h / n);
    chainB<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("B accumulate FFMA (a=x*w+a):      %.2f cyc/op\n", (double)h / n);
    int nch = 512;
    cudaFuncSetAttribute(chainC, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(chainD, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(chainE, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    nch = 256;
    chainC<<<1, 32, nch * 512 + nch * 32>>>(out, cyc, nch); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("C smem bf16 w + smem x, 1 warp:    %.2f cyc/k\n", (double)h / (nch * 8));
    nch = 128;
    chainD<<<1, 32, nch * 1024 + nch * 32>>>(out, cyc, nch); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("D smem f32 w + smem x, 1 warp:     %.2f cyc/k\n", (double)h / (nch * 8));
    int n4 = 256;
    chainE<<<1, 32, n4 * 512>>>(out, cyc, n4); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("E smem products, FADD chain:      %.2f cyc/k\n", (double)h / (n4 * 4));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
