// fadd_chain_microbench.cu -- cycles per dependent fp32 add on one warp of an otherwise idle SM (sm_100a):
// FADD chain, FFMA-by-one chain (same bits), and the same with the interleaved LDS.128 of the PageRank chain warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fadd_chain_microbench tools/fadd_chain_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void chain(float *out, long long *cyc, int n, float one, const float4 *src) {
  extern __shared__ float4 tile[];
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) tile[i] = src[i];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < n; it++) {
    const float4 *t = tile + lane;
#pragma unroll 4
    for (int q = 0; q < 64; q++) {
      const float4 v = t[q * 32];
      if (MODE == 0) { acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w); }
      else { acc = __fmaf_rn(v.x, one, acc); acc = __fmaf_rn(v.y, one, acc); acc = __fmaf_rn(v.z, one, acc); acc = __fmaf_rn(v.w, one, acc); }
    }
  }
  const long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) *cyc = t1 - t0;
}
int main() {
  float *out; long long *cyc; float4 *src;
  cudaMalloc(&out, 128); cudaMalloc(&cyc, 8); cudaMalloc(&src, 64 * 32 * 16);
  cudaMemset(src, 0, 64 * 32 * 16);
  const int n = 4096;              // 4096 x 256 columns = 1 M dependent adds
  for (int threads : {32, 1024}) {
    for (int mode = 0; mode < 2; mode++) {
      long long h = 0;
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) chain<0><<<1, threads, 64 * 32 * 16>>>(out, cyc, n, 1.0f, src);
        else chain<1><<<1, threads, 64 * 32 * 16>>>(out, cyc, n, 1.0f, src);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("%s, %4d threads in the CTA: %.2f cycles per dependent add\n", mode ? "FFMA x*1+acc" : "FADD        ", threads, (double)h / (n * 256.0));
    }
  }
  return 0;
}
