// gather_microbench.cu -- calibrates the design assumptions of DESIGN.md on a B200:
//   (1) streaming read bandwidth with 256-bit L1-bypassing loads
//   (2) random 4-byte gather rate out of a table of size T (L1 / L2 / HBM resident)
//   (3) the same gather out of shared memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_microbench gather_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void stream_read(const int4 *__restrict__ a, size_t n8, int *sink) {
  int acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    int q[8];
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                 : "l"(a + 2 * i));
    acc += q[0] ^ q[1] ^ q[2] ^ q[3] ^ q[4] ^ q[5] ^ q[6] ^ q[7];
  }
  if (acc == 0x12345678) *sink = acc;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each thread does `per` gathers at hashed indices, 8 independent per step
__global__ void gather_global(const float *__restrict__ tab, uint32_t mask, int per, float *sink) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int k = 0; k < per; k += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = __ldg(tab + (hash32(t * 7919u + (uint32_t)(k + j) * 0x9e3779b9u) & mask));
#pragma unroll
    for (int j = 0; j < 8; j++) acc += v[j];
  }
  if (acc == 1.2345f) *sink = acc;
}

__global__ void gather_shared(const float *__restrict__ tab, int n_tab, int per, float *sink) {
  extern __shared__ float s[];
  for (int i = threadIdx.x; i < n_tab; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int k = 0; k < per; k += 8) {
#pragma unroll
    for (int j = 0; j < 8; j++) acc += s[hash32(t * 7919u + (uint32_t)(k + j) * 0x9e3779b9u) % (uint32_t)n_tab];
  }
  if (acc == 1.2345f) *sink = acc;
}

int main() {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int *sink; CK(cudaMalloc(&sink, 64));
  const size_t big = (size_t)4 << 30;
  void *buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    stream_read<<<148 * 16, 256>>>((const int4 *)buf, big / 32, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("stream_read 4 GiB v8 loads: %.3f ms  %.1f GB/s\n", ms, big / ms / 1e6);
  }
  const int per = 1024;
  for (int lg = 14; lg <= 29; lg++) {     // table of 2^lg floats: 64 KB .. 2 GB
    uint32_t mask = (1u << lg) - 1;
    const int blocks = 148 * 8, threads = 256;
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaEventRecord(e0));
      gather_global<<<blocks, threads>>>((const float *)buf, mask, per, (float *)sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    double n = (double)blocks * threads * per;
    printf("gather_global table %8.2f MB: %.3f ms  %.1f Ggather/s  (%.2f gathers/clk/SM @1.9GHz)\n",
           (double)(4ull << lg) / 1e6, ms, n / ms / 1e6, n / ms / 1e6 / 148 / 1.9);
  }
  for (int kb : {16, 64, 128, 200}) {
    int n_tab = kb * 1024 / 4;
    CK(cudaFuncSetAttribute(gather_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
    const int blocks = 148 * 4, threads = 1024;
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaEventRecord(e0));
      gather_shared<<<blocks, threads, kb * 1024>>>((const float *)buf, n_tab, per, (float *)sink);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    double n = (double)blocks * threads * per;
    printf("gather_shared table %4d KB: %.3f ms  %.1f Ggather/s  (%.2f gathers/clk/SM @1.9GHz)\n", kb, ms, n / ms / 1e6,
           n / ms / 1e6 / 148 / 1.9);
  }
  return 0;
}
