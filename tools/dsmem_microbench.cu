// dsmem_microbench.cu -- random 4-byte gather out of a table spread over the shared
// memories of a thread-block cluster (distributed shared memory), B200.
// Question for DESIGN.md: can the PageRank hot-vertex table grow from one SM's
// 192 KB to C x 192 KB without falling back to the L2 gather rate (~1/clk/SM)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_microbench dsmem_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__device__ __forceinline__ float ld_cluster(uint32_t saddr_local, uint32_t rank) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(saddr_local), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
  return v;
}

// slice = floats per CTA (power of two), csize = cluster size (power of two).
// local_frac_256: of 256, how many gathers are forced to the local slice (models a skewed hot table)
__global__ void gather_dsmem(const float *__restrict__ tab, int slice_lg, int csize_lg, int per, int local_256, float *sink) {
  extern __shared__ float s[];
  cg::cluster_group cl = cg::this_cluster();
  const uint32_t slice = 1u << slice_lg;
  const uint32_t my = cl.block_rank();
  for (uint32_t i = threadIdx.x; i < slice; i += blockDim.x) s[i] = tab[my * slice + i];
  cl.sync();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s);
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int k = 0; k < per; k += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t h = hash32(t * 7919u + (uint32_t)(k + j) * 0x9e3779b9u);
      const uint32_t off = h & (slice - 1);
      uint32_t rank = (h >> slice_lg) & ((1u << csize_lg) - 1);
      if ((int)(h >> 24) < local_256) rank = my;
      v[j] = ld_cluster(sbase + off * 4, rank);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc += v[j];
  }
  if (acc == 1.2345f) *sink = acc;
  cl.sync();
}

int main() {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float *sink; CK(cudaMalloc(&sink, 64));
  float *buf; CK(cudaMalloc(&buf, 64 << 20)); CK(cudaMemset(buf, 0, 64 << 20));
  CK(cudaFuncSetAttribute(gather_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int per = 512;
  for (int kb : {64, 128}) {
    int slice_lg = (kb == 64) ? 14 : 15;
    CK(cudaFuncSetAttribute(gather_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
    for (int threads : {512, 1024}) {
      for (int clg = 0; clg <= 4; clg++) {
        for (int local_256 : {0, 128, 192}) {
          if (clg == 0 && local_256) continue;
          const int csize = 1 << clg;
          cudaLaunchConfig_t cfg = {};
          int blocks = 148 / csize * csize;
          cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = kb * 1024;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          float ms = 0;
          cudaError_t err = cudaSuccess;
          for (int rep = 0; rep < 2 && err == cudaSuccess; rep++) {
            CK(cudaEventRecord(e0));
            err = cudaLaunchKernelEx(&cfg, gather_dsmem, (const float *)buf, slice_lg, clg, per, local_256, sink);
            CK(cudaEventRecord(e1));
            if (err == cudaSuccess) err = cudaEventSynchronize(e1);
            if (err == cudaSuccess) CK(cudaEventElapsedTime(&ms, e0, e1));
          }
          if (err != cudaSuccess) { printf("cluster %2d slice %3d KB threads %4d: launch failed: %s\n", csize, kb, threads, cudaGetErrorString(err)); cudaGetLastError(); continue; }
          double n = (double)blocks * threads * per;
          printf("gather_dsmem cluster %2d x %3d KB (table %5d KB) threads %4d forced-local %3d/256: %.3f ms  %.1f Ggather/s  (%.2f gathers/clk/SM @1.9GHz, %d CTAs)\n",
                 csize, kb, csize * kb, threads, local_256, ms, n / ms / 1e6, n / ms / 1e6 / blocks / 1.9, blocks);
        }
      }
    }
  }
  return 0;
}
