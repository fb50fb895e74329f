// ordered_sum_microbench.cu -- cycles per 512-column block of gdn::ordered_row_sum (csrc/ordered_sum.cuh) for 1 / 8 warps
// of an otherwise idle SM, on a PageRank-like row (addends ~ 5e-10, 1 M columns), checked against the sequential sum.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gardenia_b200/csrc -o tools/ordered_sum_microbench tools/ordered_sum_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ordered_sum.cuh"
#undef cudaMalloc      // (common.cuh routes the library's allocations through its arena)
#undef cudaFree
using namespace gdn;
__global__ void run(const float4 *vals, uint32_t ngl, float *out, long long *cyc) {
  extern __shared__ float4 ring[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  const float s = ordered_row_sum(vals + wib, ngl, ring + (size_t)wib * kOrdDepth * 32 * 4, lane);
  const long long t1 = clock64();
  if (lane == 0) { out[wib] = s; cyc[wib] = t1 - t0; }
}
int main() {
  const uint32_t ngl = 250000;                          // groups per lane: 1 M columns
  std::vector<float> h((size_t)ngl * 32 * 4);
  srand(5);
  for (size_t i = 0; i < h.size(); i++) h[i] = 5e-10f * (0.5f + (float)rand() / RAND_MAX);
  float4 *d; float *out; long long *cyc;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 128); cudaMalloc(&cyc, 256);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)8 * kOrdDepth * 32 * 4 * sizeof(float4);
  cudaFuncSetAttribute(run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int warps : {1, 8}) {
    for (int rep = 0; rep < 2; rep++) run<<<1, warps * 32, smem>>>(d, ngl, out, cyc);
    float ho[8]; long long hc[8];
    cudaMemcpy(ho, out, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
    // sequential reference for row 0: element (q, c) of row r at h[((q * 32) + r) * 4 + c]
    volatile float seq = 0.f;
    for (uint32_t q = 0; q < ngl; q++) for (int c = 0; c < 4; c++) { volatile float t = seq + h[((size_t)q * 32) * 4 + c]; seq = t; }
    printf("%d warp(s): %.0f cycles per 512-column block (%s); row 0 sum %.9g vs sequential %.9g\n", warps, (double)hc[0] / ((ngl + 127) / 128),
           cudaGetErrorString(cudaGetLastError()), ho[0], (float)seq);
  }
  return 0;
}
