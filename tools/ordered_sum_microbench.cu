// ordered_sum_microbench.cu -- gdn::ordered_block (csrc/ordered_sum.cuh), the exact per-block emulation of a sequential
// fp32 sum, on one warp of an idle SM: cycles per 512-column block and the result against the true sequential sum
// (PageRank-like row: 1 M addends around 5e-10).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gardenia_b200/csrc -o tools/ordered_sum_microbench tools/ordered_sum_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ordered_sum.cuh"
#undef cudaMalloc      // (common.cuh routes the library's allocations through its arena)
#undef cudaFree
using namespace gdn;
__global__ void run(const float4 *vals, uint32_t n_blocks, float *out, long long *cyc) {
  __shared__ float stage[512];
  const int lane = threadIdx.x & 31;
  uint32_t ab = 0;
  const long long t0 = clock64();
  for (uint32_t b = 0; b < n_blocks; b++) {
    float v[16];
    for (int u = 0; u < 4; u++) {
      const float4 t = vals[(size_t)b * 128 + lane * 4 + u];
      v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
    }
    ab = ordered_block(ab, v, lane, stage);
  }
  const long long t1 = clock64();
  if (lane == 0) { *out = __uint_as_float(ab); *cyc = t1 - t0; }
}
int main() {
  const uint32_t n_blocks = 2048;                       // 1 M columns, row-major
  std::vector<float> h((size_t)n_blocks * 512);
  srand(5);
  for (size_t i = 0; i < h.size(); i++) h[i] = 5e-10f * (0.5f + (float)rand() / RAND_MAX);
  float4 *d; float *out; long long *cyc;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) run<<<1, 32>>>(d, n_blocks, out, cyc);
  float ho; long long hc;
  cudaMemcpy(&ho, out, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  volatile float seq = 0.f;
  for (size_t i = 0; i < h.size(); i++) { volatile float t = seq + h[i]; seq = t; }
  printf("ordered_block: %.0f cycles per 512-column block (%s); sum %.9g vs sequential %.9g (%s)\n", (double)hc / n_blocks,
         cudaGetErrorString(cudaGetLastError()), ho, (float)seq, ho == (float)seq ? "same bits" : "DIFFERENT");
  return 0;
}
