// pcie_bw.cu -- host->device / device->host copy bandwidth of this box, for the e2e budget of the one-shot entry points
// (DESIGN.md 6): cudaHostAlloc'ed vs cudaHostRegister'ed (the caller's CSR arrays are registered, include/gdn_b200.h gdn_host_pin)
// vs pageable memory, one and two streams.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_bw pcie_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double run(void *d, void *h, size_t n, bool h2d, int streams, cudaStream_t *st) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaDeviceSynchronize();
  cudaEventRecord(a, st[0]);
  for (int s = 0; s < streams; s++) {
    size_t off = n / streams * s, len = n / streams;
    if (s) cudaStreamWaitEvent(st[s], a, 0);
    if (h2d) cudaMemcpyAsync((char *)d + off, (char *)h + off, len, cudaMemcpyHostToDevice, st[s]);
    else cudaMemcpyAsync((char *)h + off, (char *)d + off, len, cudaMemcpyDeviceToHost, st[s]);
  }
  cudaEvent_t e1; cudaEventCreate(&e1);
  for (int s = 1; s < streams; s++) { cudaEventRecord(e1, st[s]); cudaStreamWaitEvent(st[0], e1, 0); }
  cudaEventRecord(b, st[0]);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return n / (ms * 1e-3) / 1e9;
}
int main() {
  const size_t n = (size_t)2 << 30;
  void *d; CK(cudaMalloc(&d, n));
  cudaStream_t st[2]; cudaStreamCreate(&st[0]); cudaStreamCreate(&st[1]);
  void *hp; CK(cudaHostAlloc(&hp, n, cudaHostAllocDefault)); memset(hp, 1, n);
  void *hr = aligned_alloc(4096, n); memset(hr, 1, n); CK(cudaHostRegister(hr, n, cudaHostRegisterDefault));
  void *hg = aligned_alloc(4096, n); memset(hg, 1, n);
  for (int rep = 0; rep < 2; rep++) {
    printf("H2D hostalloc   1 stream  %.1f GB/s\n", run(d, hp, n, true, 1, st));
    printf("H2D hostalloc   2 streams %.1f GB/s\n", run(d, hp, n, true, 2, st));
    printf("H2D registered  1 stream  %.1f GB/s\n", run(d, hr, n, true, 1, st));
    printf("H2D pageable    1 stream  %.1f GB/s\n", run(d, hg, n, true, 1, st));
    printf("D2H hostalloc   1 stream  %.1f GB/s\n", run(d, hp, n, false, 1, st));
    printf("D2H registered  1 stream  %.1f GB/s\n", run(d, hr, n, false, 1, st));
  }
  return 0;
}
