// l2_policy_microbench.cu -- how much of a gathered table does B200's L2 keep while a
// one-pass index stream flows through it?  Models the pull gather of PageRank / SpMV:
// every warp streams 256-bit chunks of an index array (4 GiB, one pass) and gathers
// table[index] (4-byte random reads, table T MB).  Variants:
//   stream policy  : evict_first + L1 no_allocate   |  plain
//   gather policy  : plain ld.global.nc | L2::evict_last | L2::evict_normal hint policy
//   persisting L2  : cudaAccessPolicyWindow over the table (set-aside = device max) on/off
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_policy_microbench l2_policy_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void fill_idx(uint32_t *idx, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) idx[i] = hash32((uint32_t)i * 2654435761u + 12345u);
}

template <int SP, int GP>
__global__ void __launch_bounds__(256, 4) stream_gather(const uint32_t *__restrict__ idx, size_t n8, const float *__restrict__ tab, uint32_t mask, float *sink) {
  float acc = 0.f;
  uint64_t pol;
  if (GP == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t q[8];
    if (SP == 0)
      asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "l"(idx + 8 * i));
    else
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "l"(idx + 8 * i));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float *p = tab + (q[j] & mask);
      if (GP == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v[j]) : "l"(p));
      else asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[j]) : "l"(p), "l"(pol));
    }
#pragma unroll
    for (int j = 0; j < 8; j++) acc += v[j];
  }
  if (acc == 1.2345f) *sink = acc;
}

typedef void (*kern_t)(const uint32_t *, size_t, const float *, uint32_t, float *);

int main() {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  float *sink; CK(cudaMalloc(&sink, 64));
  const size_t n = (size_t)1 << 30;                 // 1 Gi indices = 4 GiB stream
  uint32_t *idx; CK(cudaMalloc(&idx, n * 4));
  fill_idx<<<148 * 8, 256, 0, st>>>(idx, n);
  float *tab; CK(cudaMalloc(&tab, (size_t)512 << 20)); CK(cudaMemsetAsync(tab, 0, (size_t)512 << 20, st));
  CK(cudaStreamSynchronize(st));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20, p.accessPolicyMaxWindowSize >> 20);
  kern_t kerns[2][3] = {{stream_gather<0, 0>, stream_gather<0, 1>, stream_gather<0, 2>}, {stream_gather<1, 0>, stream_gather<1, 1>, stream_gather<1, 2>}};
  const char *spn[2] = {"stream evict_first", "stream plain      "}, *gpn[3] = {"gather plain ", "gather evict_last", "gather evict_normal"};
  for (int persist = 0; persist < 2; persist++) {
    for (int lg = 22; lg <= 25; lg++) {              // 16, 32, 64, 128 MB tables (+ 96 MB via mask trick below)
      for (int extra = 0; extra < 2; extra++) {
        if (extra && lg != 24) continue;
        // extra: 96 MB table = indices masked to 2^25 then folded
        size_t tbytes = extra ? ((size_t)96 << 20) : ((size_t)4 << lg);
        uint32_t mask = extra ? 0 : (1u << lg) - 1;
        if (extra) continue;                          // keep the run short: power-of-two tables only
        if (persist) {
          CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, p.persistingL2CacheMaxSize));
          cudaStreamAttrValue av = {};
          av.accessPolicyWindow.base_ptr = tab;
          av.accessPolicyWindow.num_bytes = tbytes < (size_t)p.accessPolicyMaxWindowSize ? tbytes : (size_t)p.accessPolicyMaxWindowSize;
          av.accessPolicyWindow.hitRatio = 1.0f;
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        for (int sp = 0; sp < 2; sp++)
          for (int gp = 0; gp < 3; gp++) {
            if (persist && gp != 0) continue;
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
              CK(cudaEventRecord(e0, st));
              kerns[sp][gp]<<<148 * 8, 256, 0, st>>>(idx, n / 8, tab, mask, sink);
              CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            printf("table %4zu MB  %s  %s  persistWindow %d : %.3f ms  %.1f Ggather/s  stream %.2f TB/s\n", tbytes >> 20, spn[sp], gpn[gp], persist, ms, n / ms / 1e6, n * 4.0 / ms / 1e9);
          }
        if (persist) {
          cudaStreamAttrValue av = {};
          av.accessPolicyWindow.num_bytes = 0;
          CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
          CK(cudaCtxResetPersistingL2Cache());
        }
      }
    }
  }
  return 0;
}
