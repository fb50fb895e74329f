// oracle/shim/cutil_subset.h -- TEST INFRASTRUCTURE ONLY (ours, not a copy).
//
// Stands in for the reference's include/cutil_subset.h when ITS CUDA kernels are built for the B200 baseline timing
// (oracle/Makefile `gpuref`, tools/ref_gpu.sh).  The reference's header defines
//     static unsigned CudaTest(const char *msg) { }          (include/cutil_subset.h:29-30)
// a value-returning function without a return statement: g++ 13 ends it in a trap (ud2 at -O0, fall-through into the
// next function at -O2), so every stock GPU driver dies right after its "Launching CUDA ... solver" line.  This header
// provides the three things the reference's pr/warp.cu, spmv/warp.cu and bfs/linear_lb.cu use from it -- the error-check
// macro, the shuffle aliases and a CudaTest that returns -- and nothing else; the kernels and their drivers are compiled
// from the reference's own sources, untouched.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// Reports a failed launch instead of ignoring it; returns 0 when the launch went through.
static inline unsigned gdn_shim_cuda_test(const char *msg) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::fprintf(stderr, "%s: %s\n", msg, cudaGetErrorString(e));
    std::exit(EXIT_FAILURE);
  }
  return 0;
}

// bfs/linear_lb.cu reaches the reference's header first through include/worklistc.h (a sibling include, which no -I order
// can redirect): its macros are then in place already, and only the CALLS of CudaTest are redirected below.
#ifndef CUDA_SAFE_CALL_NO_SYNC
#define CUDA_SAFE_CALL(call)                                                                                         \
  do {                                                                                                               \
    const cudaError_t gdn_shim_err = (call);                                                                         \
    if (gdn_shim_err != cudaSuccess) {                                                                               \
      std::fprintf(stderr, "CUDA error %d (%s) at %s:%d\n", (int)gdn_shim_err, cudaGetErrorString(gdn_shim_err),      \
                   __FILE__, __LINE__);                                                                              \
      std::exit(EXIT_FAILURE);                                                                                       \
    }                                                                                                                \
  } while (0)
#define CUDA_SAFE_CALL_NO_SYNC(call) CUDA_SAFE_CALL(call)
#define SHFL_DOWN(a, b) __shfl_down_sync(0xFFFFFFFF, a, b)
#define SHFL(a, b) __shfl_sync(0xFFFFFFFF, a, b)
#endif

#define CudaTest(msg) gdn_shim_cuda_test(msg)
