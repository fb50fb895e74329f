// pool.cu -- grow-only device arena behind the one-shot entry points.
//
// A one-shot PRSolver call at Kronecker scale 26 allocates ~18 GB in two dozen cudaMalloc calls and frees them at the
// end; the cudaFree calls (each a device synchronisation + unmapping) cost 15 ms per call and 80 ms in two of five
// (bench.py e2e, profiles/r1_bench_kron26_v5.json: 257 ms vs 325/340 ms per call with identical upload / solve /
// download times).  While a one-shot call is running (PoolScope in oneshot.cu) blocks of >= 1 MB are tracked; freeing
// one parks it in the arena instead, and the next call's allocation of the SAME size takes it back.  Repeated calls on
// the same graph -- what a caller of the reference's solver loop does -- then allocate nothing.  The arena is bounded
// (GDN_DEVICE_ARENA_GB, default 64), emptied on allocation failure and by gdn_finalize; GDN_DEVICE_ARENA=0 turns it off.
// This file must NOT include common.cuh: that header maps cudaMalloc / cudaFree onto the two functions defined here.
#include <cuda_runtime.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>

namespace gdn {

namespace {
std::mutex mu;
int depth = 0;                                        // nesting of PoolScope
std::unordered_map<void *, size_t> live;              // blocks handed out while the arena was on
std::multimap<size_t, void *> idle;                   // parked blocks by size
size_t idle_bytes = 0;
constexpr size_t kMinBlock = (size_t)1 << 20;

bool enabled_by_env() {
  static const bool on = [] { const char *e = getenv("GDN_DEVICE_ARENA"); return !(e && atoi(e) == 0); }();
  return on;
}
size_t cap_bytes() {
  static const size_t cap = [] { const char *e = getenv("GDN_DEVICE_ARENA_GB"); return (size_t)(e ? atoi(e) : 64) << 30; }();
  return cap;
}
}  // namespace

void pool_release() {
  std::multimap<size_t, void *> take;
  {
    std::lock_guard<std::mutex> lk(mu);
    take.swap(idle);
    idle_bytes = 0;
  }
  for (auto &kv : take) cudaFree(kv.second);
}

void pool_scope(bool enter) {
  std::lock_guard<std::mutex> lk(mu);
  depth += enter ? 1 : -1;
  if (depth < 0) depth = 0;
}

cudaError_t pool_malloc(void **p, size_t bytes) {
  const bool track = bytes >= kMinBlock && enabled_by_env();
  if (track) {
    std::lock_guard<std::mutex> lk(mu);
    if (depth > 0) {
      auto it = idle.find(bytes);
      if (it != idle.end()) {
        *p = it->second;
        idle.erase(it);
        idle_bytes -= bytes;
        live[*p] = bytes;
        return cudaSuccess;
      }
    }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation && idle_bytes > 0) {       // the arena may be what is in the way
    cudaGetLastError();
    pool_release();
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess && track) {
    std::lock_guard<std::mutex> lk(mu);
    if (depth > 0) live[*p] = bytes;
  }
  return e;
}

cudaError_t pool_free(void *p) {
  if (!p) return cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it != live.end()) {
      const size_t n = it->second;
      live.erase(it);
      if (depth > 0 && idle_bytes + n <= cap_bytes()) {
        idle.emplace(n, p);
        idle_bytes += n;
        return cudaSuccess;
      }
    }
  }
  return cudaFree(p);
}

}  // namespace gdn
