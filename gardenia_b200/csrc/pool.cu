// pool.cu -- bounded device arena behind the one-shot entry points.
//
// A one-shot PRSolver call at Kronecker scale 26 allocates ~18 GB in two dozen cudaMalloc calls and frees them at the
// end; the cudaFree calls (each a device synchronisation + unmapping) cost 15 ms per call and 80 ms in two of five
// (bench.py e2e, profiles/r1_bench_kron26_v5.json: 257 ms vs 325/340 ms per call with identical upload / solve /
// download times).  While a one-shot call is running (PoolScope in oneshot.cu) all blocks are tracked; freeing
// one parks it in the arena instead, and the next call's allocation of the SAME size on the SAME device takes it back.
// Repeated calls on the same graph -- what a caller of the reference's solver loop does -- then allocate nothing.
// What a caller sharing the device has to know (include/gdn_b200.h): memory parked by the last one-shot call stays
// allocated until gdn_device_trim() / gdn_finalize(), a failing allocation inside the library, or the NEXT one-shot call
// -- blocks that the call did not take back (a different graph) are freed when it returns, so the arena never holds more
// than one call's worth.  Bounded by GDN_DEVICE_ARENA_GB (default 64); GDN_DEVICE_ARENA=0 turns it off (every block is
// freed inside the call, the reference's contract, src/pr/base.cu:133-139).
// This file must NOT include common.cuh: that header maps cudaMalloc / cudaFree onto the two functions defined here.
#include <cuda_runtime.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

namespace gdn {

namespace {
struct Block { void *p; unsigned gen; };
std::mutex mu;
int depth = 0;                                              // nesting of PoolScope (all host threads)
unsigned generation = 0;                                    // one-shot calls so far
std::unordered_map<void *, std::pair<size_t, int>> live;    // blocks handed out while the arena was on: size, device
std::multimap<std::pair<int, size_t>, Block> idle;          // parked blocks by (device, size)
size_t idle_bytes = 0;
constexpr size_t kMinBlock = 1;                             // every block: a REAL cudaFree of even a small one stalled 150-380 ms in
                                                            // one call of three (GDN_TRACE, "buffers released"), so none is left

bool enabled_by_env() {
  static const bool on = [] { const char *e = getenv("GDN_DEVICE_ARENA"); return !(e && atoi(e) == 0); }();
  return on;
}
size_t cap_bytes() {
  static const size_t cap = [] { const char *e = getenv("GDN_DEVICE_ARENA_GB"); return (size_t)(e ? atoi(e) : 64) << 30; }();
  return cap;
}
int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
// blocks must be freed on their own device
void free_blocks(std::vector<std::pair<int, void *>> &v) {
  if (v.empty()) return;
  const int here = current_device();
  for (auto &b : v) { cudaSetDevice(b.first); cudaFree(b.second); }
  cudaSetDevice(here);
}
}  // namespace

void pool_release() {
  std::vector<std::pair<int, void *>> take;
  {
    std::lock_guard<std::mutex> lk(mu);
    for (auto &kv : idle) take.emplace_back(kv.first.first, kv.second.p);
    idle.clear();
    idle_bytes = 0;
  }
  free_blocks(take);
}

void pool_scope(bool enter) {
  std::vector<std::pair<int, void *>> stale;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (enter) {
      if (depth++ == 0) generation++;
    } else {
      if (--depth < 0) depth = 0;
      if (depth == 0) {
        // what this call did not take back (parked by an earlier call on another graph) goes now
        for (auto it = idle.begin(); it != idle.end();) {
          if (it->second.gen != generation) {
            stale.emplace_back(it->first.first, it->second.p);
            idle_bytes -= it->first.second;
            it = idle.erase(it);
          } else {
            ++it;
          }
        }
      }
    }
  }
  free_blocks(stale);
}

cudaError_t pool_malloc(void **p, size_t bytes) {
  const bool track = bytes >= kMinBlock && enabled_by_env();
  const int dev = track ? current_device() : 0;
  bool have_idle = false;
  if (track) {
    std::lock_guard<std::mutex> lk(mu);
    if (depth > 0) {
      auto it = idle.find(std::make_pair(dev, bytes));
      if (it != idle.end()) {
        *p = it->second.p;
        idle.erase(it);
        idle_bytes -= bytes;
        live[*p] = std::make_pair(bytes, dev);
        return cudaSuccess;
      }
    }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    { std::lock_guard<std::mutex> lk(mu); have_idle = idle_bytes > 0; }
    if (have_idle) {                                          // the arena may be what is in the way
      cudaGetLastError();
      pool_release();
      e = cudaMalloc(p, bytes);
    }
  }
  if (e == cudaSuccess && track) {
    std::lock_guard<std::mutex> lk(mu);
    if (depth > 0) live[*p] = std::make_pair(bytes, dev);
  }
  return e;
}

cudaError_t pool_free(void *p) {
  if (!p) return cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it != live.end()) {
      const size_t n = it->second.first;
      const int dev = it->second.second;
      live.erase(it);
      if (depth > 0 && idle_bytes + n <= cap_bytes()) {
        idle.emplace(std::make_pair(dev, n), Block{p, generation});
        idle_bytes += n;
        return cudaSuccess;
      }
    }
  }
  return cudaFree(p);
}

}  // namespace gdn
