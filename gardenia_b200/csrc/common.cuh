// common.cuh -- shared device helpers and the library's internal state.
// sm_100a only (B200); no CUB/Thrust, no multi-arch dispatch.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/gdn_b200.h"

// Device allocations of the library go through the arena of pool.cu (a plain cudaMalloc / cudaFree outside one-shot calls).
namespace gdn {
cudaError_t pool_malloc(void **p, size_t bytes);
cudaError_t pool_free(void *p);
void pool_release();              // really free what the arena holds (gdn_finalize, allocation failure)
void pool_scope(bool enter);      // a one-shot call is running: park freed blocks instead of freeing them
struct PoolScope {
  PoolScope() { pool_scope(true); }
  ~PoolScope() { pool_scope(false); }
};
}  // namespace gdn
#define cudaMalloc(p, n) gdn::pool_malloc((void **)(p), (n))
#define cudaFree(p) gdn::pool_free((void *)(p))

namespace gdn {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
struct Lib {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // chunked host->device copies that kernels on `stream` consume chunk by chunk
  bool stream_fill = false;             // hint set by the one-shot PageRank entry point: build the pull layout WHILE the CSR uploads
  bool pr_exact = false;                // gdn_set_pr_exact_order: graphs created from now on sum every PageRank row in column order
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pr_ev[4] = {};            // PageRank: one per iteration in flight (the host runs ahead of the device, pull.cu)
  void *pinned = nullptr;          // small pinned mailbox for per-step counters
  size_t pinned_bytes = 0;
  // grow-only page-locked scratch of the layout preprocessing (degree array, row order): no page faults on reuse, and
  // the order goes to the device at full PCIe rate without a bounce buffer
  void *arena[2] = {nullptr, nullptr};
  size_t arena_bytes[2] = {0, 0};
  // event pairs bracketing each launch of the dominant kernel of a solve
  static constexpr int kMaxKev = 512;
  cudaEvent_t kev[2 * kMaxKev] = {};
  int n_kev = 0;
};
Lib &lib();
// usage: kev_begin(); kernel<<<>>>; kev_end();  ...  after the final sync: kev_collect(st)
inline void kev_reset() { lib().n_kev = 0; }
inline void kev_begin() { Lib &l = lib(); if (l.n_kev < Lib::kMaxKev) cudaEventRecord(l.kev[2 * l.n_kev], l.stream); }
inline void kev_end() { Lib &l = lib(); if (l.n_kev < Lib::kMaxKev) { cudaEventRecord(l.kev[2 * l.n_kev + 1], l.stream); l.n_kev++; } }
inline void kev_collect(gdn_stats *st, int n_valid = 1 << 30) {
  if (!st) return;
  Lib &l = lib();
  double tot = 0;
  if (l.n_kev > n_valid) l.n_kev = n_valid;
  for (int i = 0; i < l.n_kev; i++) { float ms = 0; if (cudaEventElapsedTime(&ms, l.kev[2 * i], l.kev[2 * i + 1]) == cudaSuccess) tot += ms; }
  st->kernel_ms = tot;
  st->kernel_calls = l.n_kev;
}
Lib &lib();                        // the calling thread's device context (a gang worker's own, else the process-wide one)
void lib_bind(Lib *l);             // graph.cu: make `l` the calling thread's context (nullptr: back to the process-wide one)
// Single-process multi-GPU mode (comm.cu): gdn_init_gpus(n) starts one worker thread per GPU; the one-shot entry points
// hand each of them a row partition.  Inside a worker comm_size() / comm_rank() are the gang's.
int gang_size();                   // 0 when the mode is off
bool gang_worker();                // is the calling thread a gang worker?
int gang_run(int (*fn)(int rank, void *arg), void *arg);       // run fn on every worker, wait; first non-zero status (message kept)
void gang_barrier();               // workers only
void *gang_shared(size_t bytes);   // workers only, collective: one page-locked host buffer of at least `bytes`, the same for all
void *host_arena(int slot, size_t bytes);      // graph.cu; nullptr when page-locking fails (callers fall back to malloc)
int ensure_init();
// GDN_TRACE=1: wall-clock stage timings of the upload / preprocessing path on stderr
void trace(const char *label);

#define GDN_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      gdn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return (e_ == cudaErrorMemoryAllocation) ? GDN_ERR_NOMEM : GDN_ERR_CUDA;          \
    }                                                                                   \
  } while (0)

#define GDN_CHECK(call)            \
  do {                             \
    int rc_ = (call);              \
    if (rc_ != GDN_OK) return rc_; \
  } while (0)

// ---------------------------------------------------------------- device-side CSR
// One direction of the graph as the kernels see it.  Offsets are LOCAL to this
// GPU's row partition (32-bit whenever the local nnz fits), column indices are
// GLOBAL vertex ids.  `col` is 256-B aligned and padded by >= 16 B so that
// 128-bit loads may over-read the tail.
struct DevCsr {
  int64_t rows = 0;        // local rows
  uint64_t nnz = 0;        // local non-zeros
  bool off64 = false;
  void *rowptr = nullptr;  // uint32_t[rows+1] or uint64_t[rows+1]
  int32_t *col = nullptr;
  // row-block schedule for the pull gather (gather.cu)
  int32_t n_chunks = 0;        // light blocks: rows [chunk_row[k], chunk_row[k+1]) minus a heavy last row
  int32_t *chunk_row = nullptr;    // int32[n_chunks+1]
  int32_t n_heavy_rows = 0;
  int32_t n_heavy_segs = 0;
  // the heavy rows of at most kSpmvExactLen entries come first in the lists below: SpMV sums those by segments and the
  // longer ones in the reference's order (gather.cu, SpmvExact); the PageRank fallback walks the whole list
  int32_t n_heavy_rows_lt = 0;
  int32_t n_heavy_segs_lt = 0;
  int32_t *heavy_row = nullptr;    // int32[n_heavy_rows]      local row id
  int32_t *heavy_first = nullptr;  // int32[n_heavy_rows+1]    first segment index of that row
  int2 *heavy_seg = nullptr;       // int2[n_heavy_segs]       {local row, segment number within the row}
  float *heavy_partial = nullptr;  // float[n_heavy_segs]
};

// Banded shared-memory layout of the heavy rows of the pull layout (band.cu): the column ids of a row that fall into
// band b = [b * band, (b + 1) * band) of the degree-sorted id space are stored as 16-bit band-local ids in band b's own
// SELL-32 array (rows of a band sorted by their count in it), and a CTA that holds contrib[band b] in shared memory
// sums them; what stays behind (cold ids, rows with too few entries in a band) forms a compacted main SELL array.
struct BandLayout {
  bool tried = false, built = false;
  bool seg = false;                // segmented mode: bands are L2-sized slices, 32-bit ids, one launch per band (pr_seg_kernel)
  int32_t B = 0, band = 0, cmin = 0, dmin = 0;   // bands, ids per band, min entries of a (row, band) pair, min row length
  int64_t n_rows = 0;              // sorted rows [0, n_rows) take part (whole slices)
  uint4 *bsell = nullptr;          // 8 band-local ids (uint16) per unit; item i = units [item_ptr[i], item_ptr[i+1]), lane-interleaved
  uint64_t n_units = 0;
  int64_t *band_start = nullptr;   // [B] first new id of band b (band.cu band_of: hot prefix, then the ranks' cold slices round-robin)
  int32_t *band_len = nullptr;     // [B] ids in band b
  int32_t n_items = 0;             // one item = up to kBandSeg index groups of one band slice (32 rows)
  uint32_t *item_ptr = nullptr;    // [n_items + 1]
  int32_t n_jobs = 0;              // a job = a run of items of ONE band given to one CTA, cut into 32 warp runs
  int4 *job = nullptr;             // [n_jobs] {band, first warp-run boundary index, 0, 0}
  int32_t *job_first = nullptr;    // [n_cta + 1] jobs of CTA c
  int32_t *wrun = nullptr;         // [n_jobs * 33] item boundaries of the warp runs
  int32_t n_cta = 0;
  int32_t *irow = nullptr;         // [n_items * 32] sorted row of (item, lane), -1 = none
  long long *acc_fix = nullptr;    // [n_rows] fixed-point sum of a row's band partials (scale chosen per solve, pull.cu fix_scale_for)
  double build_ms = 0;             // wall time of band_build (reported by gdn_graph_pull_info)
  float *acc_main = nullptr;       // [n_rows] sum over the columns left in the main array
  // the compacted main SELL array and its work tables (same meaning as the PullLayout fields)
  int4 *sell = nullptr;
  uint64_t n_groups = 0;
  uint32_t *slice_ptr = nullptr;
  int32_t n_chunks = 0;
  int32_t *chunk_slice = nullptr;
  int32_t n_heavy_slices = 0, n_heavy_segs = 0;
  int32_t *heavy_slice = nullptr, *heavy_first = nullptr;
  int2 *heavy_seg = nullptr;
  float *partial = nullptr;
  uint64_t moved = 0, pairs = 0;   // statistics: entries served from shared-memory bands, (row, band) pairs
};

// Degree-sorted SELL-32 layout of the pull (in-) CSR used by PageRank (pull.cu).
struct PullLayout {
  bool prepared = false;       // host part done (orders, ids, slice pointers, work items)
  bool exact = false;          // exact-order mode: no wide-slice segments, no bands (every row summed in column order by one lane)
  uint32_t group_ch = 1024;    // int4 groups per work item (pull.cuh kGroupCh; 0xffffffff in exact-order mode)
  int32_t n_exact = 0;         // leading slices whose rows keep the reference's summation order (never cut, never banded)
  bool symmetric_order = true; // row order == column order (row new-ids are two contiguous runs)
  int P = 1, R = 0;            // communicator size / rank the ids were laid out for
  int64_t W = 0, H = 0, Hp = 0, Wc = 0, Mp = 0;   // slice width, hot ids (total / per rank), cold width, id space
  int64_t rows = 0, n_nz_rows = 0;
  int32_t n_slices = 0;
  uint64_t n_groups = 0;       // int4 groups in the SELL array
  int32_t *perm = nullptr;     // [rows]  sorted position -> old local row
  int32_t *newid = nullptr;    // [m]     old global id -> new global id
  int32_t *sdeg = nullptr;     // [rows]  length of sorted row j
  int32_t *sout = nullptr;     // [rows]  out-degree of sorted row j (directed graphs)
  int32_t *rowid = nullptr;    // [rows]  new global id of sorted row j (only when !symmetric_order)
  uint32_t *slice_ptr = nullptr;   // [n_slices+1] in int4 groups
  int4 *sell = nullptr;            // [n_groups]   built lazily on the first PageRank call
  float4 *exact_vals = nullptr;    // gathered values of the exact slices, rewritten every iteration (layout: see pull.cu exact_setup)
  int32_t x_blocks = 0, x_tiles = 0;   // ordered-sum blocks of the exact rows (ordered_sum.cuh), tiles of 32 rows x 1 block, tables
  uint32_t *x_blk_base = nullptr, *x_tile_base = nullptr /* same allocation */, *x_mx = nullptr, *x_Q = nullptr;
  double *x_S = nullptr;
  uint8_t *x_plan = nullptr;
  int32_t n_chunks = 0;
  int32_t *chunk_slice = nullptr;  // [n_chunks+1]
  int32_t n_heavy_slices = 0, n_heavy_segs = 0, n_fill_wide = 0;
  int32_t *heavy_slice = nullptr, *heavy_first = nullptr;
  int2 *heavy_seg = nullptr;
  float *partial = nullptr;        // [n_heavy_segs * 32]
  std::vector<uint32_t> h_slice_ptr;   // host copy of slice_ptr (band.cu re-derives work tables from it)
  BandLayout band;
};

}  // namespace gdn

struct gdn_graph {
  int64_t m = 0;                 // global vertex count
  int64_t row_lo = 0, row_hi = 0;
  bool symmetric = false;        // in-CSR aliases out-CSR
  bool has_out = false, has_in = false;
  gdn::DevCsr out, in;
  size_t device_bytes = 0;
  // PageRank scratch
  float *contrib[2] = {nullptr, nullptr};
  int32_t *out_degree = nullptr;     // int32[rows]; for PR on directed graphs
  double *err_partial = nullptr;     // per-warp partial L1 deltas
  double *err_trace = nullptr;       // double[GDN_MAX_PR_ITER + 8] on device (the tail slot takes sum |scores_0|)
  double *abs_partial = nullptr;     // per-warp partials of sum |scores_0| (pr_sell_load)
  int32_t *pr_done = nullptr;        // device flag: converged
  int32_t *pr_work = nullptr;        // device counter: work batches drawn by the warps of the running iteration kernel
  int n_err_partial = 0;
  gdn::PullLayout pull;
  float *scores_sorted = nullptr;    // PR scores in sorted row order during a solve
  // SpMV rows longer than kSpmvExactLen: products row-major in 512-entry blocks + the per-block tables of ordered_core.cuh
  struct SpmvExact {
    bool tried = false;
    int32_t n_rows = 0;
    uint32_t n_blocks = 0;
    uint32_t *blk_base = nullptr;    // [n_rows + 1] first block of exact row i (= heavy_row[n_heavy_rows_lt + i])
    float4 *vals = nullptr;
    double *S = nullptr;
    uint32_t *mx = nullptr, *Q = nullptr;
    uint8_t *plan = nullptr;
  } spmv_exact;
  // SpMV on skewed graphs: hot-first column ids + x scattered into that order (gather.cu spmv_hot_columns)
  bool spmv_tried = false;
  int32_t *spmv_col = nullptr;
  float *spmv_x = nullptr;
  double spmv_prep_ms = 0;
  // row partition: the contrib vectors of the other GPUs mapped here (comm.cu pull_peer_setup); [k][own rank] = contrib[k]
  bool peer_ready = false, peer_failed = false;
  float *peer_contrib[2][8] = {};
  double prep_ms[4] = {0, 0, 0, 0};  // wall time of: create (upload + host layout), SELL build, band build, BFS hubs-first copy
  // BFS scratch
  uint32_t *visited = nullptr, *front = nullptr, *next = nullptr;
  uint32_t *iso = nullptr;           // static: vertices without in-edges (+ pad bits), pre-set in `visited`
  int32_t *queue[2] = {nullptr, nullptr};
  int32_t *heavy_queue = nullptr;     // top-down: rows deferred to td_heavy ...
  uint32_t *heavy_off = nullptr;      // ... and the first 128-edge piece of each (flattened work list)
  int64_t heavy_cap = 0;
  uint8_t *deg_class = nullptr;       // uint8[m]: log-scale out-degree class of every vertex, 0 = hub (host-built at create)
  bool one_shot = false;              // graph lives for ONE solve (oneshot.cu): skip layouts that only pay off when amortised
  int32_t *col_bu = nullptr;          // bottom-up copy of the in-CSR columns, every row reordered hubs-first
  int2 *bu_head = nullptr;            // first two entries of every row of col_bu (bfs.cu bu_head_build)
  unsigned long long *bfs_reached = nullptr;
  uint32_t *xbuf = nullptr;          // partitioned BFS: receive buffer of the OR-merge (P bitmap slices)
  int32_t *parent_buf = nullptr;     // partitioned BFS: parents, P equal slices (global index)
  void *counters = nullptr;          // BfsCounters on device
  void *bfs_ctrl = nullptr, *bfs_ctrl_host = nullptr;   // BfsCtrl: device-side controller state and its pinned host copy
  int64_t n_words = 0;               // bitmap words (32-bit), padded to a multiple of 32
  int64_t bm_alloc_words = 0;        // allocated words per bitmap (>= n_words; room for the allgather slices)
  // chunked upload of the pull CSR's column array (graph.cu upload_csr_begin): col_ev[k] fires when entries
  // [0, col_end[k]) are resident.  Transient: consumed by pull_prepare / upload_csr_end.
  std::vector<cudaEvent_t> col_ev;
  std::vector<uint64_t> col_end;
  int *col_flag = nullptr;           // device verdict word shared with validate_csr (3 = column index out of range)
  const int32_t *col_host = nullptr; // pieces not queued yet: host source, device destination, next entry, end
  int32_t *col_dev = nullptr;
  uint64_t col_next = 0, col_total = 0;
};
namespace gdn { int col_upload_rest(gdn_graph *g); void band_free(BandLayout &b); }   // graph.cu: queue the remaining pieces of the column array

namespace gdn {

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
constexpr unsigned kFull = 0xffffffffu;

// Streaming 256-bit load (Blackwell LDG.E.256) of column indices / matrix
// values: read-only path, no L1 allocation (L1 is kept for the gathered
// vector), L2 evict-first so the one-pass stream does not displace the
// gathered vector from the 126 MB L2.  p must be 32-byte aligned.
__device__ __forceinline__ void ld_stream_v8(const void *p, int (&q)[8]) {
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
      : "l"(p));
}
__device__ __forceinline__ void ld_stream_v8(const void *p, float (&q)[8]) {
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(q[3]), "=f"(q[4]), "=f"(q[5]), "=f"(q[6]), "=f"(q[7])
      : "l"(p));
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Streaming 128-bit load with an L2 eviction policy (see l2_policy_evict_first).
__device__ __forceinline__ int4 ld_stream_v4(const int4 *p, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
// Gathered vector element: read-only path, L1 allocate, L2 evict-last.
__device__ __forceinline__ float ld_gather_f32(const float *p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
// Scalar streaming load (row tails / unaligned heads).
__device__ __forceinline__ int ld_stream_s32(const int *p, uint64_t pol) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

}  // namespace gdn
