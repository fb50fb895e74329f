// bfs.cu -- direction-optimizing BFS (top-down frontier expand + bottom-up
// packed-bitmap sweep).
//
// Replaces (not ports) the reference's src/bfs/{linear_*,topo_*,bottom_up,
// hybrid_*}.cu.  The controller reproduces src/bfs/omp_beamer.cc:97-171 exactly
// (alpha=15, beta=18, integer arithmetic, `scout_count = 1` after a bottom-up
// run), so the per-step direction choice is the oracle's; depths are therefore
// bit-identical to bfs_omp_beamer.
//
// Differences in representation (all invisible in the result):
//  * "visited" is a packed 32-bit-word bitmap claimed with atomicOr (the
//    reference CASes a 4-byte depth per edge, omp_beamer.cc:46-52); depth[] is
//    written exactly once per vertex, already in output form (GDN_INFINITY for
//    unreached, omp_beamer.cc:166-169);
//  * frontier bitmaps are packed 1 bit/vertex (reference GPU variants use an
//    int per vertex + thrust::reduce + cudaMemset, src/bfs/hybrid_base.cu:104-144);
//    the bottom-up step builds `next` with warp ballots -- no atomics;
//  * top-down load balance: per 32 frontier vertices, rows >= kTdHeavy go to a
//    grid-wide kernel, rows >= 32 are strip-mined by the warp, shorter rows are
//    packed edge-by-edge onto lanes via a warp prefix sum; queue appends are
//    warp-aggregated (one atomicAdd per warp per round).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
#include <algorithm>
#include <chrono>
namespace cg = cooperative_groups;

namespace gdn {

constexpr int kTdHeavy = 8192;     // largest row a single warp strip-mines; the host lowers the cut for small frontiers
constexpr int kTdPiece = 128;      // edges per work piece of a deferred ("heavy") row
constexpr int kBuReorderMax = 16384;   // rows longer than this keep their order in the hubs-first copy
constexpr int kBuSerial = 34;      // in-neighbours a lane probes alone (2 from the head array, then 4 per step) before the warp helps
constexpr int kAlpha = 15, kBeta = 18;   // src/bfs/omp_beamer.cc:111

struct BfsCounters {
  long long scout;      // TD: sum of out-degree of newly claimed vertices (omp_beamer.cc:50)
  long long awake;      // BU: vertices discovered (omp_beamer.cc:23)
  long long degsum;     // BU: sum of out-degree of discovered vertices (TEPS accounting)
  int tail;             // next-queue length
  int pad;
  long long bu_edges;   // BU: in-edges probed; TD: edges of the frontier (roofline accounting)
  long long bu_scanned; // BU: unvisited vertices swept
  unsigned long long heavy_pack;   // high 32: deferred rows, low 32: their 128-edge pieces (ONE atomic allocates both)
  int b2q_tail;         // queue length produced by a bitmap -> queue conversion (device-side controller)
  unsigned int ticket;  // blocks that have finished the current phase (device-side controller)
};

// Device-resident state of the direction-optimizing controller (src/bfs/omp_beamer.cc:129-160) -- bfs_persist below.
enum { kModeTD = 0, kModeBU = 1, kModeDone = 2 };
enum { kConvNone = 0, kConvQ2B = 1, kConvB2Q = 2 };
struct BfsCtrl {
  long long edges_to_check, scout_count, old_awake;
  long long reached, reached_deg;
  long long bu_ns;        // time spent in bottom-up sweeps (globaltimer), for the roofline's kernel share
  long long t_step;       // globaltimer at the end of the previous step
  int n_front;            // |frontier| about to be expanded
  int mode, convert;
  int cur;                // queue holding the frontier
  int fb;                 // bitmap holding the frontier
  int level, iter, n_steps;
  int aborted;            // safety net: more steps than vertices
  int pad;
  gdn_bfs_step steps[GDN_MAX_BFS_STEPS];
};

struct BfsState {
  uint32_t *visited;
  int32_t *depth;
  int32_t *parent;      // nullable
  int32_t *q_out;
  BfsCounters *cnt;
  uint32_t *mark;       // partitioned mode: discovered-bitmap (full length); nullptr on one GPU
  int64_t row_lo;       // first global vertex id owned by this GPU (rowptr is local)
};

template <typename OffT>
__device__ __forceinline__ long long td_visit(const OffT *__restrict__ rowptr, const BfsState &s, int dst, int src,
                                              int level, int lane) {
  if (s.mark) {
    // partitioned top-down: the destination may live on another GPU, so only record it;
    // claiming (depth, queue, scout) happens in bfs_absorb after the OR-merge across GPUs.
    if (dst >= 0) {
      const uint32_t w = (uint32_t)dst >> 5, bit = 1u << (dst & 31);
      if (!(s.visited[w] & bit) && !(s.mark[w] & bit)) atomicOr(&s.mark[w], bit);
    }
    return 0;
  }
  bool claimed = false;
  if (dst >= 0) {
    const uint32_t w = (uint32_t)dst >> 5, bit = 1u << (dst & 31);
    if (!(s.visited[w] & bit)) {
      const uint32_t old = atomicOr(&s.visited[w], bit);
      claimed = !(old & bit);
    }
  }
  const unsigned cm = __ballot_sync(kFull, claimed);
  long long deg = 0;
  if (cm) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s.cnt->tail, __popc(cm));
    base = __shfl_sync(kFull, base, 0);
    if (claimed) {
      s.q_out[base + __popc(cm & ((1u << lane) - 1))] = dst;
      s.depth[dst] = level;
      if (s.parent) s.parent[dst] = src;
      deg = (long long)(rowptr[dst + 1] - rowptr[dst]);
    }
  }
  return deg;
}

// Four 32-edge groups at once (td_heavy): the visited probes, the claims and the degree look-ups of the
// four destinations are independent, so their latencies overlap.  Claimed vertices are staged in a
// per-warp shared-memory buffer and appended to the next queue kTdStage at a time: one reservation
// (atomicAdd on the single queue tail) per ~30 pieces instead of one per piece -- with a million pieces
// in the largest top-down step of a Kron-26 BFS the same-address atomic was the bottleneck.
constexpr int kTdStage = 512;
__device__ __forceinline__ void td_flush(const BfsState &s, const int *buf, int &nbuf, int lane) {
  if (nbuf == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(&s.cnt->tail, nbuf);
  base = __shfl_sync(kFull, base, 0);
  __syncwarp();
  for (int i = lane; i < nbuf; i += 32) s.q_out[base + i] = buf[i];
  __syncwarp();
  nbuf = 0;
}

template <typename OffT>
__device__ __forceinline__ long long td_visit4(const OffT *__restrict__ rowptr, const BfsState &s, const int (&dst)[4], int src,
                                               int level, int lane, int *buf, int &nbuf) {
  if (s.mark) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (dst[u] >= 0) {
        const uint32_t w = (uint32_t)dst[u] >> 5, bit = 1u << (dst[u] & 31);
        if (!(s.visited[w] & bit) && !(s.mark[w] & bit)) atomicOr(&s.mark[w], bit);
      }
    }
    return 0;
  }
  uint32_t word[4];
#pragma unroll
  for (int u = 0; u < 4; u++) word[u] = dst[u] >= 0 ? s.visited[(uint32_t)dst[u] >> 5] : 0xffffffffu;
  bool claimed[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const uint32_t bit = 1u << (dst[u] & 31);
    claimed[u] = false;
    if (!(word[u] & bit)) claimed[u] = !(atomicOr(&s.visited[(uint32_t)dst[u] >> 5], bit) & bit);
  }
  unsigned cm[4];
  int total = 0;
#pragma unroll
  for (int u = 0; u < 4; u++) { cm[u] = __ballot_sync(kFull, claimed[u]); total += __popc(cm[u]); }
  long long deg = 0;
  if (total) {
    if (nbuf + total > kTdStage) td_flush(s, buf, nbuf, lane);
    int base = nbuf;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (claimed[u]) {
        buf[base + __popc(cm[u] & ((1u << lane) - 1))] = dst[u];
        s.depth[dst[u]] = level;
        if (s.parent) s.parent[dst[u]] = src;
        deg += (long long)(rowptr[dst[u] + 1] - rowptr[dst[u]]);
      }
      base += __popc(cm[u]);
    }
    nbuf += total;
  }
  return deg;
}

// TDStep, src/bfs/omp_beamer.cc:35-58.
template <typename OffT>
__device__ __forceinline__ void td_expand_dev(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *q_in,
                                              int n_in, const BfsState &s, int32_t *heavy_q, uint32_t *heavy_off, uint32_t heavy_cut, int level) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  long long scout = 0, edges = 0;
  for (int base = warp * 32; base < n_in; base += nwarps * 32) {
    const int idx = base + lane;
    int v = -1;
    OffT b = 0, e = 0;
    if (idx < n_in) { v = q_in[idx]; b = rowptr[v - s.row_lo]; e = rowptr[v - s.row_lo + 1]; }
    uint32_t deg = (uint32_t)(e - b);
    edges += deg;
    if (deg >= heavy_cut) {                              // tier 1: defer to td_heavy (flattened piece list)
      const unsigned long long old = atomicAdd(&s.cnt->heavy_pack, (1ull << 32) | (unsigned long long)((deg + kTdPiece - 1) / kTdPiece));
      heavy_q[old >> 32] = v;
      heavy_off[old >> 32] = (uint32_t)old;
      deg = 0;
    }
    unsigned med = __ballot_sync(kFull, deg >= 32u);     // tier 2: warp strip-mines the row
    while (med) {
      const int l = __ffs(med) - 1;
      med &= med - 1;
      const OffT bb = __shfl_sync(kFull, b, l), ee = __shfl_sync(kFull, e, l);
      const int src = __shfl_sync(kFull, v, l);
      for (OffT i = bb; i < ee; i += 32) {
        const OffT k = i + lane;
        const int dst = (k < ee) ? col[k] : -1;
        scout += td_visit(rowptr, s, dst, src, level, lane);
      }
    }
    if (deg >= 32u) deg = 0;
    // tier 3: rows shorter than a warp, packed edge-by-edge onto lanes
    uint32_t off = deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(kFull, off, o);
      if (lane >= o) off += n;
    }
    const uint32_t total = __shfl_sync(kFull, off, 31);
    off -= deg;                                          // exclusive
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
      const uint32_t t = t0 + lane;
      int j = 0;                                         // largest lane with off_j <= t
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int cand = j + step;
        const uint32_t o = __shfl_sync(kFull, off, cand & 31);
        if (cand < 32 && o <= t) j = cand;
      }
      const OffT bj = __shfl_sync(kFull, b, j);
      const uint32_t oj = __shfl_sync(kFull, off, j);
      const int src = __shfl_sync(kFull, v, j);
      const int dst = (t < total) ? col[bj + (t - oj)] : -1;
      scout += td_visit(rowptr, s, dst, src, level, lane);
    }
  }
  scout = warp_sum(scout);
  edges = warp_sum(edges);
  if (lane == 0 && scout) atomicAdd((unsigned long long *)&s.cnt->scout, (unsigned long long)scout);
  if (lane == 0 && edges) atomicAdd((unsigned long long *)&s.cnt->bu_edges, (unsigned long long)edges);
}
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
td_expand(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ q_in,
          int n_in, BfsState s, int32_t *heavy_q, uint32_t *heavy_off, uint32_t heavy_cut, int level) {
  if (n_in < 0) n_in = s.cnt->tail;       // partitioned mode: queue length lives on the device
  td_expand_dev<OffT>(rowptr, col, q_in, n_in, s, heavy_q, heavy_off, heavy_cut, level);
}

// Deferred rows: the (row, 128-edge piece) pairs form one flat list (heavy_off[slot] = first piece of
// row slot, monotone in slot because slot and pieces come from the same 64-bit atomic).  Every warp
// takes a contiguous range of pieces: one binary search, then a walk.  A frontier of five hubs is
// expanded by the whole grid instead of five warps.
template <typename OffT>
__device__ __forceinline__ void td_heavy_dev(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const BfsState &s,
                                             const int32_t *heavy_q, const uint32_t *heavy_off, int level, int *s_stage /* [8][kTdStage] */) {
  const unsigned long long pack = *(volatile unsigned long long *)&s.cnt->heavy_pack;
  const uint32_t nh = (uint32_t)(pack >> 32), np = (uint32_t)pack;
  if (nh == 0) return;
  const int lane = threadIdx.x & 31;
  int *buf = s_stage + (threadIdx.x >> 5) * kTdStage;
  int nbuf = 0;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t per = (np + nwarps - 1) / nwarps;
  const uint64_t p0_ = (uint64_t)warp * per;
  if (p0_ >= np) return;
  uint32_t p = (uint32_t)p0_;
  const uint32_t p1 = (uint32_t)min((uint64_t)np, p0_ + per);
  uint32_t lo = 0, hi = nh;                              // heavy_off[lo] <= p < heavy_off[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (heavy_off[mid] <= p) lo = mid; else hi = mid;
  }
  long long scout = 0;
  for (uint32_t slot = lo; p < p1; slot++) {
    const int src = heavy_q[slot];
    const OffT b = rowptr[src - s.row_lo], e = rowptr[src - s.row_lo + 1];
    const uint32_t first = heavy_off[slot];
    const uint32_t pend = min(p1, (slot + 1 < nh) ? heavy_off[slot + 1] : np);
    for (; p < pend; p++) {
      const OffT i0 = b + (OffT)(p - first) * kTdPiece;
      const OffT i1 = (e - i0 > (OffT)kTdPiece) ? i0 + kTdPiece : e;
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const OffT k = i0 + (OffT)(u * 32 + lane);
        dst[u] = (k < i1) ? col[k] : -1;
      }
      scout += td_visit4(rowptr, s, dst, src, level, lane, buf, nbuf);
    }
  }
  if (!s.mark) td_flush(s, buf, nbuf, lane);
  scout = warp_sum(scout);
  if (lane == 0 && scout) atomicAdd((unsigned long long *)&s.cnt->scout, (unsigned long long)scout);
}
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
td_heavy(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, BfsState s,
         const int32_t *__restrict__ heavy_q, const uint32_t *__restrict__ heavy_off, int level) {
  __shared__ int s_stage[8 * kTdStage];
  td_heavy_dev<OffT>(rowptr, col, s, heavy_q, heavy_off, level, s_stage);
}

// Hubs-first copy of the bottom-up CSR.  BFS depths (and the alpha/beta schedule, which only counts
// vertices) do not depend on the order in which a row's in-neighbours are probed, but the COST of
// BUStep does: the sweep stops at the first neighbour found in the frontier (src/bfs/omp_beamer.cc:19-25)
// and on a skewed graph that is almost always a hub.  Every row of at most kBuReorderMax entries is
// stably partitioned by the log-scale degree class of the neighbour (deg_class, 0 = hub); longer rows
// (hubs themselves, discovered in the first top-down steps) are copied unchanged.  One warp per row.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
hubs_first(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const uint8_t *__restrict__ cls,
           int32_t *__restrict__ out, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const OffT b = rowptr[r], e = rowptr[r + 1];
    const OffT len = e - b;
    if (len > (OffT)kBuReorderMax || len <= 1) {
      for (OffT i = b + lane; i < e; i += 32) out[i] = col[i];
      continue;
    }
    // pass 1: entries per class -> lane j holds the running write offset of class j
    uint32_t mycount = 0;
    if (len > 32) {
      for (OffT i = b; i < e; i += 32) {
        const OffT k = i + lane;
        const int kc = (k < e) ? (int)cls[col[k]] : 99;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned mj = __ballot_sync(kFull, kc == j);
          if (lane == j) mycount += __popc(mj);
        }
      }
    }
    uint32_t start = mycount;                             // exclusive scan over lanes 0..7
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const uint32_t n = __shfl_up_sync(kFull, start, o);
      if (lane >= o) start += n;
    }
    start -= mycount;
    // pass 2: stable scatter
    for (OffT i = b; i < e; i += 32) {
      const OffT k = i + lane;
      const int c = (k < e) ? col[k] : -1;
      const int kc = (k < e) ? (int)cls[c] : 99;
      unsigned mk = 0;
      uint32_t add = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const unsigned mj = __ballot_sync(kFull, kc == j);
        if (kc == j) mk = mj;
        if (lane == j) add = __popc(mj);
      }
      if (len <= 32) {                                    // single chunk: class starts from this chunk's counts
        uint32_t st1 = add;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const uint32_t n = __shfl_up_sync(kFull, st1, o);
          if (lane >= o) st1 += n;
        }
        start = st1 - add;
      }
      const uint32_t base = __shfl_sync(kFull, start, kc & 7);
      if (k < e) out[b + base + __popc(mk & lt)] = c;
      start += add;
    }
  }
}

// BUStep, src/bfs/omp_beamer.cc:13-32.  One warp sweeps 32 bitmap words (1024 vertices) at a time.
// The unvisited vertices of the group are first compacted into a shared-memory list (popc + warp scan),
// so that every lane works on a live vertex -- on a Kronecker graph half of the ids are isolated
// (pre-marked in `visited`, see iso_bitmap) and after the first sweep most of the rest are visited --
// and each lane follows TWO vertices at once: the sweep is a chain of dependent loads
// (offsets -> column -> frontier bit), so its speed is the number of chains in flight.
// A lane probes up to kBuSerial in-neighbours alone; rows that have not hit by then are scanned by the
// whole warp with a ballot early exit.  `next` is assembled in shared memory; no global atomics.
template <typename OffT>
__device__ __forceinline__ int bu_warp_scan(const int32_t *__restrict__ col, const uint32_t *front,
                                            OffT bb, OffT ee, int lane, long long &probed) {
  for (OffT i = bb; i < ee; i += 32) {
    const OffT k = i + lane;
    int src = -1;
    bool ok = false;
    if (k < ee) { src = col[k]; ok = (front[(uint32_t)src >> 5] >> (src & 31)) & 1u; probed++; }
    const unsigned bal = __ballot_sync(kFull, ok);
    if (bal) return __shfl_sync(kFull, src, __ffs(bal) - 1);
  }
  return -1;
}

// head[v] = the first two entries of v's (hubs-first) bottom-up row in 8 bytes: x = first neighbour or -1; y = second
// neighbour, -1 when there is none, or -(id) - 2 when the row goes on after it.  Most vertices of a skewed graph are
// settled by these two probes, which cost a quarter of a sector each (consecutive vertices share it) instead of a sector
// of the offsets array plus a sector of the column array.
template <typename OffT>
__global__ void bu_head_build(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, int2 *__restrict__ head, int64_t rows) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const OffT b = rowptr[r], e = rowptr[r + 1];
    int2 h = make_int2(-1, -1);
    if (e > b) h.x = col[b];
    if (e > b + 1) { h.y = col[b + 1]; if (e > b + 2) h.y = -h.y - 2; }
    head[r] = h;
  }
}

// depth == nullptr: the caller writes the depths of this level itself (the partitioned path: bfs_absorb, on every GPU).
// (Writing them at the end of the BFS from per-level `next` bitmaps was measured: the sweeps got 0.15 ms faster, the extra
// pass cost 0.19 ms.)  head == nullptr: no head array (one-shot graphs).
template <typename OffT>
__device__ __forceinline__ void bu_sweep_dev(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int2 *__restrict__ head,
                                             const uint32_t *front, uint32_t *next, uint32_t *visited, int32_t *depth,
                                             int32_t *parent, int64_t word_lo, int64_t word_hi, int64_t row_lo, bool update_visited, int level,
                                             BfsCounters *cnt, uint16_t *s_list /* [8][2048] */, uint32_t *s_next /* [8][32] */) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t g_lo = word_lo >> 5, g_hi = word_hi >> 5;   // word ranges are multiples of 32
  rowptr -= row_lo;                                          // rows are addressed by global vertex id
  if (head) head -= row_lo;
  uint16_t *list = s_list + wib * 2048;                      // [0, 1024): unvisited vertices of the group; [1024, 2048): pass B's list
  uint32_t *nx = s_next + wib * 32;
  long long awake = 0, probed = 0, swept = 0;
  for (int64_t g = g_lo + warp; g < g_hi; g += nwarps) {
    const int64_t widx = g * 32 + lane;
    const uint32_t vis = visited[widx];
    uint32_t act = ~vis;
    const int c = __popc(act);
    int off = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(kFull, off, o);
      if (lane >= o) off += n;
    }
    const int total = __shfl_sync(kFull, off, 31);
    if (total == 0) { next[widx] = 0; continue; }
    swept += c;
    nx[lane] = 0;
    off -= c;
    while (act) {
      const int bit = __ffs(act) - 1;
      act &= act - 1;
      list[off++] = (uint16_t)((lane << 5) | bit);
    }
    __syncwarp();
    const int64_t vbase = g * 1024;
    // Pass A (graphs with a head array): the first two entries of every unvisited vertex's row, four vertices per lane in
    // flight -- two short chains of loads (head -> frontier word) that settle most vertices of a skewed graph.  The
    // vertices that still have unprobed entries are compacted into list2: pass B then runs with every lane busy.
    uint16_t *list2 = list;
    int total2 = total;
    if (head) {
      list2 = list + 1024;
      total2 = 0;
      for (int i0 = 0; i0 < total; i0 += 128) {
        int l[4];
        int2 h[4];
        uint32_t w0[4], w1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int idx = i0 + u * 32 + lane;
          l[u] = idx < total ? list[idx] : -1;
          h[u] = l[u] >= 0 ? head[vbase + l[u]] : make_int2(-1, -1);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) w0[u] = h[u].x >= 0 ? front[(uint32_t)h[u].x >> 5] : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int y = h[u].y < -1 ? -h[u].y - 2 : h[u].y;
          const bool hit0 = h[u].x >= 0 && ((w0[u] >> (h[u].x & 31)) & 1u);
          w1[u] = (!hit0 && y >= 0) ? front[(uint32_t)y >> 5] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int y = h[u].y < -1 ? -h[u].y - 2 : h[u].y;
          const bool hit0 = h[u].x >= 0 && ((w0[u] >> (h[u].x & 31)) & 1u);
          const bool hit1 = !hit0 && y >= 0 && ((w1[u] >> (y & 31)) & 1u);
          const int par = hit0 ? h[u].x : hit1 ? y : -1;
          probed += (int)(h[u].x >= 0) + (int)(!hit0 && y >= 0);
          if (par >= 0) {
            const int64_t v = vbase + l[u];
            if (depth) depth[v] = level;
            if (parent) parent[v] = par;
            atomicOr(&nx[l[u] >> 5], 1u << (l[u] & 31));
          }
          const bool more = l[u] >= 0 && par < 0 && h[u].y < -1;
          const unsigned fm = __ballot_sync(kFull, par >= 0), mm = __ballot_sync(kFull, more);
          if (lane == 0) awake += __popc(fm);
          if (more) list2[total2 + __popc(mm & ((1u << lane) - 1u))] = (uint16_t)l[u];
          total2 += __popc(mm);
        }
      }
      __syncwarp();
    }
    const int skip = head ? 2 : 0;                           // entries of a row that pass A has covered
    // Pass B: the rest of the rows, two vertices per lane
    for (int i0 = 0; i0 < total2; i0 += 64) {
      const bool on_a = i0 + lane < total2, on_b = i0 + 32 + lane < total2;
      const int la = on_a ? list2[i0 + lane] : 0, lb = on_b ? list2[i0 + 32 + lane] : 0;
      const int64_t va = vbase + la, vb = vbase + lb;
      int pa = -1, pb = -1;
      const bool more_a = on_a, more_b = on_b;
      OffT ia = 0, ea = 0, ib = 0, eb = 0;
      if (more_a) { ia = rowptr[va] + (OffT)skip; ea = rowptr[va + 1]; }
      if (more_b) { ib = rowptr[vb] + (OffT)skip; eb = rowptr[vb + 1]; }
      // every lane walks ITS two rows, four entries per step (the four column loads, then the four frontier probes, are
      // independent: two memory latencies per step instead of eight), up to kBuSerial entries; what is left of longer
      // rows is scanned by the whole warp below
      const OffT lima = (ea - ia > (OffT)(kBuSerial - skip)) ? ia + (kBuSerial - skip) : ea;
      const OffT limb = (eb - ib > (OffT)(kBuSerial - skip)) ? ib + (kBuSerial - skip) : eb;
#pragma unroll 1
      for (;;) {
        const bool ga = pa < 0 && ia < lima, gb = pb < 0 && ib < limb;
        if (!__any_sync(kFull, ga || gb)) break;
        int sa[4], sb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          sa[u] = (ga && ia + u < lima) ? col[ia + u] : -1;
          sb[u] = (gb && ib + u < limb) ? col[ib + u] : -1;
        }
        uint32_t wa[4], wb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          wa[u] = sa[u] >= 0 ? front[(uint32_t)sa[u] >> 5] : 0u;
          wb[u] = sb[u] >= 0 ? front[(uint32_t)sb[u] >> 5] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {                          // first hit in row order; entries after it are not counted as probed
          if (pa < 0 && sa[u] >= 0) { probed++; if ((wa[u] >> (sa[u] & 31)) & 1u) pa = sa[u]; }
          if (pb < 0 && sb[u] >= 0) { probed++; if ((wb[u] >> (sb[u] & 31)) & 1u) pb = sb[u]; }
        }
        ia += 4; ib += 4;
      }
      // long rows that have not hit yet: the whole warp scans the remainder
      unsigned rest = __ballot_sync(kFull, more_a && pa < 0 && lima < ea);
      while (rest) {
        const int l = __ffs(rest) - 1;
        rest &= rest - 1;
        const int hit = bu_warp_scan<OffT>(col, front, __shfl_sync(kFull, lima, l), __shfl_sync(kFull, ea, l), lane, probed);
        if (lane == l) pa = hit;
      }
      rest = __ballot_sync(kFull, more_b && pb < 0 && limb < eb);
      while (rest) {
        const int l = __ffs(rest) - 1;
        rest &= rest - 1;
        const int hit = bu_warp_scan<OffT>(col, front, __shfl_sync(kFull, limb, l), __shfl_sync(kFull, eb, l), lane, probed);
        if (lane == l) pb = hit;
      }
      if (pa >= 0) {
        if (depth) depth[va] = level;
        if (parent) parent[va] = pa;
        atomicOr(&nx[la >> 5], 1u << (la & 31));
      }
      if (pb >= 0) {
        if (depth) depth[vb] = level;
        if (parent) parent[vb] = pb;
        atomicOr(&nx[lb >> 5], 1u << (lb & 31));
      }
      const unsigned fa = __ballot_sync(kFull, pa >= 0), fb = __ballot_sync(kFull, pb >= 0);
      if (lane == 0) awake += __popc(fa) + __popc(fb);
    }
    __syncwarp();
    const uint32_t nxt = nx[lane];
    next[widx] = nxt;
    if (nxt && update_visited) visited[widx] = vis | nxt;
    __syncwarp();
  }
  awake = warp_sum(awake);
  probed = warp_sum(probed);
  swept = warp_sum(swept);
  if (lane == 0 && swept) {
    atomicAdd((unsigned long long *)&cnt->bu_edges, (unsigned long long)probed);
    atomicAdd((unsigned long long *)&cnt->bu_scanned, (unsigned long long)swept);
  }
  if (lane == 0 && awake) atomicAdd((unsigned long long *)&cnt->awake, (unsigned long long)awake);
}
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bu_sweep(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int2 *__restrict__ head,
         const uint32_t *__restrict__ front, uint32_t *__restrict__ next, uint32_t *visited, int32_t *depth,
         int32_t *parent, int64_t word_lo, int64_t word_hi, int64_t row_lo, bool update_visited, int level,
         BfsCounters *cnt) {
  __shared__ uint16_t s_list[8 * 2048];
  __shared__ uint32_t s_next[8 * 32];
  bu_sweep_dev<OffT>(rowptr, col, head, front, next, visited, depth, parent, word_lo, word_hi, row_lo, update_visited, level, cnt,
                     s_list, s_next);
}

// Vertices without in-edges can never be discovered (omp_beamer.cc:17-26 finds no neighbour, TDStep never
// sees them as a destination): their bits are pre-set in `visited` so that the sweeps skip them.  Static per
// graph; a row partition marks its own rows only (the only ones it sweeps).
template <typename OffT>
__global__ void iso_bitmap(const OffT *__restrict__ in_rowptr, int64_t row_lo, int64_t row_hi, int64_t m, int64_t n_words,
                           uint32_t *__restrict__ iso) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < n_words; w += nwarps) {
    const int64_t v = w * 32 + lane;
    bool z = v >= m;                                       // pad bits count as visited
    if (v >= row_lo && v < row_hi) z = in_rowptr[v - row_lo + 1] == in_rowptr[v - row_lo];
    const unsigned mask = __ballot_sync(kFull, z);
    if (lane == 0) iso[w] = mask;
  }
}

// QueueToBitmap, src/bfs/omp_beamer.cc:60-67
__device__ __forceinline__ void queue_to_bitmap_dev(const int32_t *q, int n, uint32_t *bm) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = q[i];
    atomicOr(&bm[(uint32_t)v >> 5], 1u << (v & 31));
  }
}
__global__ void queue_to_bitmap(const int32_t *__restrict__ q, int n, uint32_t *bm) { queue_to_bitmap_dev(q, n, bm); }

// BitmapToQueue, src/bfs/omp_beamer.cc:69-79: ballot-free popc compaction, one
// atomicAdd per 1024 vertices.
__device__ __forceinline__ void bitmap_to_queue_dev(const uint32_t *bm, int64_t word_lo, int64_t word_hi, int32_t *q, int *tail) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = (word_lo >> 5) + warp; g < (word_hi >> 5); g += nwarps) {
    const int64_t widx = g * 32 + lane;
    uint32_t word = bm[widx];
    const int c = __popc(word);
    int off = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(kFull, off, o);
      if (lane >= o) off += n;
    }
    const int total = __shfl_sync(kFull, off, 31);
    if (total == 0) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(tail, total);
    base = __shfl_sync(kFull, base, 0) + off - c;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      q[base++] = (int32_t)(widx * 32 + bit);
    }
  }
}
__global__ void __launch_bounds__(256, 4)
bitmap_to_queue(const uint32_t *__restrict__ bm, int64_t word_lo, int64_t word_hi, int32_t *q, BfsCounters *cnt) {
  bitmap_to_queue_dev(bm, word_lo, word_hi, q, &cnt->tail);
}

template <typename OffT>
__global__ void bfs_init(const OffT *__restrict__ out_rowptr, int32_t *depth, int32_t *parent, uint32_t *visited,
                         int64_t m, int64_t n_words, int source, int32_t *queue0, BfsCounters *cnt,
                         uint32_t *front /* partitioned mode: bitmap holding just the source */, int64_t row_lo,
                         int64_t row_hi, const uint32_t *__restrict__ iso) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = tid; v < m; v += nth) {
    depth[v] = (v == source) ? 0 : GDN_INFINITY;
    if (parent) parent[v] = (v == source) ? source : -1;
  }
  for (int64_t w = tid; w < n_words; w += nth) {
    uint32_t word = iso[w];                                 // isolated vertices + pad bits: "visited"
    if ((int64_t)(source >> 5) == w) word |= 1u << (source & 31);
    visited[w] = word;
    if (front) front[w] = ((int64_t)(source >> 5) == w) ? (1u << (source & 31)) : 0u;
  }
  if (tid == 0) {
    queue0[0] = source;
    const bool own = source >= row_lo && source < row_hi;
    cnt->scout = own ? (long long)(out_rowptr[source - row_lo + 1] - out_rowptr[source - row_lo]) : 0;   // degrees[source], omp_beamer.cc:130
    cnt->awake = 0; cnt->degsum = 0; cnt->tail = 0; cnt->pad = 0; cnt->bu_edges = 0; cnt->bu_scanned = 0; cnt->heavy_pack = 0;
    cnt->b2q_tail = 0; cnt->ticket = 0;
  }
}

// Partitioned mode, after the frontier bitmap has been merged across GPUs: claim
// the newly discovered vertices on EVERY GPU (visited and depth are replicated),
// emit them as the next frontier bitmap, and sum the out-degree of the ones this
// GPU owns (scout_count, src/bfs/omp_beamer.cc:50).
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bfs_absorb(const uint32_t *__restrict__ merged, uint32_t *visited, uint32_t *__restrict__ front_out, int32_t *depth,
           const OffT *__restrict__ out_rowptr, int64_t n_words, int64_t own_word_lo, int64_t own_word_hi,
           int64_t row_lo, int level, BfsCounters *cnt) {
  const int lane = threadIdx.x & 31;
  long long fresh = 0, scout = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t vis = visited[w];
    uint32_t nw = merged[w] & ~vis;
    front_out[w] = nw;
    if (nw) {
      visited[w] = vis | nw;
      fresh += __popc(nw);
      const bool own = (w >= own_word_lo && w < own_word_hi);
      while (nw) {
        const int bit = __ffs(nw) - 1;
        nw &= nw - 1;
        const int64_t v = w * 32 + bit;
        depth[v] = level;
        if (own) scout += (long long)(out_rowptr[v - row_lo + 1] - out_rowptr[v - row_lo]);
      }
    }
  }
  fresh = warp_sum(fresh);
  scout = warp_sum(scout);
  if (lane == 0 && fresh) {
    atomicAdd((unsigned long long *)&cnt->awake, (unsigned long long)fresh);
    if (scout) atomicAdd((unsigned long long *)&cnt->scout, (unsigned long long)scout);
  }
}

int comm_size();
int comm_rank();
int64_t partition_width(int64_t m, int nparts);

static int bfs_alloc(gdn_graph *g) {
  const int64_t words = (g->m + 31) / 32;
  g->n_words = (words + 31) / 32 * 32;
  // allgather needs room for comm_size() equal-width slices
  const int64_t need = std::max<int64_t>(g->n_words, partition_width(g->m, comm_size()) / 32 * comm_size());
  if (g->visited && g->bm_alloc_words >= need) return GDN_OK;
  if (g->visited) {
    cudaFree(g->visited); cudaFree(g->front); cudaFree(g->next); cudaFree(g->xbuf); cudaFree(g->iso);
    g->visited = g->front = g->next = g->xbuf = g->iso = nullptr;
  }
  const size_t bm = sizeof(uint32_t) * need;
  GDN_CUDA(cudaMalloc((void **)&g->visited, bm));
  GDN_CUDA(cudaMalloc((void **)&g->front, bm));
  GDN_CUDA(cudaMalloc((void **)&g->next, bm));
  GDN_CUDA(cudaMalloc((void **)&g->iso, bm));
  {
    const DevCsr &ci = g->symmetric ? g->out : g->in;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((g->n_words + 7) / 8, (int64_t)lib().sm_count * 8));
    if (ci.off64) iso_bitmap<uint64_t><<<grid, 256, 0, lib().stream>>>((const uint64_t *)ci.rowptr, g->row_lo, g->row_hi, g->m, g->n_words, g->iso);
    else iso_bitmap<uint32_t><<<grid, 256, 0, lib().stream>>>((const uint32_t *)ci.rowptr, g->row_lo, g->row_hi, g->m, g->n_words, g->iso);
  }
  GDN_CUDA(cudaMemsetAsync(g->front, 0, bm, lib().stream));
  GDN_CUDA(cudaMemsetAsync(g->next, 0, bm, lib().stream));
  g->bm_alloc_words = need;
  g->device_bytes += 4 * bm;
  if (!g->queue[0]) {
    GDN_CUDA(cudaMalloc((void **)&g->queue[0], sizeof(int32_t) * std::max<int64_t>(g->m, 1)));
    GDN_CUDA(cudaMalloc((void **)&g->queue[1], sizeof(int32_t) * std::max<int64_t>(g->m, 1)));
    // deferred rows have >= 32 entries (td_cut), so there are at most nnz/32 of them in one step
    const int64_t hq = std::min<int64_t>(g->row_hi - g->row_lo, (int64_t)(g->out.nnz / 32)) + 2;
    GDN_CUDA(cudaMalloc((void **)&g->heavy_queue, sizeof(int32_t) * hq));
    GDN_CUDA(cudaMalloc((void **)&g->heavy_off, sizeof(uint32_t) * hq));
    g->heavy_cap = hq;
    GDN_CUDA(cudaMalloc((void **)&g->counters, sizeof(BfsCounters)));
    g->device_bytes += 2 * sizeof(int32_t) * g->m + 2 * sizeof(int32_t) * hq + sizeof(BfsCounters);
  }
  return GDN_OK;
}

// Rows at least this long are deferred to td_heavy.  `edges` = scout_count = the number of edges this
// top-down step scans (known exactly from the previous step, omp_beamer.cc:155): aim at equal work per
// warp of a full grid, never below one warp-width and never above what one warp should strip-mine.
__host__ __device__ __forceinline__ uint32_t td_cut(long long edges, int sm) {
  // a warp of td_expand owns 32 frontier rows: bound the ROW length by 1/32 of the per-warp share
  const long long per_row = edges / ((long long)sm * 4 * 8 * 32);
  return (uint32_t)(per_row < 32 ? 32 : (per_row > kTdHeavy ? kTdHeavy : per_row));
}

// Build the hubs-first copy of the bottom-up columns on first use (4 bytes per edge, ~50 ms at Kron-26).
template <typename OffT>
static int bfs_prepare(gdn_graph *g) {
  if (g->col_bu || !g->deg_class || g->one_shot || getenv("GDN_BFS_NO_REORDER")) return GDN_OK;
  const DevCsr &ci = g->symmetric ? g->out : g->in;
  if (ci.nnz == 0) return GDN_OK;
  const auto t0 = std::chrono::steady_clock::now();
  GDN_CUDA(cudaMalloc((void **)&g->col_bu, sizeof(int32_t) * ci.nnz + 256));
  g->device_bytes += sizeof(int32_t) * ci.nnz;
  hubs_first<OffT><<<lib().sm_count * 8, 256, 0, lib().stream>>>((const OffT *)ci.rowptr, ci.col, g->deg_class, g->col_bu, ci.rows);
  GDN_CUDA(cudaMalloc((void **)&g->bu_head, sizeof(int2) * std::max<int64_t>(ci.rows, 1)));
  g->device_bytes += sizeof(int2) * ci.rows;
  bu_head_build<OffT><<<lib().sm_count * 8, 256, 0, lib().stream>>>((const OffT *)ci.rowptr, g->col_bu, g->bu_head, ci.rows);
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  GDN_CUDA(cudaGetLastError());
  g->prep_ms[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return GDN_OK;
}

// ------------------------------------------------------------------ the whole BFS as ONE cooperative kernel
// The reference's CUDA variants -- and round 1 of this file -- return to the host after every level to read the
// counters and choose the next step (src/bfs/hybrid_base.cu:104-144): seven or more round trips of ~30 us, more than the
// kernels themselves at Kronecker scale 22.  Here the alpha/beta controller of src/bfs/omp_beamer.cc:135-160 lives on the
// device: a persistent grid (4 CTAs of 256 threads per SM, co-resident) walks the levels, with a grid-wide barrier between
// the phases of a step; the LAST CTA to finish a step (ticket counter) reads the step's counters, records the step and
// decides direction and conversions for the next one.  One launch, one host synchronisation per BFS.
struct PersistArgs {
  int32_t *queue[2];
  uint32_t *bm[2];
  const int2 *head;               // nullable
  uint32_t *visited;
  int32_t *depth, *parent;
  int32_t *heavy_q;
  uint32_t *heavy_off;
  BfsCounters *cnt;
  BfsCtrl *ctrl;
  int64_t m, n_words;
  int sm;
};

__device__ __forceinline__ unsigned long long bfs_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ctrl_record(BfsCtrl *c, int dir, long long frontier, long long disc, long long sc, long long edges, long long scanned) {
  const long long now = (long long)bfs_timer_ns();
  if (c->n_steps < GDN_MAX_BFS_STEPS) {
    gdn_bfs_step &b = c->steps[c->n_steps];
    b.dir = dir; b.ns = (int32_t)(now - c->t_step); b.frontier = frontier; b.discovered = disc; b.scout = sc; b.edges = edges; b.scanned = scanned;
  }
  c->n_steps++;
  c->t_step = now;
}
// `while (!queue.empty()) { if (scout_count > edges_to_check / alpha) ... else ...` (omp_beamer.cc:135-137,152-154)
__device__ __forceinline__ void ctrl_decide(BfsCtrl *c, bool in_bitmap) {
  if (c->n_front == 0) { c->mode = kModeDone; return; }
  if (c->scout_count > c->edges_to_check / kAlpha) {
    c->mode = kModeBU;
    c->convert = in_bitmap ? kConvNone : kConvQ2B;
    c->old_awake = c->n_front;                        // awake_count = queue.size(), :139
  } else {
    c->mode = kModeTD;
    c->convert = in_bitmap ? kConvB2Q : kConvNone;
    c->edges_to_check -= c->scout_count;              // :154
  }
  c->iter++;
}
// Executed by ONE thread once every CTA has finished the step.
__device__ __forceinline__ void ctrl_after_step(BfsCtrl *c, BfsCounters *cnt, int64_t m) {
  volatile BfsCounters *v = cnt;
  if (c->mode == kModeTD) {
    const long long scout = v->scout;
    const int tail = v->tail;
    ctrl_record(c, 0, c->n_front, tail, scout, v->bu_edges, 0);
    c->reached += tail; c->reached_deg += scout;
    c->scout_count = scout;                           // :155
    c->n_front = tail;
    c->cur ^= 1;                                      // queue.slide_window(), :156
    c->level++;
    ctrl_decide(c, false);
  } else {
    const long long awake = v->awake;
    ctrl_record(c, 1, c->old_awake, awake, awake, v->bu_edges, v->bu_scanned);
    c->reached += awake;
    c->fb ^= 1;                                       // front.swap(curr), :145
    c->level++;
    if (awake >= c->old_awake || awake > m / kBeta) {  // :148-149
      c->old_awake = awake;
      c->convert = kConvNone;
      c->iter++;
    } else {
      c->n_front = (int)awake;                        // BitmapToQueue, :150 (done at the head of the top-down step)
      c->scout_count = 1;                             // :151
      ctrl_decide(c, true);
    }
  }
  if (c->n_steps > m + 8) { c->aborted = 1; c->mode = kModeDone; }
  cnt->scout = 0; cnt->awake = 0; cnt->degsum = 0; cnt->tail = 0; cnt->bu_edges = 0; cnt->bu_scanned = 0; cnt->heavy_pack = 0;
  cnt->b2q_tail = 0; cnt->ticket = 0;
  __threadfence();
}
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bfs_persist(const OffT *__restrict__ orp, const int32_t *__restrict__ ocol, const OffT *__restrict__ irp,
            const int32_t *__restrict__ bu_col, PersistArgs p) {
  // the phases never overlap inside a CTA: one buffer serves td_heavy's staging and the sweep's lists
  __shared__ __align__(16) unsigned char s_buf[8 * 2048 * sizeof(uint16_t) + 8 * 32 * sizeof(uint32_t)];
  static_assert(sizeof(s_buf) >= 8 * kTdStage * sizeof(int), "staging buffer of td_heavy");
  cg::grid_group grid = cg::this_grid();
  volatile BfsCtrl *vc = p.ctrl;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.ctrl->t_step = (long long)bfs_timer_ns();     // (the first step's clock starts here)
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // a step is over when every CTA has drawn a ticket; the last one runs the controller
  auto step_done = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&p.cnt->ticket, 1u) == gridDim.x - 1) ctrl_after_step(p.ctrl, p.cnt, p.m);
    }
  };
  for (;;) {
    const int mode = vc->mode, convert = vc->convert, cur = vc->cur, fb = vc->fb, n_front = vc->n_front, level = vc->level;
    if (mode == kModeDone) break;
    int32_t *q_cur = cur ? p.queue[1] : p.queue[0], *q_nxt = cur ? p.queue[0] : p.queue[1];
    uint32_t *bm_front = fb ? p.bm[1] : p.bm[0], *bm_next = fb ? p.bm[0] : p.bm[1];
    if (mode == kModeTD) {
      if (convert == kConvB2Q) {
        bitmap_to_queue_dev(bm_front, 0, p.n_words, q_cur, &p.cnt->b2q_tail);
        grid.sync();
      }
      const BfsState bs = {p.visited, p.depth, p.parent, q_nxt, p.cnt, nullptr, 0};
      td_expand_dev<OffT>(orp, ocol, q_cur, n_front, bs, p.heavy_q, p.heavy_off, td_cut(vc->scout_count, p.sm), level + 1);
      grid.sync();
      td_heavy_dev<OffT>(orp, ocol, bs, p.heavy_q, p.heavy_off, level + 1, reinterpret_cast<int *>(s_buf));
      step_done();
      grid.sync();
    } else {
      if (convert == kConvQ2B) {
        for (int64_t w = tid; w < p.n_words; w += nth) bm_front[w] = 0;
        grid.sync();
        queue_to_bitmap_dev(q_cur, n_front, bm_front);
        grid.sync();
      }
      unsigned long long t0 = 0;
      if (tid == 0) t0 = bfs_timer_ns();
      bu_sweep_dev<OffT>(irp, bu_col, p.head, bm_front, bm_next, p.visited, p.depth, p.parent, 0, p.n_words, 0, true, level + 1, p.cnt,
                         reinterpret_cast<uint16_t *>(s_buf), reinterpret_cast<uint32_t *>(s_buf + 8 * 2048 * sizeof(uint16_t)));
      step_done();
      grid.sync();
      if (tid == 0) p.ctrl->bu_ns += (long long)(bfs_timer_ns() - t0);
    }
  }
}

// Sum of out-degree over the reached vertices (the numerator of TEPS): measurement bookkeeping, outside the timed region
// like the reference's own verifier; out[0] += degrees, out[1] += vertices.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bfs_count_reached(const OffT *__restrict__ out_rowptr, const int32_t *__restrict__ depth, int64_t m, unsigned long long *out) {
  const int lane = threadIdx.x & 31;
  unsigned long long deg = 0, cntv = 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < m; v += (int64_t)gridDim.x * blockDim.x)
    if (depth[v] != GDN_INFINITY) { deg += (unsigned long long)(out_rowptr[v + 1] - out_rowptr[v]); cntv++; }
  deg = (unsigned long long)warp_sum((long long)deg);
  cntv = (unsigned long long)warp_sum((long long)cntv);
  if (lane == 0 && cntv) { atomicAdd(out, deg); atomicAdd(out + 1, cntv); }
}

template <typename OffT>
__global__ void bfs_ctrl_init(const OffT *__restrict__ out_rowptr, int source, int64_t nnz, BfsCtrl *c, BfsCounters *cnt) {
  cnt->scout = 0;                                                              // (bfs_init left degrees[source] there for the host-driven path)
  c->edges_to_check = nnz;                                                     // g.E(), omp_beamer.cc:129
  c->scout_count = (long long)(out_rowptr[source + 1] - out_rowptr[source]);   // degrees[source], :130
  c->old_awake = 0; c->reached = 1; c->reached_deg = c->scout_count; c->bu_ns = 0;
  c->n_front = 1; c->cur = 0; c->fb = 0; c->level = 0; c->iter = 0; c->n_steps = 0; c->aborted = 0; c->pad = 0;
  c->mode = kModeTD; c->convert = kConvNone;
  ctrl_decide(c, false);
  c->t_step = 0;
}

template <typename OffT>
static int bfs_t(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st) {
  GDN_CHECK(bfs_alloc(g));
  GDN_CHECK(bfs_prepare<OffT>(g));
  cudaStream_t s = lib().stream;
  const DevCsr &co = g->out;
  const DevCsr &ci = g->symmetric ? g->out : g->in;
  const OffT *orp = (const OffT *)co.rowptr;
  const OffT *irp = (const OffT *)ci.rowptr;
  const int32_t *bu_col = g->col_bu ? g->col_bu : ci.col;
  BfsCounters *cnt = (BfsCounters *)g->counters;
  const int64_t m = g->m;
  const int sm = lib().sm_count;
  if (!g->bfs_ctrl) {
    GDN_CUDA(cudaMalloc((void **)&g->bfs_ctrl, sizeof(BfsCtrl)));
    GDN_CUDA(cudaHostAlloc((void **)&g->bfs_ctrl_host, sizeof(BfsCtrl), cudaHostAllocDefault));
  }
  BfsCtrl *ctrl = (BfsCtrl *)g->bfs_ctrl, *h = (BfsCtrl *)g->bfs_ctrl_host;
  if (!g->bfs_reached) GDN_CUDA(cudaMalloc((void **)&g->bfs_reached, 2 * sizeof(unsigned long long)));
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    int a = 0, b = 0;
    GDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, bfs_persist<uint32_t>, 256, 0));
    GDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, bfs_persist<uint64_t>, 256, 0));
    ctas_per_sm = std::max(1, std::min(4, std::min(a, b)));
  }

  // untimed, like the reference's own depth / bitmap initialisation (omp_beamer.cc:116-131 sit before t.Start())
  const int init_grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm * 8);
  bfs_init<OffT><<<init_grid, 256, 0, s>>>(orp, d_depth, d_parent, g->visited, m, g->n_words, source, g->queue[0], cnt,
                                           nullptr, 0, m, g->iso);
  bfs_ctrl_init<OffT><<<1, 1, 0, s>>>(orp, source, (int64_t)co.nnz, ctrl, cnt);

  PersistArgs pa;
  pa.queue[0] = g->queue[0]; pa.queue[1] = g->queue[1];
  pa.bm[0] = g->front; pa.bm[1] = g->next; pa.head = g->bu_head;
  pa.visited = g->visited; pa.depth = d_depth; pa.parent = d_parent;
  pa.heavy_q = g->heavy_queue; pa.heavy_off = g->heavy_off;
  pa.cnt = cnt; pa.ctrl = ctrl; pa.m = m; pa.n_words = g->n_words; pa.sm = sm;
  const int32_t *ocol = co.col;
  void *args[] = {(void *)&orp, (void *)&ocol, (void *)&irp, (void *)&bu_col, (void *)&pa};
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  GDN_CUDA(cudaLaunchCooperativeKernel((const void *)bfs_persist<OffT>, dim3(sm * ctas_per_sm), dim3(256), args, 0, s));
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaMemcpyAsync(h, ctrl, sizeof(BfsCtrl), cudaMemcpyDeviceToHost, s));
  unsigned long long h_reached[2] = {0, 0};
  if (st) {                                        // TEPS numerator: after the clock has stopped
    GDN_CUDA(cudaMemsetAsync(g->bfs_reached, 0, 2 * sizeof(unsigned long long), s));
    bfs_count_reached<OffT><<<sm * 8, 256, 0, s>>>(orp, d_depth, m, g->bfs_reached);
    GDN_CUDA(cudaMemcpyAsync(h_reached, g->bfs_reached, sizeof(h_reached), cudaMemcpyDeviceToHost, s));
  }
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (h->aborted) { set_error("BFS controller did not terminate"); return GDN_ERR_CUDA; }
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms;
    st->iterations = h->iter;
    st->n_steps = h->n_steps;
    st->kernel_launches = 3;
    st->edges_reached = (int64_t)h_reached[0];
    st->vertices_reached = (int64_t)h_reached[1];
    st->kernel_ms = (double)h->bu_ns * 1e-6;       // bottom-up sweeps, timed inside the kernel (globaltimer)
    st->kernel_calls = 0;
    for (int i = 0; i < std::min(h->n_steps, GDN_MAX_BFS_STEPS); i++) { st->steps[i] = h->steps[i]; st->kernel_calls += h->steps[i].dir; }
  }
  return GDN_OK;
}

int bfs_merge_or(gdn_graph *g, uint32_t *bm, uint32_t *xbuf);          // comm.cu
int bfs_allgather_words(gdn_graph *g, uint32_t *bm);                   // comm.cu
int allreduce_i64(long long *d_p, int n);                              // comm.cu

// Partitioned mode: parents of the vertices this GPU owns that were discovered by a top-down step (their claim went
// through the merged bitmap, which carries no source).  Depths are replicated, so any in-neighbour one level up is a valid
// parent (the reference keeps parents only in comments, src/bfs/omp_beamer.cc:12,18,22,44,47; Graph500 accepts any tree
// edge).  One warp per 32 owned vertices; a vertex that still needs a parent has its row scanned by the whole warp.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bfs_parents_from_depth(const OffT *__restrict__ in_rowptr, const int32_t *__restrict__ in_col, const int32_t *__restrict__ depth,
                       int32_t *parent, int64_t row_lo, int64_t row_hi) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v0 = row_lo + warp * 32; v0 < row_hi; v0 += nwarps * 32) {
    const int64_t v = v0 + lane;
    int d = -1;
    if (v < row_hi) { d = depth[v]; if (d == GDN_INFINITY || d == 0 || parent[v] >= 0) d = -1; }
    unsigned need = __ballot_sync(kFull, d > 0);
    while (need) {
      const int l = __ffs(need) - 1;
      need &= need - 1;
      const int64_t vv = v0 + l;
      const int want = __shfl_sync(kFull, d, l) - 1;
      const OffT b = in_rowptr[vv - row_lo], e = in_rowptr[vv - row_lo + 1];
      int found = -1;
      for (OffT i = b; i < e && found < 0; i += 32) {
        const OffT k = i + lane;
        int src = -1;
        bool ok = false;
        if (k < e) { src = in_col[k]; ok = depth[src] == want; }
        const unsigned bal = __ballot_sync(kFull, ok);
        if (bal) found = __shfl_sync(kFull, src, __ffs(bal) - 1);
      }
      if (lane == 0) parent[vv] = found;
    }
  }
}
__global__ void fill_i32(int32_t *p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void set_i32(int32_t *p, int32_t v) { *p = v; }

int allgather_i32(int32_t *buf, int64_t per_rank);                     // comm.cu (in place)

// 1-D row-partitioned BFS.  Every GPU holds its rows of the CSR and
// FULL-length visited/frontier bitmaps and depth[]; all GPUs run the identical
// alpha/beta controller on all-reduced counts, so they take the same branch.
//   bottom-up : sweep own rows against the global frontier -> own slice of `next`
//               -> ONE allgather of the bitmap slices
//   top-down  : owners expand their frontier rows and mark (possibly remote)
//               destinations in a full-length bitmap -> slices are sent to their
//               owners and OR-ed (grouped send/recv), then ONE allgather
//   absorb    : every GPU claims the merged new vertices (visited, depth) and the
//               owners add up their out-degrees -> allreduce(scout_count)
template <typename OffT>
static int bfs_multi_t(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st) {
  GDN_CHECK(bfs_alloc(g));
  GDN_CHECK(bfs_prepare<OffT>(g));
  cudaStream_t s = lib().stream;
  const DevCsr &co = g->out;
  const DevCsr &ci = g->symmetric ? g->out : g->in;
  const OffT *orp = (const OffT *)co.rowptr;
  const OffT *irp = (const OffT *)ci.rowptr;
  BfsCounters *cnt = (BfsCounters *)g->counters;
  BfsCounters *h = (BfsCounters *)lib().pinned;
  const int64_t m = g->m;
  const int sm = lib().sm_count;
  const int P = comm_size(), R = comm_rank();
  const int64_t ww = partition_width(m, P) / 32;                 // words per slice (multiple of 32)
  const int64_t own_lo = std::min<int64_t>((int64_t)R * ww, g->n_words);
  const int64_t own_hi = std::min<int64_t>(own_lo + ww, g->n_words);
  if (!g->xbuf) {
    GDN_CUDA(cudaMalloc((void **)&g->xbuf, sizeof(uint32_t) * ww * P));
    g->device_bytes += sizeof(uint32_t) * ww * P;
  }
  int64_t launches = 0;
  uint32_t *front = g->front, *next = g->next;
  // parents: a global-indexed scratch of P equal slices (the allgather's shape); bottom-up steps write the owned rows
  // directly, top-down discoveries are resolved from the depths at the end
  const int64_t W = partition_width(m, P);
  int32_t *pbuf = nullptr;
  if (d_parent) {
    if (!g->parent_buf) { GDN_CUDA(cudaMalloc((void **)&g->parent_buf, sizeof(int32_t) * W * P)); g->device_bytes += sizeof(int32_t) * W * P; }
    pbuf = g->parent_buf;
    fill_i32<<<sm * 8, 256, 0, s>>>(pbuf, W * P, -1);
    if (source >= g->row_lo && source < g->row_hi) set_i32<<<1, 1, 0, s>>>(pbuf + source, source);
  }

  const int init_grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm * 8);
  bfs_init<OffT><<<init_grid, 256, 0, s>>>(orp, d_depth, nullptr, g->visited, m, g->n_words, source, g->queue[0], cnt,
                                           front, g->row_lo, g->row_hi, g->iso);
  // global E and degrees[source]: two all-reduced scalars
  long long *d_tmp = (long long *)&cnt->awake;                  // reuse: awake <- local nnz
  long long local_nnz = (long long)co.nnz;
  GDN_CUDA(cudaMemcpyAsync(d_tmp, &local_nnz, sizeof(long long), cudaMemcpyHostToDevice, s));
  GDN_CHECK(allreduce_i64(&cnt->scout, 2));                     // {scout, awake} are adjacent
  GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());

  int64_t edges_to_check = h->awake;                  // g.E(), omp_beamer.cc:129
  int64_t scout_count = h->scout;                     // degrees[source], :130
  int64_t reached_deg = scout_count, reached = 1;
  int64_t n_front = 1;
  int level = 0, iter = 0, n_steps = 0;
  auto record = [&](int dir, int64_t frontier, int64_t disc, int64_t sc, int64_t edges, int64_t scanned) {
    if (st && n_steps < GDN_MAX_BFS_STEPS) {
      gdn_bfs_step &b = st->steps[n_steps];
      b.dir = dir; b.ns = 0; b.frontier = frontier; b.discovered = disc; b.scout = sc; b.edges = edges; b.scanned = scanned;
    }
    n_steps++;
  };
  const int own_grid = (int)std::max<int64_t>(1, std::min<int64_t>(((own_hi - own_lo) / 32 + 7) / 8, (int64_t)sm * 8));
  const int all_grid = (int)std::max<int64_t>(1, std::min<int64_t>((g->n_words + 255) / 256, (int64_t)sm * 8));
  auto absorb = [&]() -> int {
    bfs_absorb<OffT><<<all_grid, 256, 0, s>>>(next, g->visited, front, d_depth, orp, g->n_words, own_lo, own_hi,
                                              g->row_lo, level + 1, cnt);
    launches++;
    GDN_CHECK(allreduce_i64(&cnt->scout, 1));
    GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
    GDN_CUDA(cudaStreamSynchronize(s));
    return GDN_OK;
  };

  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  while (n_front > 0) {                                           // omp_beamer.cc:135
    if (scout_count > edges_to_check / kAlpha) {                  // :136
      int64_t awake = n_front, old_awake;
      do {
        ++iter;
        old_awake = awake;
        GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
        kev_begin();
        bu_sweep<OffT><<<own_grid, 256, 0, s>>>(irp, g->col_bu ? g->col_bu : ci.col, g->bu_head, front, next, g->visited, nullptr, pbuf, own_lo,
                                                own_hi, g->row_lo, false, level + 1, cnt);
        kev_end();
        launches++;
        GDN_CHECK(bfs_allgather_words(g, next));
        GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));   // absorb recounts globally
        GDN_CHECK(absorb());
        awake = h->awake;
        reached += awake; reached_deg += h->scout;
        level++;
        record(1, old_awake, awake, awake, 0, 0);
      } while ((awake >= old_awake) || (awake > m / kBeta));      // :148-149
      n_front = awake;
      scout_count = 1;                                            // :151
    } else {
      ++iter;
      edges_to_check -= scout_count;                              // :154
      GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
      GDN_CUDA(cudaMemsetAsync(next, 0, sizeof(uint32_t) * g->n_words, s));
      bitmap_to_queue<<<own_grid, 256, 0, s>>>(front, own_lo, own_hi, g->queue[0], cnt);
      BfsState bs = {g->visited, d_depth, nullptr, nullptr, cnt, next, g->row_lo};
      td_expand<OffT><<<sm * 4, 256, 0, s>>>(orp, co.col, g->queue[0], -1, bs, g->heavy_queue, g->heavy_off,
                                             td_cut(scout_count / P, sm), level + 1);
      td_heavy<OffT><<<sm * 4, 256, 0, s>>>(orp, co.col, bs, g->heavy_queue, g->heavy_off, level + 1);
      launches += 3;
      GDN_CHECK(bfs_merge_or(g, next, g->xbuf));
      GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
      GDN_CHECK(absorb());
      scout_count = h->scout;                                     // :155
      record(0, n_front, h->awake, scout_count, 0, 0);
      reached += h->awake; reached_deg += scout_count;
      n_front = h->awake;
      level++;
    }
  }
  if (d_parent) {
    bfs_parents_from_depth<OffT><<<sm * 8, 256, 0, s>>>(irp, ci.col, d_depth, pbuf, g->row_lo, g->row_hi);
    launches++;
    GDN_CHECK(allgather_i32(pbuf, W));
    GDN_CUDA(cudaMemcpyAsync(d_parent, pbuf, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, s));
  }
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms; st->iterations = iter; st->n_steps = n_steps; st->kernel_launches = launches;
    st->edges_reached = reached_deg; st->vertices_reached = reached;
    kev_collect(st);
  }
  return GDN_OK;
}

int bfs_run(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st) {
  if (!g->has_out || !(g->symmetric || g->has_in)) {
    // src/bfs/omp_beamer.cc:98-102
    set_error("This algorithm requires the reverse graph constructed for directed graph");
    return GDN_ERR_GRAPH;
  }
  if (source < 0 || source >= g->m) { set_error("source out of range"); return GDN_ERR_ARG; }
  if (g->row_lo != 0 || g->row_hi != g->m || comm_size() > 1) {
    const int64_t w = partition_width(g->m, comm_size());
    if (g->row_lo != std::min<int64_t>((int64_t)comm_rank() * w, g->m) || g->row_hi != std::min<int64_t>(g->row_lo + w, g->m)) {
      set_error("partitioned BFS: graph rows [%lld,%lld) do not match gdn_partition_rows for rank %d of %d",
                (long long)g->row_lo, (long long)g->row_hi, comm_rank(), comm_size());
      return GDN_ERR_ARG;
    }
    return g->out.off64 ? bfs_multi_t<uint64_t>(g, source, d_depth, d_parent, st) : bfs_multi_t<uint32_t>(g, source, d_depth, d_parent, st);
  }
  return g->out.off64 ? bfs_t<uint64_t>(g, source, d_depth, d_parent, st)
                      : bfs_t<uint32_t>(g, source, d_depth, d_parent, st);
}

}  // namespace gdn
