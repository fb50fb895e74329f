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

namespace gdn {

constexpr int kTdHeavy = 8192;     // rows at least this long are expanded by the whole grid
constexpr int kBuSerial = 8;       // in-neighbours a lane probes alone before the warp helps
constexpr int kAlpha = 15, kBeta = 18;   // src/bfs/omp_beamer.cc:111

struct BfsCounters {
  long long scout;      // TD: sum of out-degree of newly claimed vertices (omp_beamer.cc:50)
  long long awake;      // BU: vertices discovered (omp_beamer.cc:23)
  long long degsum;     // BU: sum of out-degree of discovered vertices (TEPS accounting)
  int tail;             // next-queue length
  int heavy_tail;       // heavy-queue length
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

// TDStep, src/bfs/omp_beamer.cc:35-58.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
td_expand(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ q_in,
          int n_in, BfsState s, int32_t *heavy_q, int level) {
  if (n_in < 0) n_in = s.cnt->tail;       // partitioned mode: queue length lives on the device
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  long long scout = 0;
  for (int base = warp * 32; base < n_in; base += nwarps * 32) {
    const int idx = base + lane;
    int v = -1;
    OffT b = 0, e = 0;
    if (idx < n_in) { v = q_in[idx]; b = rowptr[v - s.row_lo]; e = rowptr[v - s.row_lo + 1]; }
    uint32_t deg = (uint32_t)(e - b);
    if (deg >= (uint32_t)kTdHeavy) {                    // tier 1: defer to td_heavy
      heavy_q[atomicAdd(&s.cnt->heavy_tail, 1)] = v;
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
  if (lane == 0 && scout) atomicAdd((unsigned long long *)&s.cnt->scout, (unsigned long long)scout);
}

// Rows >= kTdHeavy: every warp of the grid takes 32-edge pieces of the row.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
td_heavy(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, BfsState s,
         const int32_t *__restrict__ heavy_q, int level) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nh = s.cnt->heavy_tail;
  long long scout = 0;
  for (int h = 0; h < nh; h++) {
    const int src = heavy_q[h];
    const OffT b = rowptr[src - s.row_lo], e = rowptr[src - s.row_lo + 1];
    for (OffT i = b + (OffT)warp * 32; i < e; i += (OffT)nwarps * 32) {
      const OffT k = i + lane;
      const int dst = (k < e) ? col[k] : -1;
      scout += td_visit(rowptr, s, dst, src, level, lane);
    }
  }
  scout = warp_sum(scout);
  if (lane == 0 && scout) atomicAdd((unsigned long long *)&s.cnt->scout, (unsigned long long)scout);
}

// BUStep, src/bfs/omp_beamer.cc:13-32.  One warp sweeps 32 bitmap words
// (1024 vertices) at a time; lane L owns vertex bit L of the current word.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
bu_sweep(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const OffT *__restrict__ out_rowptr,
         const uint32_t *__restrict__ front, uint32_t *__restrict__ next, uint32_t *visited, int32_t *depth,
         int32_t *parent, int64_t word_lo, int64_t word_hi, int64_t row_lo, bool update_visited, int level,
         BfsCounters *cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t g_lo = word_lo >> 5, g_hi = word_hi >> 5;   // word ranges are multiples of 32
  rowptr -= row_lo;                                          // rows are addressed by global vertex id
  out_rowptr -= row_lo;
  long long awake = 0, degsum = 0;
  for (int64_t g = g_lo + warp; g < g_hi; g += nwarps) {
    const int64_t widx = g * 32 + lane;
    const uint32_t vis = visited[widx];
    uint32_t nxt = 0;
    if (__ballot_sync(kFull, vis != 0xffffffffu) != 0) {
      for (int w = 0; w < 32; w++) {
        const uint32_t word = __shfl_sync(kFull, vis, w);
        if (word == 0xffffffffu) continue;               // warp-uniform
        const int64_t v = (g * 32 + w) * 32 + lane;
        const bool active = !((word >> lane) & 1u);
        bool found = false;
        int par = -1;
        OffT lim = 0, e = 0;
        if (active) {
          const OffT b = rowptr[v];
          e = rowptr[v + 1];
          lim = (e - b > (OffT)kBuSerial) ? b + kBuSerial : e;
          for (OffT i = b; i < lim; i++) {
            const int src = col[i];
            if ((front[(uint32_t)src >> 5] >> (src & 31)) & 1u) { found = true; par = src; break; }
          }
        }
        // long rows that have not hit yet: the whole warp scans the remainder
        unsigned rest = __ballot_sync(kFull, active && !found && lim < e);
        while (rest) {
          const int l = __ffs(rest) - 1;
          rest &= rest - 1;
          const OffT bb = __shfl_sync(kFull, lim, l), ee = __shfl_sync(kFull, e, l);
          int hit = -1;
          for (OffT i = bb; i < ee; i += 32) {
            const OffT k = i + lane;
            int src = -1;
            bool ok = false;
            if (k < ee) { src = col[k]; ok = (front[(uint32_t)src >> 5] >> (src & 31)) & 1u; }
            const unsigned bal = __ballot_sync(kFull, ok);
            if (bal) { hit = __shfl_sync(kFull, src, __ffs(bal) - 1); break; }
          }
          if (lane == l && hit >= 0) { found = true; par = hit; }
        }
        const unsigned fmask = __ballot_sync(kFull, found);
        if (found) {
          depth[v] = level;
          if (parent) parent[v] = par;
          degsum += (long long)(out_rowptr[v + 1] - out_rowptr[v]);
        }
        if (lane == w) nxt = fmask;
        if (lane == 0) awake += __popc(fmask);
      }
    }
    next[widx] = nxt;
    if (nxt && update_visited) visited[widx] = vis | nxt;
  }
  awake = warp_sum(awake);
  degsum = warp_sum(degsum);
  if (lane == 0 && awake) {
    atomicAdd((unsigned long long *)&cnt->awake, (unsigned long long)awake);
    atomicAdd((unsigned long long *)&cnt->degsum, (unsigned long long)degsum);
  }
}

// QueueToBitmap, src/bfs/omp_beamer.cc:60-67
__global__ void queue_to_bitmap(const int32_t *__restrict__ q, int n, uint32_t *bm) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int v = q[i];
    atomicOr(&bm[(uint32_t)v >> 5], 1u << (v & 31));
  }
}

// BitmapToQueue, src/bfs/omp_beamer.cc:69-79: ballot-free popc compaction, one
// atomicAdd per 1024 vertices.
__global__ void __launch_bounds__(256, 4)
bitmap_to_queue(const uint32_t *__restrict__ bm, int64_t word_lo, int64_t word_hi, int32_t *q, BfsCounters *cnt) {
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
    if (lane == 0) base = atomicAdd(&cnt->tail, total);
    base = __shfl_sync(kFull, base, 0) + off - c;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      q[base++] = (int32_t)(widx * 32 + bit);
    }
  }
}

template <typename OffT>
__global__ void bfs_init(const OffT *__restrict__ out_rowptr, int32_t *depth, int32_t *parent, uint32_t *visited,
                         int64_t m, int64_t n_words, int source, int32_t *queue0, BfsCounters *cnt,
                         uint32_t *front /* partitioned mode: bitmap holding just the source */, int64_t row_lo,
                         int64_t row_hi) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = tid; v < m; v += nth) {
    depth[v] = (v == source) ? 0 : GDN_INFINITY;
    if (parent) parent[v] = (v == source) ? source : -1;
  }
  for (int64_t w = tid; w < n_words; w += nth) {
    uint32_t word = 0;
    const int64_t lo = w * 32;
    if (lo + 32 > m) word = (lo >= m) ? 0xffffffffu : ~((1u << (int)(m - lo)) - 1u);   // pad bits: "visited"
    if ((int64_t)(source >> 5) == w) word |= 1u << (source & 31);
    visited[w] = word;
    if (front) front[w] = ((int64_t)(source >> 5) == w) ? (1u << (source & 31)) : 0u;
  }
  if (tid == 0) {
    queue0[0] = source;
    const bool own = source >= row_lo && source < row_hi;
    cnt->scout = own ? (long long)(out_rowptr[source - row_lo + 1] - out_rowptr[source - row_lo]) : 0;   // degrees[source], omp_beamer.cc:130
    cnt->awake = 0; cnt->degsum = 0; cnt->tail = 0; cnt->heavy_tail = 0;
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
    cudaFree(g->visited); cudaFree(g->front); cudaFree(g->next); cudaFree(g->xbuf);
    g->visited = g->front = g->next = g->xbuf = nullptr;
  }
  const size_t bm = sizeof(uint32_t) * need;
  GDN_CUDA(cudaMalloc((void **)&g->visited, bm));
  GDN_CUDA(cudaMalloc((void **)&g->front, bm));
  GDN_CUDA(cudaMalloc((void **)&g->next, bm));
  GDN_CUDA(cudaMemsetAsync(g->front, 0, bm, lib().stream));
  GDN_CUDA(cudaMemsetAsync(g->next, 0, bm, lib().stream));
  g->bm_alloc_words = need;
  g->device_bytes += 3 * bm;
  if (!g->queue[0]) {
    GDN_CUDA(cudaMalloc((void **)&g->queue[0], sizeof(int32_t) * std::max<int64_t>(g->m, 1)));
    GDN_CUDA(cudaMalloc((void **)&g->queue[1], sizeof(int32_t) * std::max<int64_t>(g->m, 1)));
    const int64_t hq = (int64_t)(g->out.nnz / kTdHeavy) + 2;
    GDN_CUDA(cudaMalloc((void **)&g->heavy_queue, sizeof(int32_t) * hq));
    GDN_CUDA(cudaMalloc((void **)&g->counters, sizeof(BfsCounters)));
    g->device_bytes += 2 * sizeof(int32_t) * g->m + sizeof(int32_t) * hq + sizeof(BfsCounters);
  }
  return GDN_OK;
}

template <typename OffT>
static int bfs_t(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st) {
  GDN_CHECK(bfs_alloc(g));
  cudaStream_t s = lib().stream;
  const DevCsr &co = g->out;
  const DevCsr &ci = g->symmetric ? g->out : g->in;
  const OffT *orp = (const OffT *)co.rowptr;
  const OffT *irp = (const OffT *)ci.rowptr;
  BfsCounters *cnt = (BfsCounters *)g->counters;
  BfsCounters *h = (BfsCounters *)lib().pinned;
  const int64_t m = g->m;
  const int sm = lib().sm_count;
  int64_t launches = 0;

  const int init_grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm * 8);
  bfs_init<OffT><<<init_grid, 256, 0, s>>>(orp, d_depth, d_parent, g->visited, m, g->n_words, source, g->queue[0], cnt,
                                           nullptr, 0, m);
  GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());

  int64_t edges_to_check = (int64_t)co.nnz;          // g.E(), omp_beamer.cc:129
  int64_t scout_count = h->scout;                    // degrees[source], :130
  int64_t reached_deg = scout_count, reached = 1;
  int64_t n_in = 1;
  int cur = 0, level = 0, iter = 0, n_steps = 0;
  uint32_t *front = g->front, *next = g->next;
  auto record = [&](int dir, int64_t frontier, int64_t disc, int64_t sc) {
    if (st && n_steps < GDN_MAX_BFS_STEPS) {
      gdn_bfs_step &b = st->steps[n_steps];
      b.dir = dir; b.pad = 0; b.frontier = frontier; b.discovered = disc; b.scout = sc;
    }
    n_steps++;
  };
  const int sweep_grid = (int)std::max<int64_t>(1, std::min<int64_t>((g->n_words / 32 + 7) / 8, (int64_t)sm * 8));

  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  while (n_in > 0) {                                              // omp_beamer.cc:135
    if (scout_count > edges_to_check / kAlpha) {                  // :136
      GDN_CUDA(cudaMemsetAsync(front, 0, sizeof(uint32_t) * g->n_words, s));
      queue_to_bitmap<<<(int)std::min<int64_t>((n_in + 255) / 256, sm * 8), 256, 0, s>>>(g->queue[cur], (int)n_in, front);
      launches++;
      int64_t awake = n_in, old_awake;                            // :139
      do {
        ++iter;
        old_awake = awake;
        GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
        kev_begin();
        bu_sweep<OffT><<<sweep_grid, 256, 0, s>>>(irp, ci.col, orp, front, next, g->visited, d_depth, d_parent,
                                                  0, g->n_words, 0, true, level + 1, cnt);
        kev_end();
        launches++;
        GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
        GDN_CUDA(cudaStreamSynchronize(s));
        awake = h->awake;
        reached += awake; reached_deg += h->degsum;
        level++;
        std::swap(front, next);                                   // front.swap(curr), :145
        record(1, old_awake, awake, awake);
      } while ((awake >= old_awake) || (awake > m / kBeta));      // :148-149
      GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
      bitmap_to_queue<<<sweep_grid, 256, 0, s>>>(front, 0, g->n_words, g->queue[cur], cnt);
      launches++;
      GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
      GDN_CUDA(cudaStreamSynchronize(s));
      n_in = h->tail;
      scout_count = 1;                                            // :151
    } else {
      ++iter;
      edges_to_check -= scout_count;                              // :154
      GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
      BfsState bs = {g->visited, d_depth, d_parent, g->queue[cur ^ 1], cnt, nullptr, 0};
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_in + 255) / 256, (int64_t)sm * 8));
      td_expand<OffT><<<grid, 256, 0, s>>>(orp, co.col, g->queue[cur], (int)n_in, bs, g->heavy_queue, level + 1);
      td_heavy<OffT><<<sm * 4, 256, 0, s>>>(orp, co.col, bs, g->heavy_queue, level + 1);
      launches += 2;
      GDN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(BfsCounters), cudaMemcpyDeviceToHost, s));
      GDN_CUDA(cudaStreamSynchronize(s));
      scout_count = h->scout;                                     // :155
      record(0, n_in, h->tail, scout_count);
      reached += h->tail; reached_deg += scout_count;
      n_in = h->tail;
      cur ^= 1;
      level++;
    }
  }
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms;
    st->iterations = iter;
    st->n_steps = n_steps;
    st->kernel_launches = launches;
    st->edges_reached = reached_deg;
    st->vertices_reached = reached;
    kev_collect(st);
  }
  return GDN_OK;
}

int bfs_merge_or(gdn_graph *g, uint32_t *bm, uint32_t *xbuf);          // comm.cu
int bfs_allgather_words(gdn_graph *g, uint32_t *bm);                   // comm.cu
int allreduce_i64(long long *d_p, int n);                              // comm.cu

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
static int bfs_multi_t(gdn_graph *g, int32_t source, int32_t *d_depth, gdn_stats *st) {
  GDN_CHECK(bfs_alloc(g));
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

  const int init_grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm * 8);
  bfs_init<OffT><<<init_grid, 256, 0, s>>>(orp, d_depth, nullptr, g->visited, m, g->n_words, source, g->queue[0], cnt,
                                           front, g->row_lo, g->row_hi);
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
  auto record = [&](int dir, int64_t frontier, int64_t disc, int64_t sc) {
    if (st && n_steps < GDN_MAX_BFS_STEPS) {
      gdn_bfs_step &b = st->steps[n_steps];
      b.dir = dir; b.pad = 0; b.frontier = frontier; b.discovered = disc; b.scout = sc;
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
        bu_sweep<OffT><<<own_grid, 256, 0, s>>>(irp, ci.col, orp, front, next, g->visited, d_depth, nullptr, own_lo,
                                                own_hi, g->row_lo, false, level + 1, cnt);
        kev_end();
        launches++;
        GDN_CHECK(bfs_allgather_words(g, next));
        GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));   // absorb recounts globally
        GDN_CHECK(absorb());
        awake = h->awake;
        reached += awake; reached_deg += h->scout;
        level++;
        record(1, old_awake, awake, awake);
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
      td_expand<OffT><<<sm * 4, 256, 0, s>>>(orp, co.col, g->queue[0], -1, bs, g->heavy_queue, level + 1);
      td_heavy<OffT><<<sm * 4, 256, 0, s>>>(orp, co.col, bs, g->heavy_queue, level + 1);
      launches += 3;
      GDN_CHECK(bfs_merge_or(g, next, g->xbuf));
      GDN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(BfsCounters), s));
      GDN_CHECK(absorb());
      scout_count = h->scout;                                     // :155
      record(0, n_front, h->awake, scout_count);
      reached += h->awake; reached_deg += scout_count;
      n_front = h->awake;
      level++;
    }
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
    if (d_parent) { set_error("partitioned BFS does not return parents yet"); return GDN_ERR_ARG; }
    return g->out.off64 ? bfs_multi_t<uint64_t>(g, source, d_depth, st) : bfs_multi_t<uint32_t>(g, source, d_depth, st);
  }
  return g->out.off64 ? bfs_t<uint64_t>(g, source, d_depth, d_parent, st)
                      : bfs_t<uint32_t>(g, source, d_depth, d_parent, st);
}

}  // namespace gdn
