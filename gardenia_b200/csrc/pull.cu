// pull.cu -- PageRank pull on a degree-sorted SELL-32 layout with a shared-memory
// hot-vertex table.
//
// Why (measured on this B200, profiles/r1_gather_microbench_b200.txt): a random
// 4-byte gather sustains ~1.0 gather/clk/SM out of L2, ~3/clk/SM out of L1,
// >=5/clk/SM out of shared memory, and only 40-74 G/s (0.15-0.25/clk/SM) once the
// gathered vector no longer fits L2 (268 MB at Kronecker scale 26, whose ids the
// generator permutes at random, include/generator.h:52-62).  The pull gather of
// src/pr/omp_base.cc:27-33 is therefore bound by WHERE contrib[src] lives, not by
// the colidx stream.  This file re-lays the problem out for the memory hierarchy
// once per graph (untimed, like the reference's own segmenting/tiling variants,
// include/segmenting.h) and never changes the arithmetic:
//
//  * vertices are renumbered by (out-)degree, hottest first (host-side stable
//    counting sort, O(m)); the top H (<= 48 K) contrib values are kept in a
//    192 KB shared-memory table per SM, the next few million stay L2-resident
//    because they are now contiguous;
//  * rows are sorted by length and stored as SELL-32 slices (32 rows per warp,
//    lane = row, 4 column ids per lane per 128-bit coalesced load), so a lane
//    sums ITS row in the reference's sequential fp32 order -- rows up to 128
//    non-zeros are bit-identical to pr_omp_base -- with no staging, no
//    divergence (rows of a slice have (nearly) equal length) and 8 independent
//    gathers in flight per lane;
//  * slices wider than 128 are cut into 128-column segments (one warp each) whose
//    per-row partials are added in column order by pr_sell_finalize;
//  * scores live in sorted order during the solve and are permuted back at the end.
//
// Multi-GPU: every rank sorts ITS rows; new ids are laid out as
//   [ hot slice of rank 0 | ... | hot slice of rank P-1 | cold slice of rank 0 | ... ]
// so the hot table is one contiguous prefix and each rank still owns two
// contiguous slices of contrib (two in-place NCCL allgathers per iteration).
#include "pull.cuh"
#include "ordered_sum.cuh"
#include <omp.h>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <vector>

namespace gdn {

int comm_size();
int comm_rank();
int64_t partition_width(int64_t m, int nparts);

// ------------------------------------------------------------------ host: stable sort by degree, descending
// perm[j] = index (relative to lo) of the j-th vertex of [lo,hi) in (degree desc, id asc) order.
// 268 MB work arrays: uninitialised (std::vector would zero them on ONE thread: ~60 ms each at Kron-26) and
// first-touched inside the parallel loops.
template <typename T>
struct RawBuf {
  T *p = nullptr;
  explicit RawBuf(size_t n = 0) { if (n) p = (T *)malloc(n * sizeof(T)); }
  ~RawBuf() { free(p); }
  RawBuf(const RawBuf &) = delete;
  RawBuf &operator=(const RawBuf &) = delete;
  void alloc(size_t n) { free(p); p = (T *)malloc(std::max<size_t>(n, 1) * sizeof(T)); }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
  T *data() { return p; }
};

static void sort_by_degree(const int32_t *deg, int64_t lo, int64_t hi, int32_t *perm, int32_t *sorted_deg = nullptr) {
  const int64_t n = hi - lo;
  const int DB = 4096;                   // degrees below DB: parallel counting sort; above: std::stable_sort
  const int T = std::max(1, omp_get_max_threads());
  std::vector<std::vector<uint32_t>> hist(T, std::vector<uint32_t>(DB, 0));
  std::vector<std::vector<int32_t>> big(T);
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    const int64_t a = n * t / T, b = n * (t + 1) / T;
    for (int64_t i = a; i < b; i++) {
      const int32_t d = deg[lo + i];
      if (d >= DB) big[t].push_back((int32_t)i); else hist[t][d]++;
    }
  }
  std::vector<int32_t> bigs;
  for (int t = 0; t < T; t++) bigs.insert(bigs.end(), big[t].begin(), big[t].end());
  std::stable_sort(bigs.begin(), bigs.end(), [&](int32_t x, int32_t y) { return deg[lo + x] > deg[lo + y]; });
  std::copy(bigs.begin(), bigs.end(), perm);
  if (sorted_deg) for (size_t i = 0; i < bigs.size(); i++) sorted_deg[i] = deg[lo + bigs[i]];
  // start[t][d]: descending degree, then thread order (= ascending id)
  uint64_t pos = bigs.size();
  std::vector<std::vector<uint64_t>> start(T, std::vector<uint64_t>(DB));
  for (int d = DB - 1; d >= 0; d--)
    for (int t = 0; t < T; t++) { start[t][d] = pos; pos += hist[t][d]; }
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    const int64_t a = n * t / T, b = n * (t + 1) / T;
    for (int64_t i = a; i < b; i++) {
      const int32_t d = deg[lo + i];
      if (d < DB) {
        const uint64_t pos = start[t][d]++;
        perm[pos] = (int32_t)i;
        if (sorted_deg) sorted_deg[pos] = d;
      }
    }
  }
}

template <typename T>
static int upload(gdn_graph *g, T **dptr, const T *h, size_t n) {
  GDN_CUDA(cudaMalloc((void **)dptr, std::max<size_t>(n, 4) * sizeof(T)));
  g->device_bytes += n * sizeof(T);
  GDN_CUDA(cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, lib().stream));
  return GDN_OK;
}

// Host part of the layout: orders, new ids, slice pointers, work items.  Runs in
// gdn_graph_create while the caller's host offsets are still valid.
//   row_off : offsets of the PULL (in-) CSR, global, host
//   key_off : offsets whose row lengths rank the COLUMNS by hotness (out-CSR; == row_off when symmetric)
// device halves of the layout (one GPU, symmetric order): new ids and sorted row lengths from the order
__global__ void newid_from_perm(const int32_t *__restrict__ perm, int32_t *__restrict__ newid, int64_t n) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) newid[perm[j]] = (int32_t)j;
}
template <typename OffT>
__global__ void sdeg_from_perm(const OffT *__restrict__ rowptr, const int32_t *__restrict__ perm, int32_t *__restrict__ sdeg, int64_t n) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = perm[j];
    sdeg[j] = (int32_t)(rowptr[r + 1] - rowptr[r]);
  }
}

static int pull_stream_fill(gdn_graph *g);     // below, next to the kernel it launches

template <typename HostOffT>
int pull_prepare(gdn_graph *g, const HostOffT *row_off, const HostOffT *key_off) {
  PullLayout &L = g->pull;
  const int P = comm_size(), R = comm_rank();
  const int64_t m = g->m, lo = g->row_lo, hi = g->row_hi, rows = hi - lo;
  const int64_t W = (P == 1) ? m : partition_width(m, P);
  if (lo != std::min<int64_t>((int64_t)R * W, m) || hi != std::min<int64_t>(lo + W, m)) {
    // a row range that is not this rank's gdn_partition_rows slice: keep the plain CSR path
    L.prepared = false;
    return GDN_OK;
  }
  L.P = P; L.R = R; L.W = W;
  L.Hp = std::min<int64_t>(kHotMax / P, W);
  L.H = L.Hp * P;
  L.Wc = W - L.Hp;
  L.Mp = L.H + L.Wc * P;
  L.rows = rows;
  L.symmetric_order = (key_off == row_off);
  // exact-order mode (gdn_set_pr_exact_order / GDN_PR_EXACT=1): no slice is ever cut into segments, so EVERY row is summed
  // sequentially in column order by one lane -- bit-identical to src/pr/omp_base.cc:28-30 whatever the row length
  L.exact = lib().pr_exact;
  L.group_ch = L.exact ? 0xffffffffu : (uint32_t)kGroupCh;

  trace("pull_prepare: begin");
  // A gang worker (one process, one thread per GPU) ranks only ITS vertices -- O(m / P) host work per GPU -- and the
  // renumbering of the whole graph is assembled in host memory shared by the workers.  Separate processes (and directed
  // graphs, whose column order is not the row order) rank every partition themselves.
  const bool part_only = gang_worker() && L.symmetric_order && P > 1;
  const int64_t d_lo = part_only ? lo : 0, d_hi = part_only ? hi : m;
  // degree array and row order live in the library's page-locked scratch (reused from call to call)
  RawBuf<int32_t> rdeg_own, kdeg;
  int32_t *rdeg_buf = (int32_t *)host_arena(0, sizeof(int32_t) * (size_t)std::max<int64_t>(d_hi - d_lo, 1));
  if (!rdeg_buf) { rdeg_own.alloc(std::max<int64_t>(d_hi - d_lo, 1)); rdeg_buf = rdeg_own.data(); }
  if (!rdeg_buf) { set_error("out of host memory"); return GDN_ERR_NOMEM; }
  int32_t *rdeg = rdeg_buf - d_lo;                         // indexed by global vertex id in [d_lo, d_hi)
  int64_t bad = 0;          // the device-side validation of these offsets is still in flight: do not index with garbage
#pragma omp parallel for reduction(+ : bad)
  for (int64_t v = d_lo; v < d_hi; v++) {
    rdeg[v] = (int32_t)(row_off[v + 1] - row_off[v]);
    bad += row_off[v + 1] < row_off[v] || (uint64_t)(row_off[v + 1] - row_off[v]) > 0x7fffffffull;
  }
  const int32_t *kd = rdeg;
  if (!L.symmetric_order) {
    kdeg.alloc(m);
#pragma omp parallel for reduction(+ : bad)
    for (int64_t v = 0; v < m; v++) {
      kdeg[v] = (int32_t)(key_off[v + 1] - key_off[v]);
      bad += key_off[v + 1] < key_off[v] || (uint64_t)(key_off[v + 1] - key_off[v]) > 0x7fffffffull;
    }
    kd = kdeg.data();
  }
  if (part_only) {
    // (a malformed partition must not leave the other workers waiting at the barriers below: everybody goes on, the
    // device-side validation reports the error)
    if (bad) {
#pragma omp parallel for
      for (int64_t v = d_lo; v < d_hi; v++) rdeg[v] = 0;
    }
  } else if (bad) { L.prepared = false; return GDN_OK; }      // upload_csr_end reports the malformed CSR
  // on one GPU of a symmetric graph the row order IS the column order: one sort, and newid / sdeg are
  // derived from it on the device (no 268 MB host scatter, no upload)
  const bool one_sort = L.symmetric_order && P == 1;
  RawBuf<int32_t> perm_own, newid, tmp, sdeg, rowid;
  int32_t *newid_host = nullptr;
  int32_t *perm = (int32_t *)host_arena(1, sizeof(int32_t) * (size_t)std::max<int64_t>(rows, 1));
  if (!perm) { perm_own.alloc(std::max<int64_t>(rows, 1)); perm = perm_own.data(); }
  if (!perm) { set_error("out of host memory"); return GDN_ERR_NOMEM; }
  if (one_sort) {
    sort_by_degree(kd, 0, m, perm);
  } else if (part_only) {
    int32_t *shared = (int32_t *)gang_shared(sizeof(int32_t) * (size_t)m);      // collective
    if (!shared) { set_error("out of page-locked host memory"); return GDN_ERR_NOMEM; }
    sdeg.alloc(std::max<int64_t>(rows, 1));
    if (rows > 0) sort_by_degree(rdeg, lo, hi, perm, sdeg.data());
#pragma omp parallel for
    for (int64_t j = 0; j < rows; j++)
      shared[lo + perm[j]] = (int32_t)(j < L.Hp ? (int64_t)R * L.Hp + j : L.H + (int64_t)R * L.Wc + (j - L.Hp));
    gang_barrier();                                        // every worker's slice of the renumbering is in
    newid_host = shared;
  } else {
    newid.alloc(m); tmp.alloc(W); sdeg.alloc(std::max<int64_t>(rows, 1));
    for (int q = 0; q < P; q++) {
      const int64_t qlo = std::min<int64_t>((int64_t)q * W, m), qhi = std::min<int64_t>(qlo + W, m);
      if (qhi <= qlo) continue;
      sort_by_degree(kd, qlo, qhi, tmp.data());
      const int64_t n = qhi - qlo;
#pragma omp parallel for
      for (int64_t j = 0; j < n; j++)
        newid[qlo + tmp[j]] = (int32_t)(j < L.Hp ? (int64_t)q * L.Hp + j : L.H + (int64_t)q * L.Wc + (j - L.Hp));
    }
    if (rows > 0) sort_by_degree(rdeg, lo, hi, perm, sdeg.data());
    if (!L.symmetric_order) {
      rowid.alloc(std::max<int64_t>(rows, 1));
#pragma omp parallel for
      for (int64_t j = 0; j < rows; j++) rowid[j] = newid[lo + perm[j]];
    }
    newid_host = newid.data();
  }
  trace("pull_prepare: orders");
  auto len_of = [&](int64_t j) -> int32_t { return rdeg[lo + perm[j]]; };     // length of sorted row j
  int64_t n_nz = 0;
  {
    // rows are sorted by length, descending: the non-empty ones are a prefix
    int64_t a = 0, b = rows;
    while (a < b) { const int64_t mid = (a + b) >> 1; if (len_of(mid) > 0) a = mid + 1; else b = mid; }
    n_nz = a;
  }
  L.n_nz_rows = n_nz;
  L.n_slices = (int32_t)((n_nz + 31) / 32);
  // slice pointers in int4 groups: slice s = 32 lanes x ceil(width/4) groups, width = longest (= first) row
  std::vector<uint32_t> sptr((size_t)L.n_slices + 1);
  std::vector<int32_t> width((size_t)L.n_slices + 1);
#pragma omp parallel for
  for (int32_t s = 0; s < L.n_slices; s++) width[s] = len_of((int64_t)s * 32);
  uint64_t tot = 0;
  for (int32_t s = 0; s < L.n_slices; s++) {
    sptr[s] = (uint32_t)tot;
    tot += 32ull * ((width[s] + 3) / 4);
    if (tot >= 0xffff0000ull) { L.prepared = false; return GDN_OK; }   // 32-bit group offsets would overflow: plain CSR path (gather.cu pr_t)
  }
  sptr[L.n_slices] = (uint32_t)tot;
  L.n_groups = tot;
  L.h_slice_ptr = sptr;
  // work items: chunk k owns the light slices that START in [k*CH,(k+1)*CH); wide slices are cut into segments
  // (the chunk grid stays kGroupCh in exact-order mode; only the cut of wide slices into segments is disabled)
  L.n_chunks = (int32_t)(tot / kGroupCh + 1);
  std::vector<int32_t> chunk((size_t)L.n_chunks + 1);
  {
    int32_t s = 0;
    for (int32_t k = 0; k < L.n_chunks; k++) {
      while (s < L.n_slices && sptr[s] < (uint64_t)k * kGroupCh) s++;
      chunk[k] = s;
    }
    chunk[L.n_chunks] = L.n_slices;
  }
  // Exact slices: the widest rows are never cut into segments and never banded -- every lane of ONE warp adds ITS row
  // sequentially in column order like src/pr/omp_base.cc:28-30.  On a long row the fp32 rounding of a re-ordered sum differs
  // from the reference's by ~ sqrt(length) ulps; at Kronecker scale 26 (rows of 10^6 entries that carry percents of the
  // score mass) that alone is 1.15e-6 of L1 distance -- over the 1e-6 parity bar (profiles/r2_pr_exact_threshold.txt:
  // 0.58e-6 with the one slice wider than 2^18 columns exact, 0.65e-6 with the 11 wider than 2^16, 0.17e-6 with 559).  A dependent chain of a million adds is 3.4 ms whatever
  // feeds it, so these rows are summed by the block-parallel emulation of ordered_sum.cuh (same bits, no chain); the
  // exact-order mode keeps a true chain (values gathered by the whole grid, streamed through a TMA-fed ring).
  const char *e_xc = getenv("GDN_PR_EXACT_COLS");
  const int64_t exact_cols = e_xc ? atoll(e_xc) : (L.exact ? kExactColsStrict : kExactCols);
  // ... as long as they stay cheap: their ids are mostly cold (a hub row points everywhere), one HBM sector per gather,
  // whereas the banded layout serves most of them from shared memory -- so the default mode spends at most 3 % of the
  // non-zeros (or 48 M, whichever is more; slots of the slices, padding included) on exact slices, widest first
  // (profiles/r2_pr_exact_threshold.txt).  GDN_PR_EXACT_BUDGET (slots) overrides.
  const char *e_xb = getenv("GDN_PR_EXACT_BUDGET");
  const uint64_t budget = L.exact ? ~0ull : (e_xb ? strtoull(e_xb, nullptr, 10) : std::max<uint64_t>((uint64_t)(row_off[hi] - row_off[lo]) * 3 / 100, 48ull << 20));
  L.n_exact = 0;
  while (L.n_exact < L.n_slices && width[L.n_exact] > exact_cols && (sptr[L.n_exact + 1] - sptr[L.n_exact]) > (uint32_t)kGroupCh &&
         (uint64_t)sptr[L.n_exact + 1] * 4 <= budget)
    L.n_exact++;
  std::vector<int32_t> hslice, hfirst;
  std::vector<int2> hseg;
  for (int32_t s = L.n_exact; s < L.n_slices; s++) {
    const uint32_t sz = sptr[s + 1] - sptr[s];
    if (sz <= L.group_ch) break;                         // widths are non-increasing
    hslice.push_back(s);
    hfirst.push_back((int32_t)hseg.size());
    for (uint32_t q = 0; q < (sz + kGroupCh - 1) / kGroupCh; q++) hseg.push_back(make_int2(s, (int)q));
  }
  hfirst.push_back((int32_t)hseg.size());
  L.n_heavy_slices = (int32_t)hslice.size();
  L.n_heavy_segs = (int32_t)hseg.size();
  // wide slices for the fill kernel: the whole grid strides their column tiles
  L.n_fill_wide = 0;
  for (int32_t s = 0; s < L.n_slices && width[s] > 2048; s++) L.n_fill_wide++;

  trace("pull_prepare: slices");
  cudaStream_t st = lib().stream;
  GDN_CHECK(upload(g, &L.perm, perm, (size_t)rows));
  if (one_sort) {
    // the device offsets were queued on this stream by upload_csr_begin, so they are ready for sdeg_from_perm
    GDN_CUDA(cudaMalloc((void **)&L.newid, sizeof(int32_t) * std::max<int64_t>(m, 4)));
    GDN_CUDA(cudaMalloc((void **)&L.sdeg, sizeof(int32_t) * std::max<int64_t>(rows, 4)));
    g->device_bytes += sizeof(int32_t) * (m + rows);
    const DevCsr &c = g->symmetric ? g->out : g->in;
    const int grid = lib().sm_count * 8;
    newid_from_perm<<<grid, 256, 0, st>>>(L.perm, L.newid, rows);
    if (c.off64) sdeg_from_perm<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)c.rowptr, L.perm, L.sdeg, rows);
    else sdeg_from_perm<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)c.rowptr, L.perm, L.sdeg, rows);
  } else {
    GDN_CHECK(upload(g, &L.newid, newid_host, (size_t)m));
    GDN_CHECK(upload(g, &L.sdeg, sdeg.data(), (size_t)rows));
    if (!L.symmetric_order) GDN_CHECK(upload(g, &L.rowid, rowid.data(), (size_t)rows));
  }
  GDN_CHECK(upload(g, &L.slice_ptr, sptr.data(), sptr.size()));
  GDN_CHECK(upload(g, &L.chunk_slice, chunk.data(), chunk.size()));
  if (L.n_heavy_slices) {
    GDN_CHECK(upload(g, &L.heavy_slice, hslice.data(), hslice.size()));
    GDN_CHECK(upload(g, &L.heavy_first, hfirst.data(), hfirst.size()));
    GDN_CHECK(upload(g, &L.heavy_seg, hseg.data(), hseg.size()));
    GDN_CUDA(cudaMalloc((void **)&L.partial, sizeof(float) * 32 * (size_t)L.n_heavy_segs));
    g->device_bytes += sizeof(float) * 32 * (size_t)L.n_heavy_segs;
  }
  // one-shot PageRank: build the SELL array behind the pieces of the column upload still in flight
  const bool stream_build = one_sort && !g->col_ev.empty() && L.n_slices > 0;
  cudaEvent_t staged = nullptr;
  if (stream_build) {
    GDN_CUDA(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
    GDN_CUDA(cudaEventRecord(staged, st));  // the uploads above have left the host buffers once this fires
    GDN_CHECK(col_upload_rest(g));          // the pieces upload_csr_begin held back so that the uploads above were not queued behind them
    GDN_CHECK(pull_stream_fill(g));
    GDN_CUDA(cudaEventSynchronize(staged));
    GDN_CUDA(cudaEventDestroy(staged));
  } else {
    GDN_CUDA(cudaStreamSynchronize(st));    // host buffers die here
  }
  GDN_CUDA(cudaGetLastError());
  trace("pull_prepare: uploaded");
  L.prepared = true;
  return GDN_OK;
}
template int pull_prepare<uint64_t>(gdn_graph *, const uint64_t *, const uint64_t *);
template int pull_prepare<int32_t>(gdn_graph *, const int32_t *, const int32_t *);

// Share of the SELL array held by the rows of the 64 hottest bands' worth of ids (rows and columns are ranked by the same
// degrees on a symmetric graph): 0.87 at Kronecker scale 26, 0.05 at urand-26.  "Is there a hot set?"
double pull_hot_share(const PullLayout &L) {
  if (L.h_slice_ptr.empty() || L.n_slices < 1) return 0.0;
  const int64_t hot_slices = std::min<int64_t>(L.n_slices, (int64_t)64 * kHotMax / 32 / std::max(L.P, 1));
  return (double)L.h_slice_ptr[hot_slices] / (double)std::max<uint32_t>(L.h_slice_ptr[L.n_slices], 1u);
}

// ------------------------------------------------------------------ device: fill the SELL array
// One warp per slice; 32x32 tiles are read row-wise (coalesced along a CSR row),
// renumbered through newid[], transposed in shared memory and written lane = row
// (coalesced 512-byte stores).  Padding is -1.
template <typename OffT>
__device__ __forceinline__ void fill_tile(const int32_t *__restrict__ col, const int32_t *__restrict__ newid,
                                          int (*tile)[33], OffT b_l, uint32_t d_l, uint32_t k0, uint32_t ngk,
                                          int4 *__restrict__ dst, int lane) {
  for (int rr = 0; rr < 32; rr++) {
    const OffT b = __shfl_sync(kFull, b_l, rr);
    const uint32_t d = __shfl_sync(kFull, d_l, rr);
    int c = -1;
    if (k0 + lane < d) c = newid[col[b + k0 + lane]];
    tile[rr][lane] = c;
  }
  __syncwarp();
#pragma unroll
  for (int gq = 0; gq < 8; gq++) {
    const uint32_t kk = k0 / 4 + gq;
    if (kk < ngk) dst[(size_t)kk * 32 + lane] = make_int4(tile[lane][4 * gq], tile[lane][4 * gq + 1], tile[lane][4 * gq + 2], tile[lane][4 * gq + 3]);
  }
  __syncwarp();
}

template <typename OffT>
__global__ void __launch_bounds__(256, 4)
sell_fill(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ perm,
          const int32_t *__restrict__ newid, const uint32_t *__restrict__ slice_ptr, int32_t n_slices, int64_t n_nz_rows,
          int32_t n_wide, int4 *__restrict__ sell) {
  __shared__ int tiles[8][32][33];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
  // wide slices: all warps of the grid share the column tiles of one slice at a time
  for (int32_t s = 0; s < n_wide; s++) {
    const int64_t j = (int64_t)s * 32 + lane;
    OffT b_l = 0; uint32_t d_l = 0;
    if (j < n_nz_rows) { const int32_t old = perm[j]; b_l = rowptr[old]; d_l = (uint32_t)(rowptr[old + 1] - b_l); }
    const uint32_t w = __shfl_sync(kFull, d_l, 0), ngk = (w + 3) / 4;
    for (uint64_t k0 = (uint64_t)warp * 32; k0 < w; k0 += (uint64_t)nwarps * 32)
      fill_tile<OffT>(col, newid, tiles[wib], b_l, d_l, (uint32_t)k0, ngk, sell + slice_ptr[s], lane);
  }
  for (int64_t s = n_wide + warp; s < n_slices; s += nwarps) {
    const int64_t j = s * 32 + lane;
    OffT b_l = 0; uint32_t d_l = 0;
    if (j < n_nz_rows) { const int32_t old = perm[j]; b_l = rowptr[old]; d_l = (uint32_t)(rowptr[old + 1] - b_l); }
    const uint32_t w = __shfl_sync(kFull, d_l, 0), ngk = (w + 3) / 4;
    for (uint32_t k0 = 0; k0 < w; k0 += 32)
      fill_tile<OffT>(col, newid, tiles[wib], b_l, d_l, k0, ngk, sell + slice_ptr[s], lane);
  }
}


// ------------------------------------------------------------------ device: streaming build of the SELL array
// The one-shot entry point (oneshot.cu) pays for the layout inside the call, and sell_fill above needs the whole
// column array resident before its first slice (44 ms at Kron-26, all of it waiting on 2.1 G random newid[]
// lookups).  sell_scatter builds the same array from the CSR side instead, for any range [e0, e1) of non-zeros, so it
// runs piece by piece BEHIND the PCIe copy of the column array (graph.cu upload_csr_begin, 256 MB pieces) and is hidden
// by it.  A CTA owns 1024 consecutive non-zeros: two binary searches bound the rows they belong to, the offsets of
// those rows are staged in shared memory, every thread locates the row of its four entries there and writes
//     sell[slice_ptr[j / 32] * 4 + (k / 4) * 128 + (j % 32) * 4 + k % 4] = newid[col[e]]      (j = newid[row], k = e - rowptr[row])
// into the array pre-filled with -1 (the padding).  Column ids are range-checked here (validate_csr skips them in
// this mode): a bad id raises the verdict word and is not followed.
constexpr int kScatRows = 2048;        // offsets staged per CTA; longer row runs (mostly empty rows) search in global memory

template <typename OffT>
__global__ void __launch_bounds__(256, 4)
sell_scatter(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const int32_t *__restrict__ newid,
             const uint32_t *__restrict__ slice_ptr, int32_t *__restrict__ sell, int64_t rows, uint64_t e0, uint64_t e1,
             int64_t m, int *flag) {
  __shared__ int32_t s_off[kScatRows + 1];
  __shared__ int64_t s_row[2];
  const uint64_t n_tiles = (e1 - e0 + 1023) / 1024;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t tb = e0 + tile * 1024, te = tb + 1024 < e1 ? tb + 1024 : e1;
    if (threadIdx.x == 0 || threadIdx.x == 32) {
      // last row whose first non-zero is at or before `key` (upper_bound - 1); rowptr[0] = 0 <= key
      const uint64_t key = threadIdx.x == 0 ? tb : te - 1;
      int64_t lo = 0, hi = rows;
      while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if ((uint64_t)rowptr[mid] <= key) lo = mid; else hi = mid - 1;
      }
      s_row[threadIdx.x >> 5] = lo;
    }
    __syncthreads();
    const int64_t r_lo = s_row[0], r_hi = s_row[1];
    const int64_t nr = r_hi - r_lo + 1;
    const bool staged = nr <= kScatRows;
    if (staged)
      for (int64_t i = threadIdx.x; i <= nr; i += 256) s_off[i] = (int32_t)((int64_t)rowptr[r_lo + i] - (int64_t)tb);
    __syncthreads();
    const uint64_t e = tb + 4ull * threadIdx.x;
    if (e < te) {
      int c4[4];
      if (e + 4 <= te) { const int4 q = *reinterpret_cast<const int4 *>(col + e); c4[0] = q.x; c4[1] = q.y; c4[2] = q.z; c4[3] = q.w; }
      else { for (int u = 0; u < 4; u++) c4[u] = e + u < te ? col[e + u] : 0; }
      // row of the first entry
      int64_t i;
      if (staged) {
        const int32_t key = (int32_t)(e - tb);
        int32_t lo = 0, hi = (int32_t)nr - 1;
        while (lo < hi) { const int32_t mid = (lo + hi + 1) >> 1; if (s_off[mid] <= key) lo = mid; else hi = mid - 1; }
        i = lo;
      } else {
        int64_t lo = r_lo, hi = r_hi;
        while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if ((uint64_t)rowptr[mid] <= e) lo = mid; else hi = mid - 1; }
        i = lo - r_lo;
      }
      int64_t row_end = staged ? (int64_t)s_off[i + 1] + (int64_t)tb : (int64_t)rowptr[r_lo + i + 1];
      int64_t row_beg = staged ? (int64_t)s_off[i] + (int64_t)tb : (int64_t)rowptr[r_lo + i];
      int32_t j = newid[r_lo + i];
      size_t base = (size_t)slice_ptr[j >> 5] * 4 + (size_t)(j & 31) * 4;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int64_t ee = (int64_t)e + u;
        if ((uint64_t)ee >= te) break;
        while (ee >= row_end) {                    // next non-empty row
          i++;
          row_beg = row_end;
          row_end = staged ? (int64_t)s_off[i + 1] + (int64_t)tb : (int64_t)rowptr[r_lo + i + 1];
          if (ee < row_end) { j = newid[r_lo + i]; base = (size_t)slice_ptr[j >> 5] * 4 + (size_t)(j & 31) * 4; }
        }
        const uint32_t k = (uint32_t)(ee - row_beg);
        const int c = c4[u];
        int v = -1;
        if ((unsigned)c < (unsigned)m) v = newid[c]; else atomicExch(flag, 3);
        sell[base + (size_t)(k >> 2) * 128 + (k & 3)] = v;
      }
    }
    __syncthreads();
  }
}

// Queue the streaming build behind the chunked column upload (gdn_graph::col_ev).  Everything it reads besides the
// columns (offsets, newid, slice_ptr) was queued on the library stream before.
static int pull_stream_fill(gdn_graph *g) {
  PullLayout &L = g->pull;
  const DevCsr &c = g->has_in && g->in.col ? g->in : g->out;
  cudaStream_t st = lib().stream;
  GDN_CUDA(cudaMalloc((void **)&L.sell, sizeof(int4) * std::max<uint64_t>(L.n_groups, 1) + 256));
  g->device_bytes += sizeof(int4) * L.n_groups;
  GDN_CUDA(cudaMemsetAsync(L.sell, 0xff, sizeof(int4) * std::max<uint64_t>(L.n_groups, 1) + 256, st));
  uint64_t e0 = 0;
  for (size_t k = 0; k < g->col_ev.size(); k++) {
    const uint64_t e1 = g->col_end[k];
    GDN_CUDA(cudaStreamWaitEvent(st, g->col_ev[k], 0));
    const int grid = (int)std::min<uint64_t>((e1 - e0 + 1023) / 1024, (uint64_t)lib().sm_count * 8);
    if (c.off64)
      sell_scatter<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)c.rowptr, c.col, L.newid, L.slice_ptr, (int32_t *)L.sell,
                                                   c.rows, e0, e1, g->m, g->col_flag);
    else
      sell_scatter<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)c.rowptr, c.col, L.newid, L.slice_ptr, (int32_t *)L.sell,
                                                   c.rows, e0, e1, g->m, g->col_flag);
    e0 = e1;
  }
  GDN_CUDA(cudaGetLastError());
  return GDN_OK;
}

int pull_build_sell(gdn_graph *g) {
  PullLayout &L = g->pull;
  if (L.sell || !L.prepared) return GDN_OK;
  const auto t0 = std::chrono::steady_clock::now();
  const DevCsr &c = g->in;
  GDN_CUDA(cudaMalloc((void **)&L.sell, sizeof(int4) * std::max<uint64_t>(L.n_groups, 1) + 256));
  g->device_bytes += sizeof(int4) * L.n_groups;
  if (L.n_slices > 0) {
    const int grid = lib().sm_count * 8;
    if (c.off64)
      sell_fill<uint64_t><<<grid, 256, 0, lib().stream>>>((const uint64_t *)c.rowptr, c.col, L.perm, L.newid, L.slice_ptr,
                                                          L.n_slices, L.n_nz_rows, L.n_fill_wide, L.sell);
    else
      sell_fill<uint32_t><<<grid, 256, 0, lib().stream>>>((const uint32_t *)c.rowptr, c.col, L.perm, L.newid, L.slice_ptr,
                                                          L.n_slices, L.n_nz_rows, L.n_fill_wide, L.sell);
  }
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  GDN_CUDA(cudaGetLastError());
  trace("pull_build_sell: done");
  g->prep_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return GDN_OK;
}

// ------------------------------------------------------------------ device: the PageRank iteration
// One gathered contrib value.  Three tiers (profiles/r1_gather_microbench_b200.txt): ids < H come from the
// shared-memory table; ids < warm are the part of contrib that fits the 126 MB L2 and are loaded with an
// L2 evict-last hint; colder ids (a few % of the edges of a Kronecker graph) are loaded evict-first so that
// their one-touch sectors do not push the warm part out of L2.  Written as three PREDICATED loads (no
// branches): the compiler's branchy version spent a fifth of its issue slots on reconvergence and left
// the 16 loads of a trip interleaved with them (ncu r1: stall_mio 18 %, short scoreboard 30 %).
__device__ __forceinline__ float pull_one(const SellArgs &a, int32_t hot_n, uint32_t s_hot_addr, int c, uint64_t pol_first, uint64_t pol_last) {
  float v = 0.f;
  const float *p = a.contrib_in + c;
  const int32_t t = tier_id(a, c);
  asm volatile(
      "{\n\t.reg .pred ph, pw, pc;\n\t"
      "setp.lt.u32 ph, %1, %2;\n\t"               // hot: 0 <= c < H   (c = -1 is 0xffffffff: never hot)
      "setp.ge.s32 pw, %1, %2;\n\t"               // not hot and not padding
      "setp.ge.s32 pc, %8, %3;\n\t"               // cold (by its position inside its rank's slice); never true for c = -1
      "and.pred pw, pw, !pc;\n\t"
      "@ph ld.shared.f32 %0, [%4];\n\t"
      "@pw ld.global.nc.L2::cache_hint.f32 %0, [%5], %6;\n\t"
      "@pc ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%5], %7;\n\t}"
      : "+f"(v)
      : "r"(c), "r"(hot_n), "r"(a.warm), "r"(s_hot_addr + 4u * (uint32_t)c), "l"(p), "l"(pol_last), "l"(pol_first), "r"(t));
  return v;
}

// ------------------------------------------------------------------ the pipelined iteration kernel
// A kernel that keeps ONE trip (16 gathers per lane) in flight and then waits for it was measured
// (profiles/r1_pr_tier_probe.txt) at the SUM of a 3.3 ms "stream + shared-table" floor and the L2-tier
// gathers at exactly the 1 sector/clk/SM miss-path rate -- the two never overlap, because all warps of an SM queue
// their gathers behind each other, drain together, and then wait together for the next index groups (a convoy; L1TEX
// 53 % busy).  Here a warp's work is flattened into one sequence of trips that crosses slice and item boundaries, and
// the loop is software-pipelined two trips deep:
//      request index groups of trip t+2  |  issue the gathers of trip t+1  |  add the values of trip t
// so the L1TEX miss path always holds gathers of this warp while it waits, adds and runs epilogues.  The row epilogue's
// inputs (old score, degree) are requested with the last trip of their slice.  The per-row addition order is column
// order, sequential fp32 (src/pr/omp_base.cc:28-30).
struct TripDesc {
  uint32_t g;      // first int4 unit of the trip (lane 0's)
  int32_t n;       // index groups per lane (0: this warp has no work left)
  int32_t fin;     // 1: last trip of a light slice (ref = slice), 2: last trip of a wide-slice segment (ref = item), 0: neither
  int32_t ref;
};

// Work distribution is dynamic: a warp draws batches of kBatch items from one device counter, so a warp that first walks
// the long sequential chain of an exact slice simply draws fewer batches and nobody waits for it at the end of the launch.
constexpr int kBatch = 8;
template <int G>
struct TripIter {
  const SellArgs &a;
  int32_t item, item_end;         // batch in progress; work items of one GPU fit 31 bits (checked by pr_run_sell)
  int32_t s, s_end, sa;
  uint32_t g, g_end, lo, hi;      // lo/hi: slice_ptr[sa + lane], slice_ptr[sa + lane + 1] of the current chunk
  int32_t kind, ref;
  int lane;
  __device__ __forceinline__ TripIter(const SellArgs &a_, int lane_)
      : a(a_), item(0), item_end(0), s(0), s_end(0), sa(0), g(0), g_end(0), lo(0), hi(0), kind(0), ref(0), lane(lane_) {}
  __device__ __forceinline__ void load_bounds() {
    lo = a.slice_ptr[min(sa + lane, s_end)];
    hi = a.slice_ptr[min(sa + lane + 1, s_end)];
  }
  __device__ __forceinline__ bool grab() {
    const int32_t n_items = a.n_heavy_segs + a.n_chunks;
    int32_t b = 0;
    if (lane == 0) b = atomicAdd(a.work_counter, 1);
    b = __shfl_sync(kFull, b, 0);
    if (b > 0x7fffffff / kBatch) return false;
    item = b * kBatch;
    item_end = min(item + kBatch, n_items);
    return item < n_items;
  }
  __device__ __forceinline__ TripDesc next() {
    for (;;) {
      if (g < g_end) {
        TripDesc d;
        const uint32_t left = (g_end - g) >> 5;
        // a slice (or segment) is one contiguous piece of the index array: every 8 trips ask L2 for the 4 KB that lie
        // pf_trips ahead, so that the index loads of the pipeline wait for L2 and not for HBM
        if (a.pf_trips && lane == 0 && (((g_end - g) >> 5) & 7u) == 0 && left > (uint32_t)a.pf_trips)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.sell + g + ((size_t)a.pf_trips << 5)),
                       "r"(min(8u, left - (uint32_t)a.pf_trips) * 512u) : "memory");
        d.g = g;
        d.n = left < (uint32_t)G ? (int32_t)left : G;
        g += (uint32_t)d.n << 5;
        d.fin = g >= g_end ? kind : 0;
        d.ref = ref;
        return d;
      }
      if (s < s_end) {
        if (s - sa >= 32) { sa = s; load_bounds(); }
        const uint32_t g0 = __shfl_sync(kFull, lo, s - sa), g1 = __shfl_sync(kFull, hi, s - sa);
        ref = s;
        s++;
        if (ref < a.n_exact || g1 - g0 > a.group_ch) continue;      // exact slice: the chain warps' work; wide slice: handled as segments
        g = g0; g_end = g1; kind = 1;
        continue;
      }
      if (item >= item_end && (item_end < 0 || !grab())) { item_end = -1; TripDesc d; d.g = 0; d.n = 0; d.fin = 0; d.ref = 0; return d; }
      const int32_t it = item++;
      if (it < a.n_heavy_segs) {
        const int2 hs = a.heavy_seg[it];
        const uint32_t s0 = a.slice_ptr[hs.x], s1 = a.slice_ptr[hs.x + 1];
        g = s0 + (uint32_t)hs.y * a.group_ch;
        g_end = (s1 - g > a.group_ch) ? g + a.group_ch : s1;
        kind = 2; ref = it; s = s_end = 0;
      } else {
        const int32_t k = it - a.n_heavy_segs;
        sa = a.chunk_slice[k]; s = sa; s_end = a.chunk_slice[k + 1];
        load_bounds();
        g = g_end = 0;
      }
    }
  }
};

// ------------------------------------------------------------------ exact slices, exact-order mode: gather pass + TMA-fed sequential chain
// (the default mode sums them by the block-parallel emulation of ordered_sum.cuh)
// Values of the exact slices in the layout of their index groups: vals[g] = contrib[sell[g]] (padding -> +0.0f, which
// leaves an fp32 sum unchanged).  Whole grid, one int4 group per thread and step.
__global__ void __launch_bounds__(256, 4)
pr_exact_gather_sell(SellArgs a, float4 *__restrict__ vals, uint32_t n_groups) {
  if (*a.done) return;
  const uint64_t pol = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  for (uint32_t g = blockIdx.x * 256 + threadIdx.x; g < n_groups; g += gridDim.x * 256) {
    const int4 c = ld_stream_v4(a.sell + g, pol);
    const int id[4] = {c.x, c.y, c.z, c.w};
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      v[u] = 0.f;
      if (id[u] >= 0) {
        const float *p = a.contrib_in + id[u];
        if (tier_id(a, id[u]) < a.warm) v[u] = ld_gather_f32(p, pol_last);
        else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[u]) : "l"(p), "l"(pol));
      }
    }
    vals[g] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D; SASS UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// The chain eats 128 bytes every ~4.5 cycles (the dependent fp32 add).  Its value stream comes through a ring of TMA bulk
// copies fed by a PRODUCER warp (warp 1 of the CTA: full / empty mbarriers per buffer, L2 prefetches further ahead), so
// that the chain warp itself only waits, loads from shared memory and adds.  Measured at Kronecker scale 26 (1.0 M
// columns in the top slice): 17.8 cycles per column with the 31 other warps of the CTA gathering beside it (their sectors
// queue ahead of the bulk copies on the SM's one request port to L2), 8.4 with the chain warp issuing its own copies and
// the CTA otherwise idle -- hence a dedicated producer and a CTA that does nothing else until its chain is done.
constexpr int kChainHot = 16384;          // hot-table entries of a CTA that hosts a chain warp (64 KB)
constexpr int kRingBufs = 3;              // chunks of the value stream in shared memory ...
constexpr int kRingGroups = 64;           // ... of 64 index groups = 256 columns x 32 rows x 4 B = 32 KB each
constexpr int kRingAhead = 8;             // chunks requested into L2 ahead of the ring
constexpr size_t kChainSmem = (size_t)kChainHot * sizeof(float) + 256 + (size_t)kRingBufs * kRingGroups * 32 * sizeof(float4);
// (both chain layouts fit the 192 KB of the full hot table: a launch that asks for more shared memory loses L1 on EVERY SM,
// and the L1 hits of the ordinary gathers are worth 1.7 ms per iteration at Kronecker scale 26 -- measured)
static_assert(2 * kRingBufs * sizeof(uint64_t) <= 256 && kChainSmem <= (size_t)kHotMax * sizeof(float), "chain CTA: table + barriers + ring within the full table's 192 KB");
constexpr uint32_t kChunkUnits = kRingGroups * 32;          // float4 units per chunk

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Producer warp: the chunks of exact slices first, first + stride, ... in order, each into the next ring buffer once the
// chain warp has released it.
__device__ __forceinline__ void exact_chain_producer(const SellArgs &a, float4 *ring, uint64_t *full, uint64_t *empty, int lane, int first, int stride) {
  uint32_t k = 0;                                             // running chunk index: buffer k % kRingBufs, phase k / kRingBufs
  for (int e = first; e < a.n_exact; e += stride) {
    const uint32_t g0 = a.slice_ptr[e], g1 = a.slice_ptr[e + 1];
    const uint32_t ngl = (g1 - g0) >> 5;                      // groups per lane
    const uint32_t n_chunks = (ngl + kRingGroups - 1) / kRingGroups;
    const float4 *src = a.exact_vals + g0;
    if (lane == 0)
      for (uint32_t c = 0; c < (uint32_t)kRingAhead && c < n_chunks; c++)
        tma_prefetch_l2(src + (size_t)c * kChunkUnits, min((uint32_t)kRingGroups, ngl - c * kRingGroups) * 32 * (uint32_t)sizeof(float4));
    for (uint32_t c = 0; c < n_chunks; c++, k++) {
      const uint32_t slot = k % kRingBufs;
      mbar_wait(&empty[slot], ((k / kRingBufs) & 1u) ^ 1u);   // (passes at once the first time round)
      if (lane == 0) {
        const uint32_t bytes = min((uint32_t)kRingGroups, ngl - c * kRingGroups) * 32 * (uint32_t)sizeof(float4);
        // no proxy fence: the buffer was only READ through the generic proxy and those loads had delivered before the release
        mbar_expect_tx(&full[slot], bytes);
        tma_load_1d(ring + (size_t)slot * kChunkUnits, src + (size_t)c * kChunkUnits, bytes, &full[slot]);
        const uint32_t pc = c + kRingAhead;
        if (pc < n_chunks) tma_prefetch_l2(src + (size_t)pc * kChunkUnits, min((uint32_t)kRingGroups, ngl - pc * kRingGroups) * 32 * (uint32_t)sizeof(float4));
      }
      __syncwarp();
    }
  }
}

// Chain warp (lane = row): adds the rows of its slices sequentially in column order, src/pr/omp_base.cc:28-30.
__device__ __forceinline__ void exact_chain(const SellArgs &a, const float4 *ring, uint64_t *full, uint64_t *empty, int lane, int first, int stride, double &err) {
  uint32_t k = 0;
  for (int e = first; e < a.n_exact; e += stride) {
    const uint32_t g0 = a.slice_ptr[e], g1 = a.slice_ptr[e + 1];
    const uint32_t ngl = (g1 - g0) >> 5;
    const uint32_t n_chunks = (ngl + kRingGroups - 1) / kRingGroups;
    float acc = 0.f;
    for (uint32_t c = 0; c < n_chunks; c++, k++) {
      const uint32_t slot = k % kRingBufs;
      mbar_wait(&full[slot], (k / kRingBufs) & 1u);
      const uint32_t n = min((uint32_t)kRingGroups, ngl - c * kRingGroups);
      const float4 *t = ring + (size_t)slot * kChunkUnits + lane;
      if (n == (uint32_t)kRingGroups) {
        // four groups (16 columns) are read one step ahead of the 16 dependent adds that consume them
        float4 a0 = t[0], a1 = t[32], a2 = t[64], a3 = t[96];
#pragma unroll 4
        for (int q = 4; q <= kRingGroups; q += 4) {
          float4 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
          if (q < kRingGroups) { b0 = t[q * 32]; b1 = t[(q + 1) * 32]; b2 = t[(q + 2) * 32]; b3 = t[(q + 3) * 32]; }
          acc = __fadd_rn(acc, a0.x); acc = __fadd_rn(acc, a0.y); acc = __fadd_rn(acc, a0.z); acc = __fadd_rn(acc, a0.w);
          acc = __fadd_rn(acc, a1.x); acc = __fadd_rn(acc, a1.y); acc = __fadd_rn(acc, a1.z); acc = __fadd_rn(acc, a1.w);
          acc = __fadd_rn(acc, a2.x); acc = __fadd_rn(acc, a2.y); acc = __fadd_rn(acc, a2.z); acc = __fadd_rn(acc, a2.w);
          acc = __fadd_rn(acc, a3.x); acc = __fadd_rn(acc, a3.y); acc = __fadd_rn(acc, a3.z); acc = __fadd_rn(acc, a3.w);
          a0 = b0; a1 = b1; a2 = b2; a3 = b3;
        }
      } else {
        for (uint32_t q = 0; q < n; q++) {
          const float4 v = t[q * 32];
          acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    const int64_t j = (int64_t)e * 32 + lane;
    if (j < a.n_nz_rows) pr_epilogue(a, j, acc, err);
  }
}

// G index groups (4 G gathers) per lane and trip, D trips of gathers in flight per warp, THREADS / 32 warps per SM.
template <int G, int D, int THREADS>
__device__ __forceinline__ void pr_sell_pipe_body(const SellArgs &a) {
  extern __shared__ float s_hot[];
  if (*a.done) return;
  // a CTA that hosts a chain warp trades most of its hot table for the chain's ring of value chunks
  // exact-order mode: CTA c first walks exact slices c, c + grid, ... as a true sequential chain (a.strict_chain)
  const bool chain_cta = a.strict_chain && (int)blockIdx.x < a.n_exact;
  const int32_t hot_n = chain_cta ? min(a.hot_n, kChainHot) : a.hot_n;
  for (int i = threadIdx.x; i < hot_n; i += THREADS) s_hot[i] = a.contrib_in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int32_t warp = (int32_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const uint64_t pol = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  const uint32_t s_hot_addr = (uint32_t)__cvta_generic_to_shared(s_hot);
  const int32_t *degs = a.sout ? a.sout : a.sdeg;
  double err = 0.0;
  float acc = 0.f;
  if (chain_cta) {
    // barriers, then the ring, behind the (shortened) hot table; warp 0 = chain, warp 1 = producer.  The other warps of
    // the CTA wait here too: their gathers would queue ahead of the chain's copies, and the chain is the critical path of
    // the launch -- the dynamic work queue hands their share to the other SMs meanwhile.
    unsigned char *base = reinterpret_cast<unsigned char *>(s_hot) + (size_t)kChainHot * sizeof(float);
    uint64_t *full = reinterpret_cast<uint64_t *>(base), *empty = full + kRingBufs;
    float4 *ring = reinterpret_cast<float4 *>(base + 256);
    if (threadIdx.x == 0) {
      for (int i = 0; i < kRingBufs; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) exact_chain(a, ring, full, empty, lane, (int)blockIdx.x, (int)gridDim.x, err);
    else if (threadIdx.x < 64) exact_chain_producer(a, ring, full, empty, lane, (int)blockIdx.x, (int)gridDim.x);
    __syncthreads();
  }
  TripIter<G> it(a, lane);
  int4 I[2][G];              // index groups: requested one step before their gathers are issued
  float V[D][4 * G];         // gathered values of the D trips in flight
  float S[D];                // old score / degree of the row a trip finishes (requested with its gathers)
  int32_t Dg[D];
  // a trip in flight is remembered as {fin | 4 * (n != 0), ref}
  int32_t dflag[D], dref[D];
  TripDesc dnext;
#pragma unroll
  for (int i = 0; i < D; i++) { S[i] = 0.f; Dg[i] = 1; dflag[i] = 0; dref[i] = 0; }

  auto load_idx = [&](int4 (&X)[G], const TripDesc &d) {
    const int4 *p = a.sell + d.g + lane;
#pragma unroll
    for (int u = 0; u < G; u++) X[u] = u < d.n ? ld_stream_v4(p + 32 * u, pol) : make_int4(-1, -1, -1, -1);
  };
  auto gather = [&](float (&W)[4 * G], const int4 (&X)[G], const TripDesc &d, float &Sc, int32_t &Dc) {
#pragma unroll
    for (int u = 0; u < G; u++) {
      W[4 * u + 0] = pull_one(a, hot_n, s_hot_addr, X[u].x, pol, pol_last);
      W[4 * u + 1] = pull_one(a, hot_n, s_hot_addr, X[u].y, pol, pol_last);
      W[4 * u + 2] = pull_one(a, hot_n, s_hot_addr, X[u].z, pol, pol_last);
      W[4 * u + 3] = pull_one(a, hot_n, s_hot_addr, X[u].w, pol, pol_last);
    }
    if (d.fin == 1) {
      const int64_t j = (int64_t)d.ref * 32 + lane;
      if (j < a.n_nz_rows && j >= a.n_band_rows) { Sc = __ldcs(a.scores + j); Dc = __ldcs(degs + j); }
    }
  };
  auto consume = [&](const float (&W)[4 * G], int32_t flag, int32_t ref, float Sc, int32_t Dc) {
#pragma unroll
    for (int q = 0; q < 4 * G; q++) acc = __fadd_rn(acc, W[q]);
    if ((flag & 3) == 1) {
      const int64_t j = (int64_t)ref * 32 + lane;
      if (j < a.n_band_rows) a.acc_main[j] = acc;          // whole slices: warp-uniform
      else if (j < a.n_nz_rows) pr_epilogue_pre(a, j, acc, err, Sc, Dc);
      acc = 0.f;
    } else if ((flag & 3) == 2) {
      a.partial[(size_t)ref * 32 + lane] = acc;
      acc = 0.f;
    }
  };

  dnext = it.next();
  if (dnext.n) {
    load_idx(I[0], dnext);
    // fill: trips 0 .. D-2
#pragma unroll
    for (int i = 0; i < D - 1; i++) {
      const TripDesc d = dnext;
      dnext = it.next();
      load_idx(I[(i + 1) & 1], dnext);
      gather(V[i], I[i & 1], d, S[i], Dg[i]);
      dflag[i] = d.fin | (d.n ? 4 : 0); dref[i] = d.ref;
    }
    // steady state, unrolled so that every register array index is a constant: step k issues the gathers of trip
    // base + D-1 + k and adds the values of trip base + k
    constexpr int U = (D % 2 == 0) ? D : 2 * D;
    for (;;) {
      bool stop = false;
#pragma unroll
      for (int k = 0; k < U; k++) {
        const TripDesc d = dnext;
        dnext = it.next();
        load_idx(I[(D + k) & 1], dnext);
        gather(V[(D - 1 + k) % D], I[(D - 1 + k) & 1], d, S[(D - 1 + k) % D], Dg[(D - 1 + k) % D]);
        dflag[(D - 1 + k) % D] = d.fin | (d.n ? 4 : 0); dref[(D - 1 + k) % D] = d.ref;
        consume(V[k % D], dflag[k % D], dref[k % D], S[k % D], Dg[k % D]);
        if (!dflag[(k + 1) % D]) { stop = true; break; }
      }
      if (stop) break;
    }
  }
  err = warp_sum(err);
  if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  contrib_flush(a);
}

template <int G, int D, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
pr_sell_pipe(SellArgs a) { pr_sell_pipe_body<G, D, THREADS>(a); }

// wide slices: the per-row partials of a slice's segments are added in a FIXED order -- the eight warps of a CTA each
// add a contiguous run of segments in column order, warp 0 then adds the eight run sums in run order.  (One warp per
// slice walked up to 500 dependent adds for the hub slices of Kron-26: 0.62 ms per iteration, a tenth of the gather.)
// Slices are taken widest first by CTA index, so the long ones start together.
__global__ void __launch_bounds__(256, 4)
pr_sell_finalize(SellArgs a) {
  __shared__ float s_run[8][32];
  if (*a.done) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double err = 0.0;
  for (int64_t h = blockIdx.x; h < a.n_heavy_slices; h += gridDim.x) {
    const int32_t s = a.heavy_slice[h];
    const int32_t f0 = a.heavy_first[h], f1 = a.heavy_first[h + 1];
    const int32_t n = f1 - f0;
    const int32_t q0 = f0 + (int32_t)((int64_t)n * w / 8), q1 = f0 + (int32_t)((int64_t)n * (w + 1) / 8);
    float acc = 0.f;
    int32_t q = q0;
    for (; q + 8 <= q1; q += 8) {                   // 8 independent loads, then the adds in segment order
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) t[u] = __ldcs(a.partial + (size_t)(q + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, t[u]);
    }
    for (; q < q1; q++) acc = __fadd_rn(acc, __ldcs(a.partial + (size_t)q * 32 + lane));
    s_run[w][lane] = acc;
    __syncthreads();
    if (w == 0) {
      float tot = s_run[0][lane];
#pragma unroll
      for (int u = 1; u < 8; u++) tot = __fadd_rn(tot, s_run[u][lane]);
      const int64_t j = (int64_t)s * 32 + lane;
      if (j < a.n_nz_rows) pr_epilogue(a, j, tot, err);
    }
    __syncthreads();
  }
  if (w == 0) {
    err = warp_sum(err);
    if (lane == 0) a.err_partial[a.err_slot0 + blockIdx.x] = err;
  }
  contrib_flush(a);
}

// rows without in-edges: score = base (src/pr/omp_base.cc:28-32 with an empty sum).  They reach that
// fixed point in the first iteration and never move again (L1 delta exactly 0), so this runs ONCE per
// solve and writes their constant contrib into both buffers.  DIRECTED graphs only: such a row may still have
// out-edges, so its OLD contrib must survive the first iteration's gathers -- this runs after them.  On a symmetric
// graph its contrib is x / 0, which nothing gathers, and pr_sell_load settles it on the way in.
__global__ void __launch_bounds__(256, 4)
pr_sell_isolated(SellArgs a, float *contrib_other) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double err = 0.0;
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damp, 0.f));
  for (int64_t j = a.n_nz_rows + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.rows; j += (int64_t)gridDim.x * blockDim.x) {
    const float old_score = __ldcs(a.scores + j);
    __stcs(a.scores + j, nw);
    err += (double)fabsf(__fsub_rn(nw, old_score));
    const int32_t deg = a.sout ? __ldcs(a.sout + j) : __ldcs(a.sdeg + j);
    // x / 0 without the division slow path (symmetric graphs: every such row has out-degree 0 too)
    const float cv = deg != 0 ? __fdiv_rn(nw, (float)deg)
                              : (nw > 0.f ? __int_as_float(0x7f800000) : nw < 0.f ? __int_as_float(0xff800000) : __int_as_float(0x7fc00000));
    const int64_t id = row_newid(a, j);
    contrib_store(a, id, cv, false);
    contrib_store_other(a, contrib_other, id, cv);
  }
  err = warp_sum(err);
  if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  contrib_flush(a);
}

// Scores into sorted order + the first contrib (src/pr/omp_base.cc:24-25).  Rows without in-edges are settled here,
// once per solve: score = base (:28-32 with an empty sum) is their fixed point from the first iteration on (L1 delta
// exactly 0 afterwards), so their |base - old| goes into the FIRST iteration's error partials and their constant
// contrib into both buffers; the iteration kernels never touch them.  (Not when max_iter = 0: no iteration runs.)
// abs_partial[warp] = sum |score| over the warp's rows: it bounds every partial sum of the solve (pr_run_sell) and so
// fixes the scale of the banded layout's fixed-point accumulators.
__global__ void __launch_bounds__(256, 4)
pr_sell_load(const float *__restrict__ scores_user, const int32_t *__restrict__ perm, SellArgs a, float *contrib_other,
             int settle_isolated, double *__restrict__ abs_partial) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double err = 0.0, tsum = 0.0;
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damp, 0.f));
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.rows; j += (int64_t)gridDim.x * blockDim.x) {
    const float sc = scores_user[perm[j]];
    const int32_t deg = a.sout ? __ldcs(a.sout + j) : __ldcs(a.sdeg + j);
    const int64_t id = row_newid(a, j);
    tsum += (double)fabsf(sc);
    if (j < a.n_nz_rows || !settle_isolated) {
      a.scores[j] = sc;
      contrib_store(a, id, __fdiv_rn(sc, (float)deg), true);
    } else {
      __stcs(a.scores + j, nw);
      err += (double)fabsf(__fsub_rn(nw, sc));
      // x / 0 without the division slow path (symmetric graphs: every such row has out-degree 0 too)
      const float cv = deg != 0 ? __fdiv_rn(nw, (float)deg)
                                : (nw > 0.f ? __int_as_float(0x7f800000) : nw < 0.f ? __int_as_float(0xff800000) : __int_as_float(0x7fc00000));
      contrib_store(a, id, cv, false);
      contrib_store_other(a, contrib_other, id, cv);
    }
  }
  if (settle_isolated) {
    err = warp_sum(err);
    if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  }
  tsum = warp_sum(tsum);
  if (lane == 0) abs_partial[warp] = tsum;
  contrib_flush(a);
}
__global__ void pr_sell_store(float *__restrict__ scores_user, const int32_t *__restrict__ perm,
                              const float *__restrict__ scores_sorted, int64_t rows) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < rows; j += (int64_t)gridDim.x * blockDim.x)
    scores_user[perm[j]] = scores_sorted[j];
}
// One GPU, symmetric order (new id of a vertex == its sorted row): coalesced stores, and the settled rows (half the
// vertices of a Kronecker graph) need no gather at all.
__global__ void pr_sell_store_gather(float *__restrict__ scores_user, const int32_t *__restrict__ newid,
                                     const float *__restrict__ scores_sorted, int64_t m, int64_t n_nz_rows, float iso, int use_iso) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < m; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = __ldcs(newid + v);
    __stcs(scores_user + v, (use_iso && j >= n_nz_rows) ? iso : __ldcs(scores_sorted + j));
  }
}
__global__ void gather_i32(const int32_t *__restrict__ src, const int32_t *__restrict__ perm, int32_t *__restrict__ dst, int64_t n) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) dst[j] = src[perm[j]];
}

__global__ void __launch_bounds__(256)
pr_reduce_err2(const double *__restrict__ partial, int n, double *err_trace, int iter, double eps, int32_t *done) {
  __shared__ double s[256];
  if (*done) return;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    err_trace[iter] = s[0];
    if (s[0] < eps) *done = iter + 1;
  }
}

// fixed-order sum of n partials -> *out (one block)
__global__ void __launch_bounds__(256)
pr_reduce_sum(const double *__restrict__ partial, int n, double *out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}
// multi-GPU over NCCL: the stop test on the all-reduced delta
__global__ void pr_check_done(const double *err_trace, int iter, double eps, int32_t *done) {
  if (!*done && err_trace[iter] < eps) *done = iter + 1;
}

// comm.cu
int pull_exchange(gdn_graph *g, float *contrib, double *err_slot);        // NCCL: two allgathers + allreduce of one double
int pull_peer_setup(gdn_graph *g);                                        // map the other GPUs' vectors (IPC / peer access)
bool pull_peer_ready(const gdn_graph *g);
void pull_peer_args(const gdn_graph *g, int buf_out, SellArgs &a);        // a.n_peers, a.peer_out, a.peer_other
// one block: reduce this GPU's partials, publish the sum to every GPU, barrier over the box (device-side flags), add the
// sums in rank order -> err_trace[iter] (and the stop flag when eps >= 0); with partial == nullptr only the barrier
int pull_peer_sync(gdn_graph *g, const double *partial, int n_partial, double *err_out, int iter, double eps, int32_t *done,
                   cudaStream_t s);
int pull_peer_check();

// Scale of the banded layout's fixed-point accumulators (band.cu).  Every partial sum of the solve is bounded by
// T = max(1, sum |scores_0|): sum_v |s_v^{k+1}| <= (1 - d) + d * sum_v |s_v^k| (src/pr/omp_base.cc:24-33 with every
// gathered vertex of out-degree >= 1) and a row sums a subset of contrib = s / deg.  2^e with T * 2^e <= 2^62 keeps the
// 64-bit accumulators in range; e is capped at 56 (the resolution round 1 shipped: 1.4e-17).  Returns 0 when the scores are
// not finite or so large that fewer than 30 fraction bits are left: the caller then runs the solve on the plain layout.
static double fix_scale_for(double tsum) {
  if (!(tsum >= 0.0) || !(tsum < 1e300)) return 0.0;
  const double bound = std::max(1.0, tsum) * 1.001;
  int e = 0;
  (void)frexp(bound, &e);                    // bound = f * 2^e, 0.5 <= f < 1  =>  bound < 2^e
  const int bits = std::min(56, 62 - e);
  if (bits < 30) return 0.0;
  return ldexp(1.0, bits);
}

// Buffers of the exact slices.  Exact-order mode: the slices' values in the layout of their index groups.  Default mode: the
// row-major value blocks and per-block tables of ordered_sum.cuh.
static int exact_setup(gdn_graph *g) {
  PullLayout &L = g->pull;
  if (L.exact) {
    GDN_CUDA(cudaMalloc((void **)&L.exact_vals, sizeof(float4) * (size_t)L.h_slice_ptr[L.n_exact] + 256));
    g->device_bytes += sizeof(float4) * (size_t)L.h_slice_ptr[L.n_exact];
    return GDN_OK;
  }
  std::vector<uint32_t> base((size_t)L.n_exact + 1), tbase((size_t)L.n_exact + 1);
  uint64_t tot = 0, tiles = 0;
  for (int32_t e = 0; e < L.n_exact; e++) {
    base[e] = (uint32_t)tot;
    tbase[e] = (uint32_t)tiles;
    const uint32_t ngl = (L.h_slice_ptr[e + 1] - L.h_slice_ptr[e]) >> 5;
    const uint32_t nb = (ngl + kOrdBlockGroups - 1) / kOrdBlockGroups;
    tot += 32ull * nb;
    tiles += nb;
    if (tot > 0x7fffffffull) { L.n_exact = e; break; }                 // (cannot happen below 2^31 * 512 columns; keeps the 32-bit block index honest)
  }
  base[L.n_exact] = (uint32_t)tot;
  tbase[L.n_exact] = (uint32_t)tiles;
  L.x_blocks = (int32_t)tot;
  L.x_tiles = (int32_t)tiles;
  GDN_CUDA(cudaMalloc((void **)&L.x_blk_base, sizeof(uint32_t) * 2 * base.size()));
  L.x_tile_base = L.x_blk_base + base.size();
  GDN_CUDA(cudaMemcpyAsync(L.x_blk_base, base.data(), sizeof(uint32_t) * base.size(), cudaMemcpyHostToDevice, lib().stream));
  GDN_CUDA(cudaMemcpyAsync(L.x_tile_base, tbase.data(), sizeof(uint32_t) * tbase.size(), cudaMemcpyHostToDevice, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  GDN_CUDA(cudaMalloc((void **)&L.exact_vals, sizeof(float4) * (size_t)tot * kOrdBlockGroups + 256));
  GDN_CUDA(cudaMemsetAsync(L.exact_vals, 0, sizeof(float4) * (size_t)tot * kOrdBlockGroups + 256, lib().stream));   // (padding blocks stay zero)
  GDN_CUDA(cudaMalloc((void **)&L.x_S, sizeof(double) * (size_t)tot));
  GDN_CUDA(cudaMalloc((void **)&L.x_mx, sizeof(uint32_t) * (size_t)tot));
  GDN_CUDA(cudaMalloc((void **)&L.x_Q, sizeof(uint32_t) * (size_t)tot));
  GDN_CUDA(cudaMalloc((void **)&L.x_plan, (size_t)tot + 16));
  g->device_bytes += (sizeof(float4) * kOrdBlockGroups + 17) * (size_t)tot;
  return GDN_OK;
}
static ExactArgs exact_args(const PullLayout &L) {
  ExactArgs x;
  x.n_exact = L.n_exact; x.n_blocks_total = L.x_blocks; x.blk_base = L.x_blk_base; x.vals = L.exact_vals;
  x.S = L.x_S; x.mx = L.x_mx; x.plan = L.x_plan; x.Q = L.x_Q;
  return x;
}

int pr_run_sell(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st) {
  PullLayout &L = g->pull;
  GDN_CHECK(pull_build_sell(g));
  // banded shared-memory layout of the heavy rows (band.cu): resident graphs; GDN_PR_BANDS=0 turns it off
  const char *e_bands = getenv("GDN_PR_BANDS");
  const bool bands_off = (e_bands && atoi(e_bands) <= 0) || L.exact;
  if (!g->one_shot && !bands_off) GDN_CHECK(band_build(g));
  bool banded = L.band.built && !bands_off;
  const BandLayout &bd = L.band;
  cudaStream_t s = lib().stream;
  const int sm = lib().sm_count;
  const int wpc = kSellThreads / 32;             // err_partial slots per CTA
  const bool multi = L.P > 1;
  const int32_t max_heavy_slices = std::max(L.n_heavy_slices, L.band.built ? bd.n_heavy_slices : 0);
  const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)max_heavy_slices, (int64_t)sm * 8));   // one CTA per wide slice
  const int igrid = sm * 8;                        // pr_sell_load's grid: its warps own the error partials of the settled rows
  const int bgrid = L.band.built ? band_finalize_grid(g) : 0;
  const int xgrid = (L.n_exact * 32 + 7) / 8;       // pr_exact_combine: one warp per exact row
  const int n_partial = sm * wpc + fgrid * 8 + igrid * 8 + bgrid * 8 + xgrid * 8;
  if (!g->contrib[0]) {
    GDN_CUDA(cudaMalloc((void **)&g->contrib[0], sizeof(float) * (L.Mp + 64)));
    GDN_CUDA(cudaMalloc((void **)&g->contrib[1], sizeof(float) * (L.Mp + 64)));
    GDN_CUDA(cudaMalloc((void **)&g->scores_sorted, sizeof(float) * std::max<int64_t>(L.rows, 1)));
    GDN_CUDA(cudaMalloc((void **)&g->err_trace, sizeof(double) * (GDN_MAX_PR_ITER + 8)));
    GDN_CUDA(cudaMalloc((void **)&g->pr_done, sizeof(int32_t)));
    GDN_CUDA(cudaMalloc((void **)&g->pr_work, sizeof(int32_t)));
    GDN_CUDA(cudaMalloc((void **)&g->abs_partial, sizeof(double) * igrid * 8));
    GDN_CUDA(cudaMemsetAsync(g->contrib[0], 0, sizeof(float) * (L.Mp + 64), s));
    GDN_CUDA(cudaMemsetAsync(g->contrib[1], 0, sizeof(float) * (L.Mp + 64), s));
    g->device_bytes += sizeof(float) * (2 * L.Mp + L.rows);
    if (g->out_degree) {                       // directed graph: out-degree in sorted row order
      GDN_CUDA(cudaMalloc((void **)&L.sout, sizeof(int32_t) * std::max<int64_t>(L.rows, 1)));
      gather_i32<<<sm * 8, 256, 0, s>>>(g->out_degree, L.perm, L.sout, L.rows);
    }
  }
  // row partition: map the other GPUs' vectors once per graph (GDN_PR_NCCL=1 keeps the NCCL collectives instead; they are
  // also what runs when the mapping is not possible)
  if (multi && (gang_worker() || !getenv("GDN_PR_NCCL")) && pull_peer_setup(g) != GDN_OK) cudaGetLastError();
  const bool peer = multi && pull_peer_ready(g);
  if (multi && !peer && gang_worker()) { set_error("the GPUs of the gang cannot map each other's memory"); return GDN_ERR_CUDA; }
  if (g->n_err_partial < n_partial) {
    if (g->err_partial) GDN_CUDA(cudaFree(g->err_partial));
    GDN_CUDA(cudaMalloc((void **)&g->err_partial, sizeof(double) * n_partial));
    g->n_err_partial = n_partial;
  }
  if (max_iter > GDN_MAX_PR_ITER - 1) max_iter = GDN_MAX_PR_ITER - 1;
  const size_t smem = std::max(sizeof(float) * (size_t)L.H, (L.exact && L.n_exact > 0) ? kChainSmem : (size_t)0);
  if (L.n_exact > 0 && !L.exact_vals) GDN_CHECK(exact_setup(g));
  // L2 residency tiers of the gathered vector (see pull_one): 48 MB measured best at Kron-26 (32: +4 %, 64: +5 %, 96: +17 %)
  const char *e_warm = getenv("GDN_PR_WARM_MB");
  const int64_t warm_ids = (int64_t)(e_warm ? atoi(e_warm) : 48) * (1 << 20) / 4;
  // one index group per trip, two trips of gathers in flight, 32 warps per SM: best of the sweep (profiles/r1_pr_pipe_sweep.txt)
  void (*kern)(SellArgs) = pr_sell_pipe<1, 2, kSellThreads>;
  GDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  SellArgs a = {};
  a.sell = L.sell; a.slice_ptr = L.slice_ptr; a.chunk_slice = L.chunk_slice; a.n_chunks = L.n_chunks;
  a.heavy_seg = L.heavy_seg; a.heavy_slice = L.heavy_slice; a.heavy_first = L.heavy_first;
  a.n_heavy_segs = L.n_heavy_segs; a.n_heavy_slices = L.n_heavy_slices; a.partial = L.partial;
  a.scores = g->scores_sorted; a.sdeg = L.sdeg; a.sout = L.sout; a.rowid = L.rowid;
  a.n_nz_rows = L.n_nz_rows; a.rows = L.rows; a.H = (int32_t)L.H; a.hot_n = (int32_t)L.H; a.Hp = L.Hp; a.Wc = L.Wc; a.rank = L.R;
  a.base = (1.0f - damp) / (float)(int32_t)g->m;            // src/pr/omp_base.cc:16
  a.damp = damp; a.err_partial = g->err_partial; a.done = g->pr_done;
  a.strict_chain = L.exact ? 1 : 0;
  a.group_ch = L.group_ch; a.n_exact = L.n_exact; a.work_counter = g->pr_work; a.exact_vals = L.exact_vals;
  { const char *e = getenv("GDN_PR_MAIN_PF"); a.pf_trips = e ? atoi(e) : 0; }
  a.P = L.P; a.inv_wc = L.Wc > 0 ? 1.0f / (float)L.Wc : 0.f;
  // the warm budget is shared by the ranks' slices (tier_id): every rank's hottest cold ids stay L2-resident
  a.warm = (int32_t)std::min<int64_t>(L.P > 1 ? L.H + std::max<int64_t>(warm_ids - L.H, 0) / L.P : warm_ids, 0x7fffffff);
  double *h_err = (double *)lib().pinned;          // [GDN_MAX_PR_ITER]: one slot per iteration; the last one takes sum |scores_0|
  double *h_tsum = h_err + (GDN_MAX_PR_ITER - 1);
  int64_t launches = 0;

  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  GDN_CUDA(cudaMemsetAsync(g->pr_done, 0, sizeof(int32_t), s));
  GDN_CUDA(cudaMemsetAsync(g->err_partial, 0, sizeof(double) * n_partial, s));
  a.contrib_out = g->contrib[0];
  if (peer) pull_peer_args(g, 0, a);
  // rows without in-edges: settled by pr_sell_load (symmetric graph) or by pr_sell_isolated after the first iteration
  const bool have_iso = max_iter > 0 && L.rows > L.n_nz_rows;
  const int settle = (have_iso && !a.sout) ? 1 : 0;
  a.err_slot0 = sm * wpc + fgrid * 8;
  pr_sell_load<<<igrid, 256, 0, s>>>(d_scores, L.perm, a, g->contrib[1], settle, g->abs_partial);
  launches++;
  double fix_scale = 0.0;
  if (banded) {
    // the fixed-point range of the band accumulators follows the caller's scores (src/pr/omp_base.cc:24 accepts any vector)
    double *d_tsum = g->err_trace + GDN_MAX_PR_ITER;
    pr_reduce_sum<<<1, 256, 0, s>>>(g->abs_partial, igrid * 8, d_tsum);
    launches++;
    if (peer) GDN_CHECK(pull_peer_sync(g, d_tsum, 1, d_tsum, 0, -1.0, nullptr, s));
    else GDN_CHECK(pull_exchange(g, g->contrib[0], d_tsum));
    GDN_CUDA(cudaMemcpyAsync(h_tsum, d_tsum, sizeof(double), cudaMemcpyDeviceToHost, s));
    GDN_CUDA(cudaStreamSynchronize(s));
    fix_scale = fix_scale_for(*h_tsum);
    if (fix_scale == 0.0) banded = false;          // out of range for the accumulators: this solve walks the plain layout
  } else {
    if (peer) GDN_CHECK(pull_peer_sync(g, nullptr, 0, nullptr, 0, -1.0, nullptr, s));
    else GDN_CHECK(pull_exchange(g, g->contrib[0], nullptr));
  }
  const int32_t n_heavy_slices = banded ? bd.n_heavy_slices : L.n_heavy_slices;
  if (banded) {
    a.sell = bd.sell; a.slice_ptr = bd.slice_ptr; a.chunk_slice = bd.chunk_slice; a.n_chunks = bd.n_chunks;
    a.heavy_seg = bd.heavy_seg; a.heavy_slice = bd.heavy_slice; a.heavy_first = bd.heavy_first;
    a.n_heavy_segs = bd.n_heavy_segs; a.n_heavy_slices = bd.n_heavy_slices; a.partial = bd.partial;
    a.n_band_rows = bd.n_rows; a.acc_main = bd.acc_main;
    GDN_CHECK(band_solve_begin(g, s));
  }
  if (st) st->pr_layout = banded ? (bd.seg ? 2 : 1) : 0;

  // The host runs kLook iterations ahead of the device: iteration k + kLook is queued before the delta of iteration k is
  // read (the kernels of an iteration queued after convergence return at once on the device-side `done` flag), so the
  // stream never drains between iterations.  Over NCCL the stop test needs the all-reduced delta and the collectives of a
  // dead iteration would still run: no look-ahead there.
  const int kLook = (multi && !peer) ? 0 : 2;
  constexpr int kRing = 4;
  if (!lib().pr_ev[0])
    for (int i = 0; i < kRing; i++) GDN_CUDA(cudaEventCreateWithFlags(&lib().pr_ev[i], cudaEventDisableTiming));
  // GDN_PR_KTIME=1: CUDA-event time of every launch of the third iteration, in place (not under a profiler), on stderr
  static thread_local cudaEvent_t kt_ev[16];
  static thread_local bool kt_init = false;
  const bool ktime = getenv("GDN_PR_KTIME") != nullptr;
  if (ktime && !kt_init) { for (auto &e : kt_ev) cudaEventCreate(&e); kt_init = true; }
  int kt_n = 0;
  const char *kt_name[16] = {};
  auto probe = [&](int iter, const char *name) {
    if (ktime && iter == 2 && kt_n < 16) { cudaEventRecord(kt_ev[kt_n], s); kt_name[kt_n++] = name; }
  };
  auto enqueue = [&](int iter) -> int {
    const int cur = iter & 1;
    a.contrib_in = g->contrib[cur];
    a.contrib_out = g->contrib[cur ^ 1];
    if (peer) pull_peer_args(g, cur ^ 1, a);
    a.err_slot0 = 0;
    kev_begin();
    probe(iter, "start");
    if (banded) {
      GDN_CHECK(band_launch(g, a, fix_scale, s));
      launches += band_launches(g);
      probe(iter, "band sums");
    }
    GDN_CUDA(cudaMemsetAsync(g->pr_work, 0, sizeof(int32_t), s));
    if (L.n_exact > 0 && L.exact) {                    // exact-order mode: values for the chain warps of the main kernel
      pr_exact_gather_sell<<<sm * 8, 256, 0, s>>>(a, L.exact_vals, L.h_slice_ptr[L.n_exact]);
      launches++;
    } else if (L.n_exact > 0) {                       // default mode: block-parallel ordered sum, passes 1 - 3 (ordered_sum.cuh)
      const ExactArgs xa = exact_args(L);
      pr_exact_gather<<<sm * 8, 256, 0, s>>>(a, xa);
      probe(iter, "exact gather");
      pr_exact_plan<<<(L.n_exact * 32 + 7) / 8, 256, 0, s>>>(a, xa);
      exact_qsum<<<sm * 8, 256, 0, s>>>(a.done, xa);
      launches += 3;
      probe(iter, "exact plan + qsum");
    }
    kern<<<sm, kSellThreads, smem, s>>>(a);
    probe(iter, "main sums");
    if (!banded) kev_end();
    launches++;
    if (n_heavy_slices > 0) {
      a.err_slot0 = sm * wpc;
      pr_sell_finalize<<<fgrid, 256, 0, s>>>(a);
      launches++;
    }
    if (L.n_exact > 0 && !L.exact) {                  // pass 4: the exact rows' sums, in order -> epilogue (or acc_main of a band row)
      a.err_slot0 = sm * wpc + fgrid * 8 + igrid * 8 + bgrid * 8;
      pr_exact_combine<<<xgrid, 256, 0, s>>>(a, exact_args(L));
      launches++;
      probe(iter, "exact combine");
    }
    if (banded) {
      // the iteration of the banded layout is band sums + main sums + the two finalize launches: timed as one
      a.err_slot0 = sm * wpc + fgrid * 8 + igrid * 8;
      GDN_CHECK(band_finalize_launch(g, a, fix_scale, bgrid, s));
      probe(iter, "finalize");
      kev_end();
      launches++;
    }
    if (have_iso && !settle && iter == 0) {
      a.err_slot0 = sm * wpc + fgrid * 8;
      pr_sell_isolated<<<igrid, 256, 0, s>>>(a, g->contrib[cur]);
      launches++;
    } else if (have_iso && iter == 1) {            // the settled rows' deltas belonged to the first iteration
      GDN_CUDA(cudaMemsetAsync(g->err_partial + sm * wpc + fgrid * 8, 0, sizeof(double) * igrid * 8, s));
    }
    if (peer) {
      GDN_CHECK(pull_peer_sync(g, g->err_partial, n_partial, g->err_trace, iter, eps, g->pr_done, s));
      launches++;
    } else {
      pr_reduce_err2<<<1, 256, 0, s>>>(g->err_partial, n_partial, g->err_trace, iter, multi ? -1.0 : eps, g->pr_done);
      launches++;
      if (multi) {
        GDN_CHECK(pull_exchange(g, g->contrib[cur ^ 1], g->err_trace + iter));
        pr_check_done<<<1, 1, 0, s>>>(g->err_trace, iter, eps, g->pr_done);
        launches++;
      }
    }
    GDN_CUDA(cudaMemcpyAsync(h_err + iter, g->err_trace + iter, sizeof(double), cudaMemcpyDeviceToHost, s));
    GDN_CUDA(cudaEventRecord(lib().pr_ev[iter % kRing], s));
    return GDN_OK;
  };
  int queued = 0, checked = 0, last = max_iter;      // last = index of the iteration that met the stop test (max_iter: none did)
  for (;;) {
    while (queued < max_iter && queued - checked <= kLook) { GDN_CHECK(enqueue(queued)); queued++; }
    if (checked == queued) break;
    GDN_CUDA(cudaEventSynchronize(lib().pr_ev[checked % kRing]));
    const double err = h_err[checked];
    if (st) st->pr_err[checked] = err;
    if (err < eps) { last = checked; break; }                // src/pr/omp_base.cc:36
    checked++;
  }
  if (L.P == 1 && L.symmetric_order && L.rows == g->m && !L.rowid)
    pr_sell_store_gather<<<sm * 8, 256, 0, s>>>(d_scores, L.newid, g->scores_sorted, g->m, L.n_nz_rows,
                                                 a.base /* = base + damp * 0 */, settle);
  else
    pr_sell_store<<<sm * 8, 256, 0, s>>>(d_scores, L.perm, g->scores_sorted, L.rows);
  launches++;
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (peer) GDN_CHECK(pull_peer_check());
  if (ktime && kt_n > 1) {
    fprintf(stderr, "[gdn] iteration 3, per launch (ms):");
    for (int i = 1; i < kt_n; i++) { float ms = 0; cudaEventElapsedTime(&ms, kt_ev[i - 1], kt_ev[i]); fprintf(stderr, "  %s %.3f", kt_name[i], ms); }
    fprintf(stderr, "\n");
  }
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms;
    st->kernel_launches = launches;
    st->iterations = last + 1;                               // printf("iterations = %d", iter+1), :38
    kev_collect(st, std::min(last + 1, max_iter));           // (iterations queued past the stop test were no-ops)
  }
  return GDN_OK;
}

}  // namespace gdn
