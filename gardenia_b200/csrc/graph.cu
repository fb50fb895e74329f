// graph.cu -- library state, device-resident CSR (gdn_graph) and the C-ABI
// entry points of include/gdn_b200.h that touch the device.
//
// gdn_graph_create replaces the cudaMalloc+cudaMemcpy prologue that every
// reference CUDA solver repeats (e.g. src/pr/warp.cu:137-155): offsets are
// narrowed to 32 bits on the device when the local nnz allows it (the reference
// narrows uint64 -> int unconditionally), column indices stay global.
#include "common.cuh"
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <vector>
#include <algorithm>

namespace gdn {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static thread_local Lib *tl_lib = nullptr;
static Lib &process_lib() { static Lib l; return l; }
Lib &lib() { return tl_lib ? *tl_lib : process_lib(); }
void lib_bind(Lib *l) { tl_lib = l; }

void trace(const char *label) {
  static const bool on = getenv("GDN_TRACE") != nullptr;
  if (!on) return;
  static double last = 0;
  const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  fprintf(stderr, "[gdn trace] %-28s +%9.2f ms\n", label, last == 0 ? 0.0 : now - last);
  last = now;
}

int ensure_init() {
  if (lib().inited) return GDN_OK;
  return gdn_init(0);
}

void *host_arena(int slot, size_t bytes) {
  Lib &l = lib();
  if (l.arena_bytes[slot] >= bytes) return l.arena[slot];
  if (l.arena[slot]) cudaFreeHost(l.arena[slot]);
  l.arena[slot] = nullptr; l.arena_bytes[slot] = 0;
  const size_t want = bytes + bytes / 8;
  if (cudaHostAlloc(&l.arena[slot], want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); l.arena[slot] = nullptr; return nullptr; }
  l.arena_bytes[slot] = want;
  return l.arena[slot];
}

// Queue the pieces of a chunked column upload that upload_csr_begin held back (see there).
int col_upload_rest(gdn_graph *g) {
  cudaStream_t cs = lib().copy_stream;
  const uint64_t piece = 64ull << 20;                    // entries (multiple of 1024)
  while (g->col_next < g->col_total) {
    const uint64_t e0 = g->col_next, e1 = std::min<uint64_t>(g->col_total, e0 + piece);
    GDN_CUDA(cudaMemcpyAsync(g->col_dev + e0, g->col_host + e0, sizeof(int32_t) * (e1 - e0), cudaMemcpyHostToDevice, cs));
    cudaEvent_t ev;
    GDN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    GDN_CUDA(cudaEventRecord(ev, cs));
    g->col_ev.push_back(ev);
    g->col_end.push_back(e1);
    g->col_next = e1;
  }
  return GDN_OK;
}

int build_schedule(gdn_graph *g, DevCsr &c);
template <typename HostOffT>
int pull_prepare(gdn_graph *g, const HostOffT *row_off, const HostOffT *key_off);   // pull.cu
int bfs_run(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st);
int pr_run(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st);
int spmv_run(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y, gdn_stats *st);
int pull_peer_release(gdn_graph *g);       // comm.cu
void gang_stop();                          // comm.cu

// in[i] - base -> out[i]
template <typename InT, typename OutT>
__global__ void convert_offsets(const InT *__restrict__ in, OutT *__restrict__ out, int64_t n, uint64_t base) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (OutT)((uint64_t)in[i] - base);
}

template <typename OffT>
__global__ void row_lengths(const OffT *__restrict__ rp, int32_t *__restrict__ deg, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    deg[i] = (int32_t)(rp[i + 1] - rp[i]);
}

// Checks monotone offsets and 0 <= col < m; flag[0] != 0 on violation.
template <typename OffT>
__global__ void validate_csr(const OffT *__restrict__ rp, const int32_t *__restrict__ col, int64_t rows,
                             uint64_t nnz, int64_t m, int *flag, bool check_cols) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = tid; r < rows; r += nth)
    if (rp[r] > rp[r + 1]) atomicExch(flag, 1);
  if (tid == 0 && (rp[0] != 0 || (uint64_t)rp[rows] != nnz)) atomicExch(flag, 2);
  if (!check_cols) return;                 // the streaming layout build checks every column id as it consumes it
  for (uint64_t e = tid; e < nnz; e += nth) {
    const int32_t c = col[e];
    if (c < 0 || c >= m) atomicExch(flag, 3);
  }
}

// Upload of one CSR in two halves so that host-side preprocessing can run while the (pinned) column
// array is still crossing PCIe: upload_csr_begin queues the copies and the validation kernel on the
// library stream, upload_csr_end waits, checks the verdict and builds the row-block schedule.
struct PendingUpload {
  void *tmp = nullptr;
  int *flag = nullptr;
  int *h_flag = nullptr;      // slot in the pinned mailbox
  bool want_schedule = false;
  bool chunked = false;       // column array copied piecewise on the copy stream (gdn_graph::col_ev)
};

template <typename HostOffT>
static int upload_csr_begin(gdn_graph *g, DevCsr &c, const HostOffT *h_rowptr, const int32_t *h_col, int64_t row_lo,
                            int64_t row_hi, bool want_schedule, int slot, PendingUpload &pu) {
  cudaStream_t st = lib().stream;
  c.rows = row_hi - row_lo;
  const uint64_t base = (uint64_t)h_rowptr[row_lo];
  c.nnz = (uint64_t)h_rowptr[row_hi] - base;
  c.off64 = c.nnz >= 0xffff0000ull;
  const size_t off_bytes = (c.off64 ? 8 : 4) * (size_t)(c.rows + 1);
  GDN_CUDA(cudaMalloc(&c.rowptr, off_bytes));
  const size_t col_bytes = sizeof(int32_t) * c.nnz + 256;          // slack for 256-bit over-reads
  GDN_CUDA(cudaMalloc((void **)&c.col, col_bytes));
  g->device_bytes += off_bytes + col_bytes;
  // offsets: stage the host-typed slice, narrow/rebase on the device
  HostOffT *tmp = nullptr;
  GDN_CUDA(cudaMalloc((void **)&tmp, sizeof(HostOffT) * (c.rows + 1)));
  pu.tmp = tmp;
  GDN_CUDA(cudaMemcpyAsync(tmp, h_rowptr + row_lo, sizeof(HostOffT) * (c.rows + 1), cudaMemcpyHostToDevice, st));
  const int grid = (int)std::min<int64_t>((c.rows + 256) / 256, (int64_t)lib().sm_count * 8);
  if (c.off64) convert_offsets<HostOffT, uint64_t><<<grid, 256, 0, st>>>(tmp, (uint64_t *)c.rowptr, c.rows + 1, base);
  else convert_offsets<HostOffT, uint32_t><<<grid, 256, 0, st>>>(tmp, (uint32_t *)c.rowptr, c.rows + 1, base);
  GDN_CUDA(cudaMalloc((void **)&pu.flag, sizeof(int)));
  GDN_CUDA(cudaMemsetAsync(pu.flag, 0, sizeof(int), st));
  GDN_CUDA(cudaMemsetAsync((char *)c.col + sizeof(int32_t) * c.nnz, 0, 256, st));
  const bool chunked = lib().stream_fill && want_schedule && c.nnz > 0;
  if (chunked) {
    // One-shot PageRank: the column array crosses PCIe in 256 MB pieces on the copy stream, one event per piece, and
    // pull_prepare launches the layout build piece by piece behind them (pull.cu sell_scatter).
    cudaStream_t cs = lib().copy_stream;
    cudaEvent_t ready;
    GDN_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    GDN_CUDA(cudaEventRecord(ready, st));                 // allocations / memsets above are ordered before the copies
    GDN_CUDA(cudaStreamWaitEvent(cs, ready, 0));
    GDN_CUDA(cudaEventDestroy(ready));
    // Copies execute in issue order on the one host->device engine, whatever their stream: the row order that the
    // host computes meanwhile (pull_prepare, ~1 ns per vertex) would queue behind EVERY piece issued now and the build
    // could not start before the whole array is in.  So only ~64 bytes per vertex of pieces go now -- about the time
    // the host needs -- and pull_prepare issues the rest right after its own uploads (col_upload_rest).
    const uint64_t piece = 64ull << 20;                    // entries (multiple of 1024)
    const uint64_t first = std::min<uint64_t>(c.nnz, std::max<uint64_t>(piece, ((uint64_t)g->m * 16 + piece - 1) / piece * piece));
    g->col_host = h_col + base; g->col_dev = c.col; g->col_next = 0; g->col_total = first;
    GDN_CHECK(col_upload_rest(g));
    g->col_total = c.nnz;
    g->col_flag = pu.flag;
  } else {
    GDN_CUDA(cudaMemcpyAsync(c.col, h_col + base, sizeof(int32_t) * c.nnz, cudaMemcpyHostToDevice, st));
  }
  const int vgrid = lib().sm_count * 8;
  if (c.off64) validate_csr<uint64_t><<<vgrid, 256, 0, st>>>((const uint64_t *)c.rowptr, c.col, c.rows, c.nnz, g->m, pu.flag, !chunked);
  else validate_csr<uint32_t><<<vgrid, 256, 0, st>>>((const uint32_t *)c.rowptr, c.col, c.rows, c.nnz, g->m, pu.flag, !chunked);
  pu.chunked = chunked;
  pu.h_flag = (int *)((char *)lib().pinned + 1024) + slot;
  *pu.h_flag = 0;
  if (!chunked) GDN_CUDA(cudaMemcpyAsync(pu.h_flag, pu.flag, sizeof(int), cudaMemcpyDeviceToHost, st));   // chunked: read in _end, after the build
  pu.want_schedule = want_schedule;
  return GDN_OK;
}

static int upload_csr_end(gdn_graph *g, DevCsr &c, PendingUpload &pu) {
  if (!pu.flag) return GDN_OK;
  if (pu.chunked) {
    // every piece must have landed (and every layout kernel that consumed one has been queued on the stream)
    GDN_CHECK(col_upload_rest(g));
    for (cudaEvent_t ev : g->col_ev) { cudaStreamWaitEvent(lib().stream, ev, 0); }
    GDN_CUDA(cudaMemcpyAsync(pu.h_flag, pu.flag, sizeof(int), cudaMemcpyDeviceToHost, lib().stream));
  }
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  if (pu.chunked) {
    for (cudaEvent_t ev : g->col_ev) cudaEventDestroy(ev);
    g->col_ev.clear(); g->col_end.clear(); g->col_flag = nullptr; g->col_host = nullptr; g->col_dev = nullptr;
    if (!g->pull.sell && *pu.h_flag == 0) {
      // the layout build did not run (no SELL layout for this graph): nobody has looked at the column ids yet
      int *d_flag = pu.flag;
      if (c.off64) validate_csr<uint64_t><<<lib().sm_count * 8, 256, 0, lib().stream>>>((const uint64_t *)c.rowptr, c.col, c.rows, c.nnz, g->m, d_flag, true);
      else validate_csr<uint32_t><<<lib().sm_count * 8, 256, 0, lib().stream>>>((const uint32_t *)c.rowptr, c.col, c.rows, c.nnz, g->m, d_flag, true);
      GDN_CUDA(cudaMemcpyAsync(pu.h_flag, pu.flag, sizeof(int), cudaMemcpyDeviceToHost, lib().stream));
      GDN_CUDA(cudaStreamSynchronize(lib().stream));
    }
  }
  const int hflag = *pu.h_flag;
  cudaFree(pu.tmp);
  cudaFree(pu.flag);
  pu.tmp = nullptr; pu.flag = nullptr;
  GDN_CUDA(cudaGetLastError());
  if (hflag) {
    set_error("malformed CSR (%s)", hflag == 1 ? "row offsets not monotone" : hflag == 2 ? "offset ends" : "column index out of range");
    return GDN_ERR_GRAPH;
  }
  // one-shot PageRank with its SELL layout already built: the row-block schedule of the plain CSR is never used
  if (pu.want_schedule && !(lib().stream_fill && g->pull.sell)) GDN_CHECK(build_schedule(g, c));
  return GDN_OK;
}

static void free_csr(DevCsr &c) {
  cudaFree(c.rowptr); cudaFree(c.col); cudaFree(c.chunk_row); cudaFree(c.heavy_row);
  cudaFree(c.heavy_first); cudaFree(c.heavy_seg); cudaFree(c.heavy_partial);
  c = DevCsr();
}

template <typename HostOffT>
static int graph_create_t(int64_t m, int64_t nnz, const HostOffT *out_rowptr, const int32_t *out_colidx,
                          const HostOffT *in_rowptr, const int32_t *in_colidx, int64_t row_lo, int64_t row_hi,
                          gdn_graph **out) {
  GDN_CHECK(ensure_init());
  if (!out || m <= 0 || m > 0x7fffffffll || nnz < 0 || row_lo < 0 || row_hi > m || row_lo > row_hi) {
    set_error("gdn_graph_create: bad sizes (m=%lld nnz=%lld rows=[%lld,%lld))", (long long)m, (long long)nnz,
              (long long)row_lo, (long long)row_hi);
    return GDN_ERR_ARG;
  }
  if (!out_rowptr && !in_rowptr) { set_error("gdn_graph_create: no CSR given"); return GDN_ERR_ARG; }
  gdn_graph *g = new gdn_graph();
  g->m = m; g->row_lo = row_lo; g->row_hi = row_hi;
  int rc = GDN_OK;
  const auto t_create = std::chrono::steady_clock::now();
  trace("graph_create: begin");
  PendingUpload pu_out, pu_in;
  const bool sym = out_rowptr && (!in_rowptr || (in_rowptr == out_rowptr && in_colidx == out_colidx));
  if (sym) {
    // one CSR serves both directions (symmetrized graph, include/csr_graph.h:241-246)
    rc = upload_csr_begin(g, g->out, out_rowptr, out_colidx, row_lo, row_hi, true, 0, pu_out);
    g->symmetric = true;
    g->has_out = g->has_in = true;
  } else if (!out_rowptr) {
    // pull-only graph: PageRank additionally needs gdn_graph_set_out_degree
    rc = upload_csr_begin(g, g->in, in_rowptr, in_colidx, row_lo, row_hi, true, 1, pu_in);
    g->has_in = true;
  } else {
    rc = upload_csr_begin(g, g->out, out_rowptr, out_colidx, row_lo, row_hi, false, 0, pu_out);
    if (rc == GDN_OK) rc = upload_csr_begin(g, g->in, in_rowptr, in_colidx, row_lo, row_hi, true, 1, pu_in);
    g->has_out = g->has_in = true;
  }
  trace("graph_create: CSR uploaded");
  if (rc == GDN_OK && out_rowptr && g->has_in) {
    // log-scale out-degree class of EVERY vertex (the host still has the full offsets here; a row
    // partition only keeps its own rows on the device): class 0 = hubs ... 7 = leaves, thresholds
    // avg_degree * 4^j.  bfs.cu orders each bottom-up row hubs-first with it.
    std::vector<uint8_t> cls((size_t)m);
    const double avg = std::max(1.0, (double)nnz / (double)m);
#pragma omp parallel for
    for (int64_t v = 0; v < m; v++) {
      const double d = (double)(out_rowptr[v + 1] - out_rowptr[v]) / avg;
      int c = 7;
      for (double t = 0.25; c > 0 && d >= t; t *= 4.0) c--;
      cls[v] = (uint8_t)c;
    }
    if (cudaMalloc((void **)&g->deg_class, (size_t)m + 16) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); rc = GDN_ERR_NOMEM; }
    else {
      g->device_bytes += (size_t)m;
      if (cudaMemcpy(g->deg_class, cls.data(), (size_t)m, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("deg_class upload failed"); rc = GDN_ERR_CUDA; }
    }
  }
  trace("graph_create: degree classes");
  if (rc == GDN_OK && g->has_in) {
    // degree-sorted SELL layout for the PageRank pull (host part; the SELL array itself is built on first use)
    const HostOffT *row_off = (g->symmetric || !in_rowptr) ? out_rowptr : in_rowptr;
    if (!out_rowptr) row_off = in_rowptr;
    const HostOffT *key_off = out_rowptr ? out_rowptr : row_off;
    rc = pull_prepare<HostOffT>(g, row_off, key_off);
  }
  trace("graph_create: pull layout");
  // the copies queued by upload_csr_begin have been running under the host work above
  {
    const int r1 = upload_csr_end(g, g->out, pu_out), r2 = upload_csr_end(g, g->in, pu_in);
    if (rc == GDN_OK) rc = r1 != GDN_OK ? r1 : r2;
  }
  if (rc == GDN_OK && sym) g->in = g->out;      // shallow alias; freed once
  if (rc == GDN_OK && !sym && out_rowptr) {
    // PageRank divides by the OUT degree (src/pr/omp_base.cc:25)
    cudaStream_t st = lib().stream;
    const int64_t rows = row_hi - row_lo;
    if (cudaMalloc((void **)&g->out_degree, sizeof(int32_t) * (rows + 1)) != cudaSuccess) {
      cudaGetLastError();
      set_error("out of device memory");
      rc = GDN_ERR_NOMEM;
    } else {
      g->device_bytes += sizeof(int32_t) * (rows + 1);
      const int grid = (int)std::min<int64_t>((rows + 256) / 256, (int64_t)lib().sm_count * 8);
      if (g->out.off64) row_lengths<uint64_t><<<grid, 256, 0, st>>>((const uint64_t *)g->out.rowptr, g->out_degree, rows);
      else row_lengths<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)g->out.rowptr, g->out_degree, rows);
      if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("row_lengths failed"); rc = GDN_ERR_CUDA; }
    }
  }
  trace("graph_create: uploads finished");
  if (rc != GDN_OK) { gdn_graph_destroy(g); return rc; }
  g->prep_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count();
  *out = g;
  return GDN_OK;
}

// gen-1 callers (int offsets) behind a row partition: the gang's workers (oneshot.cu)
int graph_create_i32_part(int32_t m, int32_t nnz, const int32_t *out_rp, const int32_t *out_ci, const int32_t *in_rp, const int32_t *in_ci,
                          int64_t row_lo, int64_t row_hi, gdn_graph **g) {
  return graph_create_t<int32_t>(m, nnz, out_rp, out_ci, in_rp, in_ci, row_lo, row_hi, g);
}

}  // namespace gdn

using namespace gdn;

extern "C" {

int gdn_version(void) { return GDN_VERSION; }
const char *gdn_last_error(void) { return g_err; }

int gdn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gdn_init(int device) {
  Lib &l = lib();
  if (l.inited && l.device == device) return GDN_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libgdn_b200 has no CPU fallback (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    return GDN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return GDN_ERR_ARG; }
  cudaDeviceProp p;
  GDN_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    return GDN_ERR_NO_DEVICE;
  }
  if (l.inited) gdn_finalize();
  GDN_CUDA(cudaSetDevice(device));
  l.device = device;
  l.sm_count = p.multiProcessorCount;
  GDN_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
  GDN_CUDA(cudaStreamCreateWithFlags(&l.copy_stream, cudaStreamNonBlocking));
  GDN_CUDA(cudaEventCreate(&l.ev0));
  GDN_CUDA(cudaEventCreate(&l.ev1));
  for (int i = 0; i < 2 * Lib::kMaxKev; i++) GDN_CUDA(cudaEventCreate(&l.kev[i]));
  l.pinned_bytes = 4096;
  GDN_CUDA(cudaHostAlloc(&l.pinned, l.pinned_bytes, cudaHostAllocDefault));
  { const char *e = getenv("GDN_PR_EXACT"); if (e) l.pr_exact = atoi(e) != 0; }
  l.inited = true;
  return GDN_OK;
}

int gdn_finalize(void) {
  if (!gang_worker()) gang_stop();      // the workers of gdn_init_gpus finalize their own contexts
  Lib &l = lib();
  if (!l.inited) return GDN_OK;
  cudaStreamSynchronize(l.stream);
  pool_release();
  cudaStreamDestroy(l.stream);
  cudaStreamSynchronize(l.copy_stream);
  cudaStreamDestroy(l.copy_stream);
  cudaEventDestroy(l.ev0);
  cudaEventDestroy(l.ev1);
  for (cudaEvent_t e : l.pr_ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2 * Lib::kMaxKev; i++) cudaEventDestroy(l.kev[i]);
  cudaFreeHost(l.pinned);
  for (int i = 0; i < 2; i++) if (l.arena[i]) cudaFreeHost(l.arena[i]);
  l = Lib();
  return GDN_OK;
}

int gdn_graph_create(int64_t m, int64_t nnz, const uint64_t *out_rowptr, const int32_t *out_colidx,
                     const uint64_t *in_rowptr, const int32_t *in_colidx, int64_t row_lo, int64_t row_hi,
                     gdn_graph **g) {
  return graph_create_t<uint64_t>(m, nnz, out_rowptr, out_colidx, in_rowptr, in_colidx, row_lo, row_hi, g);
}

int gdn_graph_create_i32(int32_t m, int32_t nnz, const int32_t *out_row_offsets, const int32_t *out_column_indices,
                         const int32_t *in_row_offsets, const int32_t *in_column_indices, gdn_graph **g) {
  return graph_create_t<int32_t>(m, nnz, out_row_offsets, out_column_indices, in_row_offsets, in_column_indices, 0, m, g);
}

int gdn_graph_set_out_degree(gdn_graph *g, const int32_t *h_out_degree) {
  if (!g || !h_out_degree) { set_error("gdn_graph_set_out_degree: null argument"); return GDN_ERR_ARG; }
  const int64_t rows = g->row_hi - g->row_lo;
  if (!g->out_degree) {
    GDN_CUDA(cudaMalloc((void **)&g->out_degree, sizeof(int32_t) * (rows + 1)));
    g->device_bytes += sizeof(int32_t) * (rows + 1);
  }
  GDN_CUDA(cudaMemcpyAsync(g->out_degree, h_out_degree, sizeof(int32_t) * rows, cudaMemcpyHostToDevice, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  return GDN_OK;
}

int gdn_graph_destroy(gdn_graph *g) {
  if (!g) return GDN_OK;
  pull_peer_release(g);          // row partition: the peers unmap this graph's vectors before they are freed
  for (cudaEvent_t ev : g->col_ev) cudaEventDestroy(ev);
  if (g->symmetric) { free_csr(g->out); g->in = DevCsr(); }
  else { free_csr(g->out); free_csr(g->in); }
  cudaFree(g->contrib[0]); cudaFree(g->contrib[1]); cudaFree(g->out_degree); cudaFree(g->err_partial);
  cudaFree(g->spmv_col); cudaFree(g->spmv_x);
  { auto &x = g->spmv_exact; cudaFree(x.blk_base); cudaFree(x.vals); cudaFree(x.S); cudaFree(x.mx); cudaFree(x.Q); cudaFree(x.plan); }
  cudaFree(g->err_trace); cudaFree(g->pr_done); cudaFree(g->pr_work); cudaFree(g->abs_partial);
  cudaFree(g->visited); cudaFree(g->front); cudaFree(g->next); cudaFree(g->iso); cudaFree(g->queue[0]); cudaFree(g->queue[1]);
  cudaFree(g->heavy_queue); cudaFree(g->heavy_off); cudaFree(g->deg_class); cudaFree(g->col_bu); cudaFree(g->bu_head); cudaFree(g->bfs_reached);
  cudaFree(g->counters); cudaFree(g->xbuf); cudaFree(g->bfs_ctrl); cudaFree(g->parent_buf);
  if (g->bfs_ctrl_host) cudaFreeHost(g->bfs_ctrl_host);
  {
    gdn::PullLayout &L = g->pull;
    cudaFree(L.perm); cudaFree(L.newid); cudaFree(L.sdeg); cudaFree(L.sout); cudaFree(L.rowid); cudaFree(L.slice_ptr);
    cudaFree(L.sell); cudaFree(L.exact_vals); cudaFree(L.x_blk_base); cudaFree(L.x_mx); cudaFree(L.x_Q); cudaFree(L.x_S); cudaFree(L.x_plan);
    cudaFree(L.chunk_slice); cudaFree(L.heavy_slice); cudaFree(L.heavy_first); cudaFree(L.heavy_seg);
    cudaFree(L.partial); cudaFree(g->scores_sorted);
    gdn::band_free(L.band);
  }
  cudaGetLastError();
  delete g;
  return GDN_OK;
}

int gdn_graph_info(const gdn_graph *g, int64_t info[8]) {
  if (!g || !info) return GDN_ERR_ARG;
  info[0] = g->m; info[1] = (int64_t)g->in.nnz; info[2] = g->row_lo; info[3] = g->row_hi;
  info[4] = g->in.n_chunks; info[5] = g->in.n_heavy_segs; info[6] = (int64_t)g->device_bytes;
  info[7] = g->in.off64 ? 64 : 32;
  return GDN_OK;
}

int gdn_graph_pull_info(const gdn_graph *g, int64_t info[8]) {
  if (!g || !info) return GDN_ERR_ARG;
  const gdn::BandLayout &b = g->pull.band;
  info[0] = b.built ? (b.seg ? 2 : 1) : 0; info[1] = b.B; info[2] = b.band; info[3] = b.n_rows;
  info[4] = (int64_t)b.moved; info[5] = (int64_t)b.pairs; info[6] = b.n_items;
  info[7] = (int64_t)(b.built ? b.n_groups : g->pull.n_groups);
  return GDN_OK;
}

int gdn_graph_prep_ms(const gdn_graph *g, double ms[4]) {
  if (!g || !ms) return GDN_ERR_ARG;
  for (int i = 0; i < 4; i++) ms[i] = g->prep_ms[i];
  ms[2] = g->pull.band.build_ms;
  return GDN_OK;
}

int gdn_set_pr_exact_order(int on) {
  lib().pr_exact = on != 0;
  return GDN_OK;
}

int gdn_device_trim(void) {
  pool_release();
  return GDN_OK;
}

int gdn_bfs_resident(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent, gdn_stats *st) {
  if (!g || !d_depth) { set_error("gdn_bfs_resident: null argument"); return GDN_ERR_ARG; }
  if (st) memset(st, 0, sizeof(*st));
  return bfs_run(g, source, d_depth, d_parent, st);
}

int gdn_pagerank_resident(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st) {
  if (!g || !d_scores || max_iter < 0) { set_error("gdn_pagerank_resident: bad argument"); return GDN_ERR_ARG; }
  if (!g->has_in) { set_error("PageRank pull needs the in-CSR"); return GDN_ERR_GRAPH; }
  if (!g->symmetric && !g->out_degree) { set_error("PageRank on a pull-only graph needs gdn_graph_set_out_degree"); return GDN_ERR_GRAPH; }
  if (st) memset(st, 0, sizeof(*st));
  return pr_run(g, d_scores, damp, eps, max_iter, st);
}

int gdn_spmv_resident(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y, gdn_stats *st) {
  if (!g || !d_Ax || !d_x || !d_y) { set_error("gdn_spmv_resident: null argument"); return GDN_ERR_ARG; }
  if (!g->has_in) { set_error("SpMV needs the in-CSR"); return GDN_ERR_GRAPH; }
  if (st) memset(st, 0, sizeof(*st));
  return spmv_run(g, d_Ax, d_x, d_y, st);
}

int gdn_dev_alloc(size_t bytes, void **d_ptr) {
  GDN_CHECK(ensure_init());
  GDN_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return GDN_OK;
}
int gdn_dev_free(void *d_ptr) { GDN_CUDA(cudaFree(d_ptr)); return GDN_OK; }
int gdn_memcpy_h2d(void *d, const void *h, size_t bytes) {
  GDN_CHECK(ensure_init());
  GDN_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  return GDN_OK;
}
int gdn_memcpy_d2h(void *h, const void *d, size_t bytes) {
  GDN_CHECK(ensure_init());
  GDN_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  return GDN_OK;
}
int gdn_host_pin(void *h_ptr, size_t bytes) {
  GDN_CHECK(ensure_init());
  GDN_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable));      // (portable: every GPU of a gang copies from it)
  return GDN_OK;
}
int gdn_host_unpin(void *h_ptr) {
  GDN_CUDA(cudaHostUnregister(h_ptr));
  return GDN_OK;
}
int gdn_device_sync(void) {
  GDN_CHECK(ensure_init());
  GDN_CUDA(cudaDeviceSynchronize());
  return GDN_OK;
}

}  // extern "C"
