// oneshot.cu -- the reference-facing one-shot entry points: HOST pointers in,
// HOST results out; upload, solve, download and free inside the call, exactly
// the ownership contract of the reference's CUDA solvers
// (src/pr/base.cu:83-101,133-139; src/bfs/linear_base.cu:34-89;
// src/spmv/base.cu:28-76).  These are what BFSSolver / PRSolver / SpmvSolver
// shims bind.
#include "common.cuh"
#include <algorithm>
#include <chrono>
#include <cstring>

using namespace gdn;
namespace gdn {
int64_t partition_width(int64_t m, int nparts);
int graph_create_i32_part(int32_t m, int32_t nnz, const int32_t *out_rp, const int32_t *out_ci, const int32_t *in_rp, const int32_t *in_ci,
                          int64_t row_lo, int64_t row_hi, gdn_graph **g);
}

namespace {
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { GDN_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); return GDN_OK; }
};
struct GraphGuard {
  gdn_graph *g = nullptr;
  ~GraphGuard() { gdn_graph_destroy(g); }
};
int h2d(void *d, const void *h, size_t bytes) {
  GDN_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, lib().stream));
  return GDN_OK;
}
int d2h(void *h, const void *d, size_t bytes) {
  GDN_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  return GDN_OK;
}

template <typename OffT>
int create(int64_t m, int64_t nnz, const OffT *orp, const int32_t *oci, const OffT *irp, const int32_t *ici,
           gdn_graph **g);
template <>
int create<uint64_t>(int64_t m, int64_t nnz, const uint64_t *orp, const int32_t *oci, const uint64_t *irp,
                     const int32_t *ici, gdn_graph **g) {
  return gdn_graph_create(m, nnz, orp, oci, irp, ici, 0, m, g);
}
template <>
int create<int32_t>(int64_t m, int64_t nnz, const int32_t *orp, const int32_t *oci, const int32_t *irp,
                    const int32_t *ici, gdn_graph **g) {
  return gdn_graph_create_i32((int32_t)m, (int32_t)nnz, orp, oci, irp, ici, g);
}

template <typename OffT>
int bfs_oneshot(int64_t m, int64_t nnz, const OffT *orp, const int32_t *oci, const OffT *irp, const int32_t *ici,
                int32_t source, int32_t *depth_out, int32_t *parent_out, gdn_stats *st) {
  if (!orp || !oci || !depth_out) { set_error("gdn_bfs: null argument"); return GDN_ERR_ARG; }
  GDN_CHECK(ensure_init());
  PoolScope arena;                 // declared before every buffer of the call: they are parked, not freed, on the way out
  const double t0 = now_ms();
  GraphGuard gg;
  GDN_CHECK(create<OffT>(m, nnz, orp, oci, irp, ici, &gg.g));
  gg.g->one_shot = true;
  DevBuf depth, parent;
  GDN_CHECK(depth.alloc(sizeof(int32_t) * m));
  if (parent_out) GDN_CHECK(parent.alloc(sizeof(int32_t) * m));
  const double t1 = now_ms();
  gdn_stats local;
  gdn_stats *s = st ? st : &local;
  GDN_CHECK(gdn_bfs_resident(gg.g, source, (int32_t *)depth.p, (int32_t *)parent.p, s));
  const double t2 = now_ms();
  GDN_CHECK(d2h(depth_out, depth.p, sizeof(int32_t) * m));
  if (parent_out) GDN_CHECK(d2h(parent_out, parent.p, sizeof(int32_t) * m));
  const double t3 = now_ms();
  s->h2d_ms = t1 - t0; s->d2h_ms = t3 - t2;
  s->h2d_bytes = (int64_t)(sizeof(OffT) * (m + 1) + sizeof(int32_t) * nnz) * ((irp && irp != orp) ? 2 : 1);
  s->d2h_bytes = (int64_t)sizeof(int32_t) * m * (parent_out ? 2 : 1);
  return GDN_OK;
}

// ---- PageRank on the gang (gdn_init_gpus): every worker uploads ITS rows of the caller's CSR over its own PCIe link,
// solves its partition with the others (peer-mapped exchange, comm.cu) and writes its slice of the scores back.
template <typename OffT>
struct PrJob {
  int64_t m, nnz;
  const OffT *irp;
  const int32_t *ici, *out_degree;
  float *scores;
  float damp;
  double eps;
  int max_iter;
  gdn_stats st[8];
};
template <typename OffT>
int create_part(int64_t m, int64_t nnz, const OffT *irp, const int32_t *ici, int64_t lo, int64_t hi, gdn_graph **g);
template <>
int create_part<uint64_t>(int64_t m, int64_t nnz, const uint64_t *irp, const int32_t *ici, int64_t lo, int64_t hi, gdn_graph **g) {
  return gdn_graph_create(m, nnz, nullptr, nullptr, irp, ici, lo, hi, g);
}
template <>
int create_part<int32_t>(int64_t m, int64_t nnz, const int32_t *irp, const int32_t *ici, int64_t lo, int64_t hi, gdn_graph **g) {
  return graph_create_i32_part((int32_t)m, (int32_t)nnz, nullptr, nullptr, irp, ici, lo, hi, g);
}
template <typename OffT>
int pr_oneshot_rank(int rank, void *arg) {
  PrJob<OffT> &J = *(PrJob<OffT> *)arg;
  PoolScope arena;
  const int64_t W = partition_width(J.m, gang_size());
  const int64_t lo = std::min<int64_t>((int64_t)rank * W, J.m), hi = std::min<int64_t>(lo + W, J.m), rows = hi - lo;
  DevBuf sc, od;
  GDN_CHECK(sc.alloc(sizeof(float) * std::max<int64_t>(rows, 1)));
  GDN_CHECK(od.alloc(sizeof(int32_t) * (std::max<int64_t>(rows, 1) + 1)));
  GDN_CHECK(h2d(sc.p, J.scores + lo, sizeof(float) * rows));
  GDN_CHECK(h2d(od.p, J.out_degree + lo, sizeof(int32_t) * rows));
  GraphGuard gg;
  GDN_CHECK(create_part<OffT>(J.m, J.nnz, J.irp, J.ici, lo, hi, &gg.g));
  gg.g->one_shot = true;
  gg.g->out_degree = (int32_t *)od.p;      // the graph owns it from here
  od.p = nullptr;
  GDN_CHECK(gdn_pagerank_resident(gg.g, (float *)sc.p, J.damp, J.eps, J.max_iter, &J.st[rank]));
  GDN_CHECK(d2h(J.scores + lo, sc.p, sizeof(float) * rows));
  return GDN_OK;
}

template <typename OffT>
int pr_oneshot(int64_t m, int64_t nnz, const OffT *irp, const int32_t *ici, const int32_t *out_degree,
               float *scores, float damp, double eps, int max_iter, gdn_stats *st) {
  if (!irp || !ici || !out_degree || !scores) { set_error("gdn_pagerank_pull: null argument"); return GDN_ERR_ARG; }
  GDN_CHECK(ensure_init());
  if (gang_size() > 1 && !gang_worker() && m >= (int64_t)gang_size() * 65536) {
    // (small graphs stay on one GPU: a partition must be worth a GPU, and every rank needs rows of its own)
    PrJob<OffT> *J = new PrJob<OffT>();
    J->m = m; J->nnz = nnz; J->irp = irp; J->ici = ici; J->out_degree = out_degree; J->scores = scores;
    J->damp = damp; J->eps = eps; J->max_iter = max_iter;
    const double t0 = now_ms();
    const int rc = gang_run(pr_oneshot_rank<OffT>, J);
    const double t1 = now_ms();
    if (rc == GDN_OK && st) {
      *st = J->st[0];
      for (int r = 1; r < gang_size(); r++) {
        st->solve_ms = std::max(st->solve_ms, J->st[r].solve_ms);
        st->kernel_launches += J->st[r].kernel_launches;
      }
      st->h2d_ms = (t1 - t0) - st->solve_ms; st->d2h_ms = 0;      // upload + layout + download of all GPUs, overlapped
      st->h2d_bytes = (int64_t)(sizeof(OffT) * (m + 1) + sizeof(int32_t) * nnz + 8 * m);
      st->d2h_bytes = (int64_t)sizeof(float) * m;
    }
    delete J;
    return rc;
  }
  struct TraceAt { const char *what; ~TraceAt() { trace(what); } };
  TraceAt t_exit{"pr_oneshot: arena scope closed"};
  PoolScope arena;                 // declared before every buffer of the call: they are parked, not freed, on the way out
  TraceAt t_freed{"pr_oneshot: buffers released"};
  const double t0 = now_ms();
  // the two small inputs go first (queued, not waited for); then the CSR, whose column array crosses PCIe in
  // pieces with the PageRank layout built behind them (Lib::stream_fill, graph.cu / pull.cu)
  DevBuf sc, od;
  GDN_CHECK(sc.alloc(sizeof(float) * m));
  GDN_CHECK(od.alloc(sizeof(int32_t) * (m + 1)));
  GDN_CHECK(h2d(sc.p, scores, sizeof(float) * m));
  GDN_CHECK(h2d(od.p, out_degree, sizeof(int32_t) * m));
  GraphGuard gg;
  lib().stream_fill = true;
  const int rc = create<OffT>(m, nnz, nullptr, nullptr, irp, ici, &gg.g);
  lib().stream_fill = false;
  GDN_CHECK(rc);
  gg.g->one_shot = true;
  gg.g->out_degree = (int32_t *)od.p;      // the graph owns it from here (gdn_graph_destroy frees it)
  od.p = nullptr;
  gg.g->device_bytes += sizeof(int32_t) * (m + 1);
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  const double t1 = now_ms();
  gdn_stats local;
  gdn_stats *s = st ? st : &local;
  trace("pr_oneshot: inputs resident");
  GDN_CHECK(gdn_pagerank_resident(gg.g, (float *)sc.p, damp, eps, max_iter, s));
  trace("pr_oneshot: solved");
  const double t2 = now_ms();
  GDN_CHECK(d2h(scores, sc.p, sizeof(float) * m));
  const double t3 = now_ms();
  trace("pr_oneshot: downloaded");
  s->h2d_ms = t1 - t0; s->d2h_ms = t3 - t2;
  s->h2d_bytes = (int64_t)(sizeof(OffT) * (m + 1) + sizeof(int32_t) * nnz + 8 * m);
  s->d2h_bytes = (int64_t)sizeof(float) * m;
  return GDN_OK;
}

template <typename OffT>
int spmv_oneshot(int64_t m, int64_t nnz, const OffT *Ap, const int32_t *Aj, const float *Ax, const float *x,
                 float *y, gdn_stats *st) {
  if (!Ap || !Aj || !Ax || !x || !y) { set_error("gdn_spmv_csr: null argument"); return GDN_ERR_ARG; }
  GDN_CHECK(ensure_init());
  PoolScope arena;                 // declared before every buffer of the call: they are parked, not freed, on the way out
  const double t0 = now_ms();
  GraphGuard gg;
  GDN_CHECK(create<OffT>(m, nnz, nullptr, nullptr, Ap, Aj, &gg.g));
  gg.g->one_shot = true;
  DevBuf dAx, dx, dy;
  GDN_CHECK(dAx.alloc(sizeof(float) * nnz + 256));
  GDN_CHECK(dx.alloc(sizeof(float) * m));
  GDN_CHECK(dy.alloc(sizeof(float) * m));
  GDN_CHECK(h2d(dAx.p, Ax, sizeof(float) * nnz));
  GDN_CHECK(h2d(dx.p, x, sizeof(float) * m));
  GDN_CHECK(h2d(dy.p, y, sizeof(float) * m));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  const double t1 = now_ms();
  gdn_stats local;
  gdn_stats *s = st ? st : &local;
  GDN_CHECK(gdn_spmv_resident(gg.g, (const float *)dAx.p, (const float *)dx.p, (float *)dy.p, s));
  const double t2 = now_ms();
  GDN_CHECK(d2h(y, dy.p, sizeof(float) * m));
  const double t3 = now_ms();
  s->h2d_ms = t1 - t0; s->d2h_ms = t3 - t2;
  s->h2d_bytes = (int64_t)(sizeof(OffT) * (m + 1) + 8 * nnz + 8 * m);
  s->d2h_bytes = (int64_t)sizeof(float) * m;
  return GDN_OK;
}
}  // namespace

extern "C" {

int gdn_bfs(int64_t m, int64_t nnz, const uint64_t *out_rowptr, const int32_t *out_colidx,
            const uint64_t *in_rowptr, const int32_t *in_colidx, int32_t source, int32_t *depth_out,
            int32_t *parent_out, gdn_stats *st) {
  return bfs_oneshot<uint64_t>(m, nnz, out_rowptr, out_colidx, in_rowptr, in_colidx, source, depth_out, parent_out, st);
}
int gdn_bfs_i32(int32_t m, int32_t nnz, const int32_t *out_row_offsets, const int32_t *out_column_indices,
                const int32_t *in_row_offsets, const int32_t *in_column_indices, int32_t source,
                int32_t *depth_out, int32_t *parent_out, gdn_stats *st) {
  return bfs_oneshot<int32_t>(m, nnz, out_row_offsets, out_column_indices, in_row_offsets, in_column_indices, source,
                              depth_out, parent_out, st);
}
int gdn_pagerank_pull(int64_t m, int64_t nnz, const uint64_t *in_rowptr, const int32_t *in_colidx,
                      const int32_t *out_degree, float *scores_inout, float damp, double eps, int max_iter,
                      gdn_stats *st) {
  return pr_oneshot<uint64_t>(m, nnz, in_rowptr, in_colidx, out_degree, scores_inout, damp, eps, max_iter, st);
}
int gdn_pagerank_pull_i32(int32_t m, int32_t nnz, const int32_t *in_row_offsets, const int32_t *in_column_indices,
                          const int32_t *out_degree, float *scores_inout, float damp, double eps, int max_iter,
                          gdn_stats *st) {
  return pr_oneshot<int32_t>(m, nnz, in_row_offsets, in_column_indices, out_degree, scores_inout, damp, eps, max_iter, st);
}
int gdn_spmv_csr(int64_t m, int64_t nnz, const uint64_t *Ap, const int32_t *Aj, const float *Ax, const float *x,
                 float *y_inout, gdn_stats *st) {
  return spmv_oneshot<uint64_t>(m, nnz, Ap, Aj, Ax, x, y_inout, st);
}
int gdn_spmv_csr_i32(int32_t m, int32_t nnz, const int32_t *Ap, const int32_t *Aj, const float *Ax, const float *x,
                     float *y_inout, gdn_stats *st) {
  return spmv_oneshot<int32_t>(m, nnz, Ap, Aj, Ax, x, y_inout, st);
}

}  // extern "C"
