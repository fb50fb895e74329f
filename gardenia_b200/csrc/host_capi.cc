// host_capi.cc -- C-ABI over the host-side graph layer (readers + generator).
// Pure host code: works without a GPU.  See include/gdn_b200.h.
#include <omp.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include "../../include/gdn_b200.h"
#include "../host/csr_graph.hpp"
#include "../host/generator.hpp"
#include "../host/graph_io.hpp"

namespace gdn {
void set_error(const char *fmt, ...);
// build.cu: the CSR builder on the GPU (same result as build_symmetric_csr)
int gpu_build_symmetric_csr(const void *el_host, int64_t n_edges, int64_t *m_out, uint64_t **rowptr_out, int32_t **col_out,
                            uint64_t *nnz_out, int32_t *maxdeg_out, double ms[4]);
}

struct gdn_host_graph {
  gdn::Graph g;                 // gen-2 loader / generator
  gdn::Csr1 c1;                 // gen-1 loader
  bool gen1 = false;
  uint64_t *rowptr64 = nullptr; // gen-1: widened offsets
  ~gdn_host_graph() {
    free(c1.row_offsets); free(c1.column_indices); free(c1.weight); free(c1.degree);
    delete[] rowptr64;
  }
};

extern "C" {

int gdn_read_graph(const char *prefix, const char *filetype, int symmetrize, int need_reverse,
                   gdn_host_graph **out) {
  if (!prefix || !filetype || !out) { gdn::set_error("gdn_read_graph: null argument"); return GDN_ERR_ARG; }
  gdn_host_graph *hg = new gdn_host_graph();
  std::string ft = filetype;
  {
    const std::string path = prefix;                       // "auto" on a .sg path: the serialized-graph reader
    if (ft == "auto" && path.size() > 3 && path.compare(path.size() - 3, 3, ".sg") == 0) ft = "sg";
  }
  if (ft == "auto") {
    gdn::ReadOpts o;
    o.symmetrize = symmetrize != 0;
    o.verbose = false;
    int rc = gdn::read_graph_file(prefix, hg->c1, o);
    if (rc != 0) {
      delete hg;
      gdn::set_error("cannot read %s (gen-1 reader error %d)", prefix, rc);
      return rc == -1 ? GDN_ERR_IO : GDN_ERR_GRAPH;
    }
    hg->gen1 = true;
    hg->rowptr64 = new uint64_t[hg->c1.m + 1];
    for (int i = 0; i <= hg->c1.m; i++) hg->rowptr64[i] = (uint64_t)hg->c1.row_offsets[i];
  } else {
    int rc = hg->g.load(prefix, ft, symmetrize != 0, need_reverse != 0, false);
    if (rc != gdn::kLoadOk) {
      delete hg;
      gdn::set_error("cannot load %s as %s (error %d)", prefix, filetype, rc);
      return rc == gdn::kLoadNoFile ? GDN_ERR_IO : (rc == gdn::kLoadBadType ? GDN_ERR_ARG : GDN_ERR_GRAPH);
    }
  }
  *out = hg;
  return GDN_OK;
}

int gdn_generate(char kind, int scale, int degree, gdn_host_graph **out) {
  if (!out || (kind != 'g' && kind != 'u') || scale < 1 || scale > 30 || degree < 1) {
    gdn::set_error("gdn_generate: bad argument");
    return GDN_ERR_ARG;
  }
  gdn_host_graph *hg = new gdn_host_graph();
  gdn::generate_graph(hg->g, kind == 'u', scale, degree);
  *out = hg;
  return GDN_OK;
}

// Edge list -> symmetric, sorted, duplicate-free, self-loop-free CSR, built on the GPU (csrc/build.cu).
int gdn_build_csr_gpu(int64_t n_edges, const int32_t *pairs, gdn_host_graph **out, double *ms /* nullable [4] */) {
  if (!out || n_edges < 0 || (n_edges > 0 && !pairs)) { gdn::set_error("gdn_build_csr_gpu: bad argument"); return GDN_ERR_ARG; }
  int64_t m = 0;
  uint64_t *rowptr = nullptr, nnz = 0;
  int32_t *col = nullptr, maxdeg = 0;
  const int rc = gdn::gpu_build_symmetric_csr(pairs, n_edges, &m, &rowptr, &col, &nnz, &maxdeg, ms);
  if (rc != GDN_OK) return rc;
  gdn_host_graph *hg = new gdn_host_graph();
  hg->g.adopt_symmetric((VertexId)m, nnz, rowptr, col, (VertexId)maxdeg);
  *out = hg;
  return GDN_OK;
}

// gdn_generate with the builder on the GPU: the edge streams are drawn on the host (they are defined by libstdc++'s
// mt19937 / distributions, include/generator.h:64-114), the CSR is built by csrc/build.cu.  ms (nullable, 5 entries):
// edge generation, upload + keys, sort, unique + offsets, download.
int gdn_generate_gpu(char kind, int scale, int degree, gdn_host_graph **out, double *ms) {
  if (!out || (kind != 'g' && kind != 'u') || scale < 1 || scale > 30 || degree < 1) {
    gdn::set_error("gdn_generate_gpu: bad argument");
    return GDN_ERR_ARG;
  }
  const int64_t n_edges = (int64_t(1) << scale) * degree;
  gdn::EdgePair32 *el = new gdn::EdgePair32[n_edges];
  const double t0 = omp_get_wtime();
  if (kind == 'u') gdn::make_uniform_el(scale, degree, el);
  else gdn::make_rmat_el(scale, degree, el);
  if (ms) ms[0] = (omp_get_wtime() - t0) * 1e3;
  const int rc = gdn_build_csr_gpu(n_edges, reinterpret_cast<const int32_t *>(el), out, ms ? ms + 1 : nullptr);
  delete[] el;
  return rc;
}

int gdn_host_graph_free(gdn_host_graph *hg) { delete hg; return GDN_OK; }
int64_t gdn_host_graph_m(const gdn_host_graph *hg) { return hg->gen1 ? hg->c1.m : hg->g.V(); }
int64_t gdn_host_graph_nnz(const gdn_host_graph *hg) { return hg->gen1 ? hg->c1.nnz : (int64_t)hg->g.E(); }
int gdn_host_graph_symmetric(const gdn_host_graph *hg) {
  return hg->gen1 ? 0 : (hg->g.has_reverse_graph() && hg->g.in_rowptr() == hg->g.out_rowptr());
}
int gdn_host_graph_has_reverse(const gdn_host_graph *hg) { return hg->gen1 ? 0 : hg->g.has_reverse_graph(); }
const uint64_t *gdn_host_graph_out_rowptr(const gdn_host_graph *hg) { return hg->gen1 ? hg->rowptr64 : hg->g.out_rowptr(); }
const int32_t *gdn_host_graph_out_colidx(const gdn_host_graph *hg) { return hg->gen1 ? hg->c1.column_indices : hg->g.out_colidx(); }
const uint64_t *gdn_host_graph_in_rowptr(const gdn_host_graph *hg) {
  return (hg->gen1 || !hg->g.has_reverse_graph()) ? nullptr : hg->g.in_rowptr();
}
const int32_t *gdn_host_graph_in_colidx(const gdn_host_graph *hg) {
  return (hg->gen1 || !hg->g.has_reverse_graph()) ? nullptr : hg->g.in_colidx();
}
const int32_t *gdn_host_graph_weights(const gdn_host_graph *hg) { return hg->gen1 ? hg->c1.weight : nullptr; }

int gdn_host_graph_write_sg(const gdn_host_graph *hg, const char *path, int offset_bytes) {
  if (!hg || !path || hg->gen1) { gdn::set_error("gdn_host_graph_write_sg: bad argument"); return GDN_ERR_ARG; }
  if (hg->g.write_sg(path, offset_bytes) != 0) {
    gdn::set_error("cannot write %s (offset width %d; a directed graph needs its reverse CSR)", path, offset_bytes);
    return GDN_ERR_IO;
  }
  return GDN_OK;
}

int gdn_host_graph_write_bin(const gdn_host_graph *hg, const char *prefix) {
  if (!hg || !prefix || hg->gen1) { gdn::set_error("gdn_host_graph_write_bin: bad argument"); return GDN_ERR_ARG; }
  return hg->g.write_bin(prefix) == 0 ? GDN_OK : GDN_ERR_IO;
}

int gdn_set_host_threads(int n) {
  if (n < 1) return GDN_ERR_ARG;
  omp_set_num_threads(n);
  return GDN_OK;
}

int gdn_fill_uniform(uint32_t seed, int64_t n, float *out) {
  if (!out || n < 0) return GDN_ERR_ARG;
  std::mt19937 rng(seed);
  for (int64_t i = 0; i < n; i++) out[i] = (float)(rng() >> 8) * (1.0f / 16777216.0f);
  return GDN_OK;
}

int gdn_pick_sources(const gdn_host_graph *hg, int n, int32_t *sources) {
  if (!hg || !sources || n < 0) return GDN_ERR_ARG;
  const int64_t m = gdn_host_graph_m(hg);
  const uint64_t *rp = gdn_host_graph_out_rowptr(hg);
  if (rp[m] == 0) { gdn::set_error("graph has no edges"); return GDN_ERR_GRAPH; }
  std::mt19937 rng(gdn::kRandSeed);
  std::uniform_int_distribution<int32_t> ud(0, (int32_t)(m - 1));
  for (int i = 0; i < n; i++) {
    int32_t s;
    do { s = ud(rng); } while (rp[s + 1] == rp[s]);
    sources[i] = s;
  }
  return GDN_OK;
}

}  // extern "C"
