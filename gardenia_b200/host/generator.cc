// generator.cc -- see generator.hpp.
#include "generator.hpp"
#include <omp.h>
#include <algorithm>
#include <random>
#include <vector>

namespace gdn {

// generator.h:81-114.  Each 2^18-edge block reseeds mt19937 with kRandSeed+block
// so the stream is independent of the thread count; R-MAT quadrant
// probabilities A=.57 B=.19 C=.19, one float draw per bit of the id.
void make_rmat_el(int scale, int degree, EdgePair32 *el) {
  const float A = 0.57f, B = 0.19f, C = 0.19f;
  const int64_t num_vertices = int64_t(1) << scale;
  const int64_t num_edges = num_vertices * degree;
#pragma omp parallel
  {
    std::mt19937 rng;
    std::uniform_real_distribution<float> udist(0, 1.0f);
#pragma omp for schedule(dynamic, 4)
    for (int64_t block = 0; block < num_edges; block += kGenBlockSize) {
      rng.seed(kRandSeed + block / kGenBlockSize);
      const int64_t hi = std::min(block + kGenBlockSize, num_edges);
      for (int64_t e = block; e < hi; e++) {
        VertexId src = 0, dst = 0;
        for (int depth = 0; depth < scale; depth++) {
          float r = udist(rng);
          src <<= 1;
          dst <<= 1;
          if (r < A + B) {
            if (r > A) dst++;
          } else {
            src++;
            if (r > A + B + C) dst++;
          }
        }
        el[e].u = src;
        el[e].v = dst;
      }
    }
  }
  // PermuteIDs, generator.h:52-62: std::shuffle with mt19937(kRandSeed)
  std::vector<VertexId> perm(num_vertices);
#pragma omp parallel for
  for (int64_t n = 0; n < num_vertices; n++) perm[n] = (VertexId)n;
  std::mt19937 prng(kRandSeed);
  std::shuffle(perm.begin(), perm.end(), prng);
#pragma omp parallel for
  for (int64_t e = 0; e < num_edges; e++) {
    el[e].u = perm[el[e].u];
    el[e].v = perm[el[e].v];
  }
}

// generator.h:64-79.  The reference builds Edge(udist(rng), udist(rng)), whose
// argument order is unspecified; synthetic graphs are always symmetrized, so
// the resulting CSR does not depend on it.
void make_uniform_el(int scale, int degree, EdgePair32 *el) {
  const int64_t num_vertices = int64_t(1) << scale;
  const int64_t num_edges = num_vertices * degree;
#pragma omp parallel
  {
    std::mt19937 rng;
    std::uniform_int_distribution<VertexId> udist(0, (VertexId)(num_vertices - 1));
#pragma omp for schedule(dynamic, 4)
    for (int64_t block = 0; block < num_edges; block += kGenBlockSize) {
      rng.seed(kRandSeed + block / kGenBlockSize);
      const int64_t hi = std::min(block + kGenBlockSize, num_edges);
      for (int64_t e = block; e < hi; e++) {
        VertexId a = udist(rng);
        VertexId b = udist(rng);
        el[e].u = a;
        el[e].v = b;
      }
    }
  }
}

VertexId build_symmetric_csr(const EdgePair32 *el, int64_t n_edges, int64_t &m_out,
                             uint64_t *&rowptr, VertexId *&col, uint64_t &nnz) {
  // FindMaxVertexID, builder.h:66-75
  VertexId max_seen = 0;
#pragma omp parallel for reduction(max : max_seen)
  for (int64_t e = 0; e < n_edges; e++) max_seen = std::max(max_seen, std::max(el[e].u, el[e].v));
  const int64_t m = (int64_t)max_seen + 1;
  m_out = m;
  // CountDegrees (symmetrize: both endpoints), builder.h:76-87
  std::vector<uint32_t> deg(m, 0);
#pragma omp parallel for
  for (int64_t e = 0; e < n_edges; e++) {
#pragma omp atomic
    deg[el[e].u]++;
#pragma omp atomic
    deg[el[e].v]++;
  }
  std::vector<uint64_t> off(m + 1);
  {
    uint64_t s = 0;
    for (int64_t i = 0; i < m; i++) { off[i] = s; s += deg[i]; }
    off[m] = s;
  }
  std::vector<uint64_t> cur(off.begin(), off.end() - 1);
  VertexId *raw = new VertexId[std::max<uint64_t>(off[m], 1)];
  // MakeCSR scatter, builder.h:220-236
#pragma omp parallel for
  for (int64_t e = 0; e < n_edges; e++) {
    uint64_t p;
#pragma omp atomic capture
    p = cur[el[e].u]++;
    raw[p] = el[e].v;
#pragma omp atomic capture
    p = cur[el[e].v]++;
    raw[p] = el[e].u;
  }
  // SquishCSR, builder.h:152-183: sort, unique, remove self
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < m; v++) {
    VertexId *b = raw + off[v], *e = raw + off[v + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    e = std::remove(b, e, (VertexId)v);
    deg[v] = (uint32_t)(e - b);
  }
  rowptr = new uint64_t[m + 1];
  {
    uint64_t s = 0;
    for (int64_t i = 0; i < m; i++) { rowptr[i] = s; s += deg[i]; }
    rowptr[m] = s;
  }
  nnz = rowptr[m];
  col = new VertexId[std::max<uint64_t>(nnz, 1)];
  VertexId maxdeg = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : maxdeg)
  for (int64_t v = 0; v < m; v++) {
    std::copy(raw + off[v], raw + off[v] + deg[v], col + rowptr[v]);
    maxdeg = std::max(maxdeg, (VertexId)deg[v]);
  }
  delete[] raw;
  return maxdeg;
}

void generate_graph(Graph &g, bool uniform, int scale, int degree) {
  const int64_t num_edges = (int64_t(1) << scale) * degree;
  EdgePair32 *el = new EdgePair32[num_edges];
  if (uniform) make_uniform_el(scale, degree, el);
  else make_rmat_el(scale, degree, el);
  int64_t m;
  uint64_t *rowptr, nnz;
  VertexId *col;
  VertexId maxdeg = build_symmetric_csr(el, num_edges, m, rowptr, col, nnz);
  delete[] el;
  g.adopt_symmetric((VertexId)m, nnz, rowptr, col, maxdeg);
}

}  // namespace gdn
