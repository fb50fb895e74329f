// graph_io.cc -- see graph_io.hpp.
#include "graph_io.hpp"
#include <sys/time.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace gdn {
namespace {

struct WEdge { int src, dst; WeightT wt; };

struct Text {
  std::string buf;
  const char *p = nullptr, *end = nullptr;
  bool open(const char *path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    buf.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    p = buf.data();
    end = p + buf.size();
    return true;
  }
  // returns [b,e) of the next line (without '\n'); false at EOF
  bool line(const char *&b, const char *&e) {
    if (p >= end) return false;
    b = p;
    while (p < end && *p != '\n') p++;
    e = p;
    if (p < end) p++;
    if (e > b && e[-1] == '\r') e--;
    return true;
  }
};

bool next_int(const char *&p, const char *e, long &v) {
  while (p < e && (*p == ' ' || *p == '\t')) p++;
  const char *q = p;
  bool neg = false;
  if (q < e && (*q == '-' || *q == '+')) { neg = *q == '-'; q++; }
  if (q >= e || *q < '0' || *q > '9') return false;
  long x = 0;
  while (q < e && *q >= '0' && *q <= '9') { x = x * 10 + (*q - '0'); q++; }
  // swallow a fractional part / exponent so "3.5" parses as 3 like sscanf("%d") would stop
  p = q;
  v = neg ? -x : x;
  return true;
}
void skip_token(const char *&p, const char *e) {
  while (p < e && *p != ' ' && *p != '\t') p++;
}

// fill_data (graph_io.h:25-143): stable sort by (src,dst), drop self loops and
// duplicates, build int offsets.
int fill(int m, int nnz_hdr, std::vector<WEdge> &edges, Csr1 &out, const ReadOpts &o) {
  for (const WEdge &e : edges)
    if (e.src < 0 || e.src >= m || e.dst < 0 || e.dst >= m) return -3;
  if (o.sorted) {
    if (o.verbose) printf("Sorting the neighbor lists...");
    std::stable_sort(edges.begin(), edges.end(), [](const WEdge &a, const WEdge &b) {
      return a.src != b.src ? a.src < b.src : a.dst < b.dst;
    });
    if (o.verbose) printf(" Done\n");
  } else {
    std::stable_sort(edges.begin(), edges.end(), [](const WEdge &a, const WEdge &b) { return a.src < b.src; });
  }
  int num_selfloops = 0, num_redundents = 0;
  size_t w = 0;
  if (o.remove_selfloops) {
    if (o.verbose) printf("Removing self loops...");
    for (size_t i = 0; i < edges.size(); i++) {
      if (edges[i].src == edges[i].dst) { num_selfloops++; continue; }
      edges[w++] = edges[i];
    }
    edges.resize(w);
    if (o.verbose) printf(" %d selfloops are removed\n", num_selfloops);
  }
  if (o.remove_redundents) {
    if (o.verbose) printf("Removing redundent edges...");
    w = 0;
    for (size_t i = 0; i < edges.size(); i++) {
      if (w > 0 && edges[w - 1].src == edges[i].src && edges[w - 1].dst == edges[i].dst) { num_redundents++; continue; }
      edges[w++] = edges[i];
    }
    edges.resize(w);
    if (o.verbose) printf(" %d redundent edges are removed\n", num_redundents);
  }
  int count = (int)edges.size();
  out.m = m;
  out.row_offsets = (IndexT *)malloc((size_t)(m + 1) * sizeof(IndexT));
  out.column_indices = (IndexT *)malloc((size_t)std::max(count, 1) * sizeof(IndexT));
  out.weight = (WeightT *)malloc((size_t)std::max(count, 1) * sizeof(WeightT));
  out.degree = (int *)malloc((size_t)std::max(m, 1) * sizeof(int));
  std::fill(out.row_offsets, out.row_offsets + m + 1, 0);
  for (const WEdge &e : edges) out.row_offsets[e.src + 1]++;
  for (int i = 0; i < m; i++) out.row_offsets[i + 1] += out.row_offsets[i];
  for (int i = 0; i < count; i++) { out.column_indices[i] = edges[i].dst; out.weight[i] = edges[i].wt; }
  for (int i = 0; i < m; i++) out.degree[i] = out.row_offsets[i + 1] - out.row_offsets[i];
  if (!o.symmetrize && count + num_selfloops + num_redundents != nnz_hdr && o.verbose)
    printf("Error reading graph, number of edges in edge list %d != %d\n", count, nnz_hdr);
  out.nnz = count;
  if (o.verbose) printf("num_vertices %d num_edges %d\n", m, count);
  return 0;
}

// shared by mtx / el: "src dst [wt]" lines, 1-based (graph_io.h:305-333)
void add_edge(std::vector<WEdge> &edges, int src, int dst, WeightT wt, bool symmetrize, bool transpose) {
  if (symmetrize && src != dst) { edges.push_back({dst, src, wt}); transpose = false; }
  if (!transpose) edges.push_back({src, dst, wt});
  else edges.push_back({dst, src, wt});
}

int read_triplets(Text &t, int nnz, std::vector<WEdge> &edges, const ReadOpts &o) {
  const char *b, *e;
  WeightT wt = 1;
  for (int i = 0; i < nnz; i++) {
    if (!t.line(b, e)) break;
    long s, d, w;
    const char *q = b;
    if (!next_int(q, e, s) || !next_int(q, e, d)) continue;
    if (next_int(q, e, w)) wt = (WeightT)w; else wt = 1;
    if (wt < 0) wt = -wt;
    add_edge(edges, (int)s - 1, (int)d - 1, wt, o.symmetrize, o.transpose);
  }
  return 0;
}

}  // namespace

int mtx2csr(const char *path, Csr1 &out, const ReadOpts &o) {
  if (o.verbose) printf("Reading (.mtx) input file %s\n", path);
  Text t;
  if (!t.open(path)) return -1;
  const char *b, *e;
  do { if (!t.line(b, e)) return -2; } while (b < e && *b == '%');
  long m, n, nnz;
  const char *q = b;
  if (!next_int(q, e, m) || !next_int(q, e, n) || !next_int(q, e, nnz)) return -2;
  if (m != n && o.verbose) printf("Warning, m(%ld) != n(%ld)\n", m, n);
  if (o.verbose) printf("Before cleaning, the original num_vertices %ld num_edges %ld\n", m, nnz);
  std::vector<WEdge> edges;
  edges.reserve((size_t)nnz * (o.symmetrize ? 2 : 1));
  read_triplets(t, (int)nnz, edges, o);
  out.n = (int)n;
  return fill((int)m, (int)nnz, edges, out, o);
}

int el2csr(const char *path, Csr1 &out, const ReadOpts &o) {
  if (o.verbose) printf("Reading edgelist (.el) input file %s\n", path);
  Text t;
  if (!t.open(path)) return -1;
  const char *b, *e;
  if (!t.line(b, e)) return -2;
  long m, nnz;
  const char *q = b;
  if (!next_int(q, e, m) || !next_int(q, e, nnz)) return -2;
  if (o.verbose) printf("Before cleaning, the original num_vertices %ld num_edges %ld\n", m, nnz);
  std::vector<WEdge> edges;
  read_triplets(t, (int)nnz, edges, o);
  out.n = (int)m;
  return fill((int)m, (int)nnz, edges, out, o);
}

int graph2csr(const char *path, Csr1 &out, const ReadOpts &o) {
  if (o.verbose) printf("Reading .graph input file %s\n", path);
  Text t;
  if (!t.open(path)) return -1;
  const char *b, *e;
  if (!t.line(b, e)) return -2;
  long m, nnz, fmt = 0;
  const char *q = b;
  if (!next_int(q, e, m) || !next_int(q, e, nnz)) return -2;
  next_int(q, e, fmt);
  const bool edge_weights = (fmt % 10) == 1;
  if (o.verbose) printf("Before cleaning, the original num_vertices %ld num_edges %ld\n", m, nnz);
  std::vector<WEdge> edges;
  for (int src = 0; src < (int)m; src++) {               // graph_io.h:260-282: one line per vertex
    if (!t.line(b, e)) break;
    q = b;
    long d, w = 1;
    while (next_int(q, e, d)) {
      if (edge_weights && !next_int(q, e, w)) w = 1;
      // the file already holds both directions; symmetrize only clears transpose (:267-272)
      bool transpose = o.transpose && !(o.symmetrize && src != (int)d - 1);
      if (!transpose) edges.push_back({src, (int)d - 1, (WeightT)w});
      else edges.push_back({(int)d - 1, src, (WeightT)w});
    }
  }
  out.n = (int)m;
  return fill((int)m, (int)nnz, edges, out, o);
}

int gr2csr(const char *path, Csr1 &out, const ReadOpts &o) {
  if (o.verbose) printf("Reading 9th DIMACS (.gr) input file %s\n", path);
  Text t;
  if (!t.open(path)) return -1;
  const char *b, *e;
  do { if (!t.line(b, e)) return -2; } while (b < e && *b == 'c');
  if (b >= e || *b != 'p') return -2;
  const char *q = b + 1;
  while (q < e && (*q == ' ' || *q == '\t')) q++;
  skip_token(q, e);                                      // "sp"
  long m, nnz;
  if (!next_int(q, e, m) || !next_int(q, e, nnz)) return -2;
  if (o.verbose) printf("Before cleaning, the original num_vertices %ld num_edges %ld\n", m, nnz);
  std::vector<WEdge> edges;
  edges.reserve((size_t)nnz * (o.symmetrize ? 2 : 1));
  bool zero_based = false;
  long read = 0;
  while (read < nnz && t.line(b, e)) {
    if (b >= e || *b == 'c') continue;
    if (*b != 'a') continue;
    q = b + 1;
    long s, d;
    if (!next_int(q, e, s) || !next_int(q, e, d)) continue;
    if (s == 0 || d == 0) zero_based = true;
    // weights are ignored: every arc gets wt 1 (graph_io.h:172-191)
    if (o.symmetrize) { edges.push_back({(int)d, (int)s, 1}); edges.push_back({(int)s, (int)d, 1}); }
    else if (!o.transpose) edges.push_back({(int)s, (int)d, 1});
    else edges.push_back({(int)d, (int)s, 1});
    read++;
  }
  if (!zero_based) for (WEdge &w : edges) { w.src--; w.dst--; }
  out.n = (int)m;
  return fill((int)m, (int)nnz, edges, out, o);
}

static bool has(const char *s, const char *suffix) { return strstr(s, suffix) != nullptr; }

int read_graph_file(const char *path, Csr1 &out, const ReadOpts &o) {
  if (has(path, ".mtx")) return mtx2csr(path, out, o);       // graph_io.h:361-366: same precedence
  if (has(path, ".graph")) return graph2csr(path, out, o);
  if (has(path, ".gr")) return gr2csr(path, out, o);
  if (has(path, ".el")) return el2csr(path, out, o);
  return -4;
}

}  // namespace gdn

void read_graph(int argc, char *argv[], int &m, int &n, int &nnz, IndexT *&row_offsets,
                IndexT *&column_indices, int *&degree, WeightT *&weight, bool is_symmetrize,
                bool is_transpose, bool sorted, bool remove_selfloops, bool remove_redundents) {
  (void)argc;
  struct timeval t0, t1;
  gettimeofday(&t0, NULL);
  gdn::ReadOpts o;
  o.symmetrize = is_symmetrize; o.transpose = is_transpose; o.sorted = sorted;
  o.remove_selfloops = remove_selfloops; o.remove_redundents = remove_redundents;
  gdn::Csr1 c;
  int rc = gdn::read_graph_file(argv[1], c, o);
  if (rc == -4) { printf("Unrecognizable input file format\n"); exit(0); }
  if (rc != 0) { fprintf(stderr, "read_graph: cannot read %s (error %d)\n", argv[1], rc); exit(1); }
  gettimeofday(&t1, NULL);
  printf("\truntime [%s] = %f ms.\n", "read_graph",
         1000.0 * (t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) / 1000.0);
  printf("Calculating degree...");
  m = c.m; n = c.n; nnz = c.nnz;
  row_offsets = c.row_offsets; column_indices = c.column_indices; weight = c.weight; degree = c.degree;
  printf(" Done\n");
}
