// csr_graph.cc -- see csr_graph.hpp.  Reader semantics follow the reference's
// include/csr_graph.h:55-250 (cited inline); the algorithms are our own.
#include "csr_graph.hpp"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <omp.h>
#include <cerrno>
#include <vector>

namespace gdn {

// Exclusive prefix sum of 64-bit counts, blocked over threads.
static void prefix_sum(const uint64_t *cnt, int64_t n, uint64_t *out /* n+1 */) {
  int nt = omp_get_max_threads();
  std::vector<uint64_t> part(nt + 1, 0);
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num();
    int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    uint64_t s = 0;
    for (int64_t i = lo; i < hi; i++) s += cnt[i];
    part[t + 1] = s;
#pragma omp barrier
#pragma omp single
    for (int i = 0; i < nt; i++) part[i + 1] += part[i];
    s = part[t];
    for (int64_t i = lo; i < hi; i++) { uint64_t c = cnt[i]; out[i] = s; s += c; }
  }
  out[n] = part[nt];
}

VertexId build_csr(int64_t m, const EdgePair32 *edges, int64_t n_edges, bool transpose,
                   bool remove_self, uint64_t *&rowptr, VertexId *&col, uint64_t &nnz) {
  std::vector<uint64_t> cnt(m + 1, 0);
#pragma omp parallel for
  for (int64_t e = 0; e < n_edges; e++) {
    VertexId s = transpose ? edges[e].v : edges[e].u;
#pragma omp atomic
    cnt[s]++;
  }
  std::vector<uint64_t> off(m + 1);
  prefix_sum(cnt.data(), m, off.data());
  std::vector<uint64_t> cur(off.begin(), off.end() - 1);
  VertexId *raw = new VertexId[std::max<uint64_t>(off[m], 1)];
#pragma omp parallel for
  for (int64_t e = 0; e < n_edges; e++) {
    VertexId s = transpose ? edges[e].v : edges[e].u;
    VertexId d = transpose ? edges[e].u : edges[e].v;
    uint64_t pos;
#pragma omp atomic capture
    pos = cur[s]++;
    raw[pos] = d;
  }
  // sort + unique (+ drop self) each row in place; cnt[] becomes the new degree
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < m; v++) {
    VertexId *b = raw + off[v], *e = raw + off[v + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    if (remove_self) e = std::remove(b, e, (VertexId)v);
    cnt[v] = (uint64_t)(e - b);
  }
  rowptr = new uint64_t[m + 1];
  prefix_sum(cnt.data(), m, rowptr);
  nnz = rowptr[m];
  col = new VertexId[std::max<uint64_t>(nnz, 1)];
  VertexId maxdeg = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : maxdeg)
  for (int64_t v = 0; v < m; v++) {
    std::copy(raw + off[v], raw + off[v] + cnt[v], col + rowptr[v]);
    maxdeg = std::max(maxdeg, (VertexId)cnt[v]);
  }
  delete[] raw;
  return maxdeg;
}

void transpose_csr(int64_t m, const uint64_t *rowptr, const VertexId *col,
                   uint64_t *&t_rowptr, VertexId *&t_col) {
  uint64_t nnz = rowptr[m];
  std::vector<uint64_t> cnt(m + 1, 0);
#pragma omp parallel for
  for (int64_t e = 0; e < (int64_t)nnz; e++) {
#pragma omp atomic
    cnt[col[e]]++;
  }
  t_rowptr = new uint64_t[m + 1];
  prefix_sum(cnt.data(), m, t_rowptr);
  t_col = new VertexId[std::max<uint64_t>(nnz, 1)];
  // Sources are visited in ascending order inside each destination bucket only
  // if the fill is sequential per bucket; do a parallel fill and sort rows.
  std::vector<uint64_t> cur(t_rowptr, t_rowptr + m);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < m; v++)
    for (uint64_t e = rowptr[v]; e < rowptr[v + 1]; e++) {
      uint64_t pos;
#pragma omp atomic capture
      pos = cur[col[e]]++;
      t_col[pos] = (VertexId)v;
    }
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t v = 0; v < m; v++) std::sort(t_col + t_rowptr[v], t_col + t_rowptr[v + 1]);
}

void Graph::release() {
  if (reverse_vertices_ != vertices_) delete[] reverse_vertices_;
  if (reverse_edges_ != edges_) delete[] reverse_edges_;
  if (map_bytes_[0]) { munmap(vertices_, map_bytes_[0]); vertices_ = nullptr; }
  if (map_bytes_[1]) { munmap(edges_, map_bytes_[1]); edges_ = nullptr; }
  map_bytes_[0] = map_bytes_[1] = 0;
  delete[] vertices_;
  delete[] edges_;
  vertices_ = reverse_vertices_ = nullptr;
  edges_ = reverse_edges_ = nullptr;
}

// csr_graph.h:234-247: reverse CSR is built only for (!symmetrize && need_reverse);
// a symmetrized graph aliases the forward arrays.
void Graph::finish(bool symmetrize, bool need_reverse, bool verbose) {
  directed_ = false;
  has_reverse_ = false;
  if (!symmetrize && need_reverse) {
    transpose_csr(n_vertices_, vertices_, edges_, reverse_vertices_, reverse_edges_);
    directed_ = true;
    has_reverse_ = true;
    if (verbose) printf("This graph maintains both incomming and outgoing edge-list\n");
  }
  if (symmetrize) {
    if (verbose) printf("This graph is symmetrized\n");
    reverse_vertices_ = vertices_;
    reverse_edges_ = edges_;
    has_reverse_ = true;
  }
}

static const char *skip_ws(const char *p, const char *end) {
  while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
  return p;
}
// Parse a decimal integer; returns false if none at p.
static bool parse_int(const char *&p, const char *end, long &out) {
  p = skip_ws(p, end);
  const char *q = p;
  bool neg = false;
  if (q < end && (*q == '-' || *q == '+')) { neg = (*q == '-'); q++; }
  if (q >= end || *q < '0' || *q > '9') return false;
  long v = 0;
  while (q < end && *q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); q++; }
  out = neg ? -v : v;
  p = q;
  return true;
}

int Graph::load_mtx(const std::string &fname, bool symmetrize, bool need_reverse, bool verbose) {
  if (verbose) std::cout << "Reading (.mtx) input file " << fname << "\n";
  std::ifstream in(fname.c_str(), std::ios::binary);
  if (!in) return kLoadNoFile;
  std::string buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const char *p = buf.data(), *end = p + buf.size();
  auto next_eol = [&](const char *q) { while (q < end && *q != '\n') q++; return q; };
  // csr_graph.h:86-92: skip header lines whose first character is '%'
  while (p < end && *p == '%') { p = next_eol(p); if (p < end) p++; }
  const char *eol = next_eol(p);
  long m = 0, n = 0, nnz_hdr = 0;
  {
    const char *q = p;
    if (!parse_int(q, eol, m) || !parse_int(q, eol, n)) return kLoadBadHeader;
    parse_int(q, eol, nnz_hdr);
  }
  if (m != n && verbose) printf("Warning, m(%ld) != n(%ld)\n", m, n);
  p = eol < end ? eol + 1 : end;
  std::vector<EdgePair32> el;
  el.reserve(nnz_hdr > 0 ? (size_t)nnz_hdr * (symmetrize ? 2 : 1) : 16);
  // csr_graph.h:66-73,105-118: skip empty and '#' lines; stop at the first
  // line that does not start with two integers; weights are ignored.
  while (p < end) {
    eol = next_eol(p);
    const char *q = p;
    bool blank = (eol == p) || (*p == '#') || (eol - p == 1 && *p == '\r');
    if (!blank) {
      long a, b;
      if (!parse_int(q, eol, a) || !parse_int(q, eol, b)) break;
      if (a != b) {                                   // self loop dropped, :108
        if (a < 1 || a > m || b < 1 || b > m) return kLoadBadVertex;
        el.push_back({(VertexId)(a - 1), (VertexId)(b - 1)});
        if (symmetrize) el.push_back({(VertexId)(b - 1), (VertexId)(a - 1)});   // :113-116
      }
    }
    p = eol < end ? eol + 1 : end;
  }
  n_vertices_ = (VertexId)m;
  uint64_t raw = el.size();
  max_degree_ = build_csr(m, el.data(), (int64_t)el.size(), false, false, vertices_, edges_, n_edges_);
  if (verbose) {
    printf("Removing redundent edges... %d redundent edges are removed\n", (int)(raw - n_edges_));
    std::cout << "|V| " << n_vertices_ << " |E| " << n_edges_ << "\n";
  }
  finish(symmetrize, need_reverse, verbose);
  return kLoadOk;
}

template <typename T>
static bool read_all(const std::string &fname, T *dst, size_t n) {
  FILE *f = fopen(fname.c_str(), "rb");
  if (!f) return false;
  size_t got = fread(dst, sizeof(T), n, f);
  fclose(f);
  return got == n;
}

// csr_graph.h:218-233
int Graph::load_bin(const std::string &prefix, bool symmetrize, bool need_reverse, bool verbose) {
  std::ifstream meta((prefix + ".meta.txt").c_str());
  if (!meta) return kLoadNoFile;
  long long nv = 0, ne = 0, maxd = 0;
  int vid_size = 0;
  meta >> nv >> ne >> vid_size >> maxd;
  if (!meta || vid_size != (int)sizeof(VertexId)) return kLoadBadHeader;
  n_vertices_ = (VertexId)nv;
  n_edges_ = (uint64_t)ne;
  max_degree_ = (VertexId)maxd;
  if (verbose) std::cout << "|V| " << n_vertices_ << " |E| " << n_edges_ << "\n";
  vertices_ = new uint64_t[nv + 1];
  edges_ = new VertexId[std::max<long long>(ne, 1)];
  if (!read_all(prefix + ".vertex.bin", vertices_, (size_t)nv + 1) ||
      !read_all(prefix + ".edge.bin", edges_, (size_t)ne)) {
    std::cerr << "Failed to open file: " << prefix << ".{vertex,edge}.bin\n";
    return kLoadNoFile;
  }
  finish(symmetrize, need_reverse, verbose);
  return kLoadOk;
}

// Serialized graph of the GAP-style reader (include/reader.h:259-316):
//   bool directed | SGOffset num_edges | SGOffset num_vertices | SGOffset offsets[nv + 1] | int32 neighs[ne]
//   and, for a directed graph, the inverse CSR: offsets[nv + 1] | neighs[ne].
// SGOffset is `int` in the reference (include/graph.h:77-79) and int64_t in upstream GAP files: the width is recognised
// from the file size, which either layout determines exactly.  The file holds a finished CSR: it is adopted as it is
// (symmetrize has nothing to do, as in include/builder.h:264); only its consistency is checked.
template <typename Off>
static int read_sg_body(FILE *f, long long fsize, bool directed, bool want_inverse, VertexId &nv_out, uint64_t &ne_out,
                        uint64_t *&rowptr, VertexId *&col, uint64_t *&t_rowptr, VertexId *&t_col) {
  Off hdr[2];
  if (fseek(f, 1, SEEK_SET) != 0 || fread(hdr, sizeof(Off), 2, f) != 2) return kLoadBadHeader;
  const long long ne = (long long)hdr[0], nv = (long long)hdr[1];
  if (ne < 0 || nv < 0 || nv > 0x7fffffffll) return kLoadBadHeader;
  const long long part = (nv + 1) * (long long)sizeof(Off) + ne * (long long)sizeof(VertexId);
  if (fsize != 1 + 2 * (long long)sizeof(Off) + (directed ? 2 : 1) * part) return kLoadBadHeader;
  auto read_csr = [&](uint64_t *&rp, VertexId *&ci) -> int {
    rp = new uint64_t[nv + 1];
    ci = new VertexId[std::max<long long>(ne, 1)];
    if (sizeof(Off) == sizeof(uint64_t)) {
      if (fread(rp, sizeof(uint64_t), (size_t)nv + 1, f) != (size_t)nv + 1) return kLoadBadHeader;
    } else {
      Off *tmp = new Off[nv + 1];
      const bool ok = fread(tmp, sizeof(Off), (size_t)nv + 1, f) == (size_t)nv + 1;
      for (long long i = 0; ok && i <= nv; i++) rp[i] = (uint64_t)(int64_t)tmp[i];
      delete[] tmp;
      if (!ok) return kLoadBadHeader;
    }
    if (fread(ci, sizeof(VertexId), (size_t)ne, f) != (size_t)ne) return kLoadBadHeader;
    if (rp[0] != 0 || rp[nv] != (uint64_t)ne) return kLoadBadVertex;
    long long bad = 0;
#pragma omp parallel for reduction(+ : bad)
    for (long long v = 0; v < nv; v++) bad += rp[v + 1] < rp[v] || rp[v + 1] > (uint64_t)ne;
    if (bad) return kLoadBadVertex;
#pragma omp parallel for reduction(+ : bad)
    for (long long e = 0; e < ne; e++) bad += ci[e] < 0 || ci[e] >= nv;
    return bad ? kLoadBadVertex : kLoadOk;
  };
  int rc = read_csr(rowptr, col);
  if (rc == kLoadOk && directed && want_inverse) rc = read_csr(t_rowptr, t_col);
  nv_out = (VertexId)nv;
  ne_out = (uint64_t)ne;
  return rc;
}

int Graph::load_sg(const std::string &fname, bool need_reverse, bool verbose) {
  if (verbose) std::cout << "Reading (.sg) input file " << fname << "\n";
  FILE *f = fopen(fname.c_str(), "rb");
  if (!f) return kLoadNoFile;
  fseek(f, 0, SEEK_END);
  const long long fsize = ftell(f);
  unsigned char dir = 0;
  if (fsize < 9 || fseek(f, 0, SEEK_SET) != 0 || fread(&dir, 1, 1, f) != 1 || dir > 1) { fclose(f); return kLoadBadHeader; }
  const bool directed = dir != 0;
  uint64_t *rp = nullptr, *trp = nullptr;
  VertexId *ci = nullptr, *tci = nullptr;
  VertexId nv = 0;
  uint64_t ne = 0;
  int rc = read_sg_body<int32_t>(f, fsize, directed, need_reverse, nv, ne, rp, ci, trp, tci);
  if (rc == kLoadBadHeader) {                      // not the reference's 32-bit offsets: an upstream GAP file?
    delete[] rp; delete[] ci; delete[] trp; delete[] tci;
    rp = trp = nullptr; ci = tci = nullptr;
    rc = read_sg_body<int64_t>(f, fsize, directed, need_reverse, nv, ne, rp, ci, trp, tci);
  }
  fclose(f);
  if (rc != kLoadOk) { delete[] rp; delete[] ci; delete[] trp; delete[] tci; return rc; }
  VertexId maxd = 0;
#pragma omp parallel for reduction(max : maxd)
  for (VertexId v = 0; v < nv; v++) maxd = std::max(maxd, (VertexId)(rp[v + 1] - rp[v]));
  if (verbose) std::cout << "|V| " << nv << " |E| " << ne << "\n";
  if (!directed) adopt_symmetric(nv, ne, rp, ci, maxd);
  else if (need_reverse) adopt_directed(nv, ne, rp, ci, trp, tci, maxd);
  else {
    release();
    n_vertices_ = nv; n_edges_ = ne; vertices_ = rp; edges_ = ci; max_degree_ = maxd;
    directed_ = true; has_reverse_ = false;
  }
  return kLoadOk;
}

int Graph::write_sg(const std::string &fname, int offset_bytes) const {
  if (offset_bytes != 4 && offset_bytes != 8) return -1;
  if (offset_bytes == 4 && n_edges_ > 0x7fffffffull) return -1;       // does not fit the reference's SGOffset
  if (directed_ && !has_reverse_) return -1;                          // a directed .sg carries its inverse
  FILE *f = fopen(fname.c_str(), "wb");
  if (!f) return -1;
  const unsigned char dir = directed_ ? 1 : 0;
  bool ok = fwrite(&dir, 1, 1, f) == 1;
  auto put = [&](uint64_t v) {
    if (offset_bytes == 4) { const int32_t w = (int32_t)v; ok = ok && fwrite(&w, 4, 1, f) == 1; }
    else { const int64_t w = (int64_t)v; ok = ok && fwrite(&w, 8, 1, f) == 1; }
  };
  put(n_edges_);
  put((uint64_t)n_vertices_);
  auto put_csr = [&](const uint64_t *rp, const VertexId *ci) {
    if (offset_bytes == 8) ok = ok && fwrite(rp, 8, (size_t)n_vertices_ + 1, f) == (size_t)n_vertices_ + 1;
    else {
      int32_t *tmp = new int32_t[(size_t)n_vertices_ + 1];
      for (int64_t i = 0; i <= n_vertices_; i++) tmp[i] = (int32_t)rp[i];
      ok = ok && fwrite(tmp, 4, (size_t)n_vertices_ + 1, f) == (size_t)n_vertices_ + 1;
      delete[] tmp;
    }
    ok = ok && fwrite(ci, sizeof(VertexId), n_edges_, f) == n_edges_;
  };
  put_csr(vertices_, edges_);
  if (directed_) put_csr(reverse_vertices_, reverse_edges_);
  return fclose(f) == 0 && ok ? 0 : -1;
}

static void *map_file(const std::string &fname, size_t bytes) {
  const int fd = open(fname.c_str(), O_RDONLY);
  if (fd < 0) return nullptr;
  struct stat st;
  void *p = MAP_FAILED;
  if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes && bytes > 0)
    p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);       // private: page-lockable, never written
  close(fd);
  return p == MAP_FAILED ? nullptr : p;
}

int Graph::load_bin_mapped(const std::string &prefix, bool verbose) {
  std::ifstream meta((prefix + ".meta.txt").c_str());
  if (!meta) return kLoadNoFile;
  long long nv = 0, ne = 0, maxd = 0;
  int vid_size = 0;
  meta >> nv >> ne >> vid_size >> maxd;
  if (!meta || vid_size != (int)sizeof(VertexId) || nv < 0 || ne <= 0) return kLoadBadHeader;
  const size_t vb = sizeof(uint64_t) * ((size_t)nv + 1), eb = sizeof(VertexId) * (size_t)ne;
  void *v = map_file(prefix + ".vertex.bin", vb), *e = map_file(prefix + ".edge.bin", eb);
  if (!v || !e) {
    if (v) munmap(v, vb);
    if (e) munmap(e, eb);
    return kLoadNoFile;
  }
  n_vertices_ = (VertexId)nv; n_edges_ = (uint64_t)ne; max_degree_ = (VertexId)maxd;
  vertices_ = (uint64_t *)v; edges_ = (VertexId *)e;
  map_bytes_[0] = vb; map_bytes_[1] = eb;
  if (verbose) std::cout << "|V| " << n_vertices_ << " |E| " << n_edges_ << " (mapped)\n";
  finish(true, false, verbose);
  return kLoadOk;
}

int Graph::load(const std::string &prefix, const std::string &filetype, bool symmetrize,
                bool need_reverse, bool verbose) {
  release();
  int rc;
  if (filetype == "bin:mmap") {
    if (!symmetrize) return kLoadBadType;                  // (the reverse graph of a directed one would have to be built)
    rc = load_bin_mapped(prefix, verbose);
    if (rc != kLoadOk) return rc;
    if (max_degree_ == 0 || max_degree_ >= n_vertices_) return kLoadDegenerate;
    return kLoadOk;
  }
  if (filetype == "sg") {
    // the .sg of the GAP-style drivers (include/reader.h:259-316) keeps whatever graph it was built from: the
    // degenerate-graph check of the gen-2 constructor does not apply
    const bool has_suffix = prefix.size() > 3 && prefix.compare(prefix.size() - 3, 3, ".sg") == 0;
    return load_sg(has_suffix ? prefix : prefix + ".sg", need_reverse, verbose);
  }
  if (filetype == "mtx") rc = load_mtx(prefix + ".mtx", symmetrize, need_reverse, verbose);
  else if (filetype == "bin") rc = load_bin(prefix, symmetrize, need_reverse, verbose);
  else return kLoadBadType;
  if (rc != kLoadOk) return rc;
  // csr_graph.h:248
  if (max_degree_ == 0 || max_degree_ >= n_vertices_) return kLoadDegenerate;
  return kLoadOk;
}

Graph::Graph(std::string prefix, std::string filetype, bool symmetrize, bool need_reverse) {
  int rc = load(prefix, filetype, symmetrize, need_reverse, true);
  if (rc == kLoadNoFile) { std::cout << "File not available\n"; exit(1); }
  if (rc != kLoadOk) exit(1);
}

void Graph::adopt_symmetric(VertexId m, uint64_t nnz, uint64_t *rowptr, VertexId *col, VertexId max_degree) {
  release();
  n_vertices_ = m; n_edges_ = nnz; vertices_ = rowptr; edges_ = col; max_degree_ = max_degree;
  reverse_vertices_ = vertices_; reverse_edges_ = edges_;
  directed_ = false; has_reverse_ = true;
}

void Graph::adopt_directed(VertexId m, uint64_t nnz, uint64_t *rowptr, VertexId *col,
                           uint64_t *t_rowptr, VertexId *t_col, VertexId max_degree) {
  release();
  n_vertices_ = m; n_edges_ = nnz; vertices_ = rowptr; edges_ = col; max_degree_ = max_degree;
  reverse_vertices_ = t_rowptr; reverse_edges_ = t_col;
  directed_ = true; has_reverse_ = true;
}

int Graph::write_bin(const std::string &prefix) const {
  FILE *f = fopen((prefix + ".meta.txt").c_str(), "w");
  if (!f) return -1;
  fprintf(f, "%d\n%llu\n%d\n%d\n", n_vertices_, (unsigned long long)n_edges_, (int)sizeof(VertexId), max_degree_);
  fclose(f);
  f = fopen((prefix + ".vertex.bin").c_str(), "wb");
  if (!f) return -1;
  fwrite(vertices_, sizeof(uint64_t), (size_t)n_vertices_ + 1, f);
  fclose(f);
  f = fopen((prefix + ".edge.bin").c_str(), "wb");
  if (!f) return -1;
  fwrite(edges_, sizeof(VertexId), n_edges_, f);
  fclose(f);
  return 0;
}

}  // namespace gdn
