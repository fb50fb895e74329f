// csr_graph.hpp -- host CSR container with the interface of the reference's
// gen-2 `Graph` (include/csr_graph.h:46-351): same constructor contract
// (prefix, filetype, symmetrize, need_reverse), same accessors
// (V/E/N/in_neigh/out_neigh/get_degree/out_rowptr/out_colidx/in_rowptr/
// in_colidx/has_reverse_graph), uint64 row offsets and int32 column indices.
//
// Written from scratch: edges are bucketed with a counting sort and each row
// is sorted + uniqued in parallel (the reference does a serial std::sort plus
// an O(deg^2) vector::erase dedup, include/csr_graph.h:122-143).  The result
// (sorted, self-loop-free, deduplicated rows) is identical.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <utility>
#include "gdn_types.hpp"

namespace gdn {

// A neighbour range; same shape as the reference's VertexSet (csr_graph.h:13-36).
class VertexSet {
  const VertexId *ptr_;
  VertexId size_;
 public:
  VertexSet() : ptr_(nullptr), size_(0) {}
  VertexSet(const VertexId *p, VertexId s) : ptr_(p), size_(s) {}
  VertexId size() const { return size_; }
  const VertexId *begin() const { return ptr_; }
  const VertexId *end() const { return ptr_ + size_; }
};

struct EdgePair32 { VertexId u, v; };

// Bucket `edges` by source, sort every row, drop duplicates (and self loops if
// asked).  Returns max degree.  rowptr/col are new[]-allocated.
VertexId build_csr(int64_t m, const EdgePair32 *edges, int64_t n_edges, bool transpose,
                   bool remove_self, uint64_t *&rowptr, VertexId *&col, uint64_t &nnz);

// Transposed CSR (csr_graph.h:170-194); rows come out sorted by construction.
void transpose_csr(int64_t m, const uint64_t *rowptr, const VertexId *col,
                   uint64_t *&t_rowptr, VertexId *&t_col);

enum LoadStatus { kLoadOk = 0, kLoadNoFile = -1, kLoadBadHeader = -2, kLoadBadVertex = -3,
                  kLoadBadType = -4, kLoadDegenerate = -5 };

class Graph {
  bool directed_ = false;
  bool has_reverse_ = false;
  VertexId n_vertices_ = 0;
  uint64_t n_edges_ = 0;
  VertexId max_degree_ = 0;
  uint64_t *vertices_ = nullptr, *reverse_vertices_ = nullptr;
  VertexId *edges_ = nullptr, *reverse_edges_ = nullptr;
  size_t map_bytes_[2] = {0, 0};          // > 0: vertices_ / edges_ are file mappings (load_bin_mapped), not new[] blocks
  void release();
  void finish(bool symmetrize, bool need_reverse, bool verbose);

 public:
  Graph() {}
  // Reference contract (csr_graph.h:211-250): prints the same lines and
  // exit(1)s on a missing file or a degenerate graph.
  Graph(std::string prefix, std::string filetype = "bin", bool symmetrize = false,
        bool need_reverse = false);
  ~Graph() { release(); }
  Graph(const Graph &) = delete;
  Graph &operator=(const Graph &) = delete;

  // Non-exiting loaders (used by the C-ABI host layer).
  int load(const std::string &prefix, const std::string &filetype, bool symmetrize,
           bool need_reverse, bool verbose);
  int load_mtx(const std::string &fname, bool symmetrize, bool need_reverse, bool verbose);
  int load_bin(const std::string &prefix, bool symmetrize, bool need_reverse, bool verbose);
  // The binary triple mapped copy-on-write instead of read: N processes of one box share ONE copy of a cached graph
  // in the page cache (filetype "bin:mmap"; symmetric graphs, nothing is ever written).
  int load_bin_mapped(const std::string &prefix, bool verbose);
  // Serialized graph of the GAP-style reader (include/reader.h:259-316); 32- or 64-bit offsets, recognised by file size.
  int load_sg(const std::string &fname, bool need_reverse, bool verbose);
  int write_sg(const std::string &fname, int offset_bytes = 4) const;
  // Adopt an already squished symmetric CSR (generator output); takes ownership.
  void adopt_symmetric(VertexId m, uint64_t nnz, uint64_t *rowptr, VertexId *col, VertexId max_degree);
  // Adopt a directed CSR pair; takes ownership.
  void adopt_directed(VertexId m, uint64_t nnz, uint64_t *rowptr, VertexId *col,
                      uint64_t *t_rowptr, VertexId *t_col, VertexId max_degree);
  // Write the reference's binary triple (csr_graph.h:218-233).
  int write_bin(const std::string &prefix) const;

  VertexSet N(VertexId v) const { return VertexSet(edges_ + vertices_[v], VertexId(vertices_[v + 1] - vertices_[v])); }
  VertexSet out_neigh(VertexId v, VertexId start_offset = 0) const {
    uint64_t b = vertices_[v], e = vertices_[v + 1];
    b += std::min<uint64_t>(start_offset, e - b);
    return VertexSet(edges_ + b, VertexId(e - b));
  }
  VertexSet in_neigh(VertexId v) const {
    return VertexSet(reverse_edges_ + reverse_vertices_[v], VertexId(reverse_vertices_[v + 1] - reverse_vertices_[v]));
  }
  VertexId V() const { return n_vertices_; }
  size_t E() const { return n_edges_; }
  size_t size() const { return size_t(n_vertices_); }
  size_t sizeEdges() const { return n_edges_; }
  VertexId num_vertices() const { return n_vertices_; }
  size_t num_edges() const { return n_edges_; }
  VertexId get_degree(VertexId v) const { return VertexId(vertices_[v + 1] - vertices_[v]); }
  VertexId out_degree(VertexId v) const { return get_degree(v); }
  uint64_t edge_begin(VertexId v) const { return vertices_[v]; }
  uint64_t edge_end(VertexId v) const { return vertices_[v + 1]; }
  VertexId getEdgeDst(uint64_t e) const { return edges_[e]; }
  VertexId get_max_degree() const { return max_degree_; }
  bool is_directed() const { return directed_; }
  bool has_reverse_graph() const { return has_reverse_; }
  uint64_t *out_rowptr() const { return vertices_; }
  VertexId *out_colidx() const { return edges_; }
  uint64_t *in_rowptr() const { return reverse_vertices_; }
  VertexId *in_colidx() const { return reverse_edges_; }
};

}  // namespace gdn
