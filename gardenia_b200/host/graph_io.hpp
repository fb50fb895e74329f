// graph_io.hpp -- gen-1 text readers over raw row_offsets/column_indices
// arrays, with the calling convention of the reference's include/graph_io.h:
//   read_graph(argc, argv, m, n, nnz, row_offsets, column_indices, degree,
//              weight, is_symmetrize, is_transpose, sorted, remove_selfloops,
//              remove_redundents)                      (graph_io.h:357-377)
// and the per-format entry points mtx2csr / graph2csr / gr2csr / el2csr
// (graph_io.h:288-338, 247-285, 146-199, 202-244).  Arrays are malloc()ed and
// owned by the caller, 0-based, rows sorted, self loops and duplicates removed
// (fill_data, graph_io.h:25-143).
//
// Deliberate differences from the reference (documented in DESIGN.md):
//  * duplicates keep the weight of the first occurrence in file order (the
//    reference's unstable std::sort leaves this unspecified);
//  * a METIS .graph header with fmt=1/11 is honoured (neighbour,weight pairs);
//    the reference reads every token as a neighbour;
//  * a DIMACS .gr file that uses vertex id 0 is taken as 0-based (the
//    reference would index vertices[-1]; datasets/4.gr is such a file);
//  * out-of-range vertex ids return an error instead of undefined behaviour.
#pragma once
#include "gdn_types.hpp"

namespace gdn {

struct Csr1 {           // result of a gen-1 read
  int m = 0, n = 0, nnz = 0;
  IndexT *row_offsets = nullptr;
  IndexT *column_indices = nullptr;
  WeightT *weight = nullptr;
  int *degree = nullptr;
};

struct ReadOpts {
  bool symmetrize = false, transpose = false, sorted = true, remove_selfloops = true,
       remove_redundents = true, verbose = true;
};

// Non-exiting readers: 0 on success, <0 on error.
int mtx2csr(const char *path, Csr1 &out, const ReadOpts &o);
int graph2csr(const char *path, Csr1 &out, const ReadOpts &o);
int gr2csr(const char *path, Csr1 &out, const ReadOpts &o);
int el2csr(const char *path, Csr1 &out, const ReadOpts &o);
int read_graph_file(const char *path, Csr1 &out, const ReadOpts &o);   // dispatch on suffix

}  // namespace gdn

// Reference-compatible free function (graph_io.h:357-377): same argument list,
// prints the same progress lines, exit(0) on an unrecognised suffix.
void read_graph(int argc, char *argv[], int &m, int &n, int &nnz, IndexT *&row_offsets,
                IndexT *&column_indices, int *&degree, WeightT *&weight,
                bool is_symmetrize = false, bool is_transpose = false, bool sorted = true,
                bool remove_selfloops = true, bool remove_redundents = true);
