// generator.hpp -- deterministic Kronecker (R-MAT) / uniform-random graph
// generator + CSR builder, 64-bit safe.
//
// Same edge streams as the reference's GAP-derived generator
// (include/generator.h:39-127, seed include/misc.h:18, block 2^18) and the same
// final graph as its builder + squish (include/builder.h:66-119,152-195,
// 220-274; synthetic graphs are force-symmetrized, include/command_line.h:75-76;
// m = max id + 1, include/builder.h:244-245), but with int64 edge counts and
// offsets so that scales 26-27 work (the reference's `int` SGOffset overflows
// there, include/graph.h:77-79).
#pragma once
#include <cstdint>
#include "csr_graph.hpp"

namespace gdn {

static const int64_t kRandSeed = 27491095;       // include/misc.h:18
static const int64_t kGenBlockSize = 1 << 18;    // include/generator.h:126

// Raw (unsymmetrized, unpermuted-for-uniform) edge list: num_edges = degree << scale.
// `el` must hold num_edges entries.
void make_rmat_el(int scale, int degree, EdgePair32 *el);      // generator.h:81-114 incl. PermuteIDs
void make_uniform_el(int scale, int degree, EdgePair32 *el);   // generator.h:64-79

// Symmetrize + squish (sort/unique/remove-self per row).  Returns max degree;
// m_out = max id + 1.
VertexId build_symmetric_csr(const EdgePair32 *el, int64_t n_edges, int64_t &m_out,
                             uint64_t *&rowptr, VertexId *&col, uint64_t &nnz);

// One call: generate + build.  uniform=false -> Kronecker (-g), true -> uniform (-u).
void generate_graph(Graph &g, bool uniform, int scale, int degree = 16);

}  // namespace gdn
