// ref_gen.cc -- harness AROUND the reference's GAP-derived generator/builder.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; oracle/Makefile compiles it
// against the reference headers where they lie (include/builder.h,
// include/generator.h, include/command_line.h).  It instantiates the
// reference Builder exactly as SURVEY §3.4 describes and writes the squished
// CSR in the reference's own binary triple (include/csr_graph.h:218-233):
//   <prefix>.meta.txt   "m nnz sizeof(vid) max_degree"
//   <prefix>.vertex.bin uint64[m+1]
//   <prefix>.edge.bin   int32[nnz]
// The reference builder is `int`-limited (include/graph.h:77-79), so this is
// only usable up to scale 25; it is the cross-check for our 64-bit generator.
//
// With -S <file.sg> the graph comes from the reference's serialized-graph reader instead (include/reader.h:259-316,
// Reader::ReadSerializedGraph called directly: Builder::MakeGraph, include/builder.h:258-274, reads the file and then
// overwrites the graph with an empty edge list's) -- the cross-check of our .sg writer and reader.
//
// usage: ref_gen (-g <scale> | -u <scale> | -S <file.sg>) [-k <degree>] -o <prefix>
#include "common.h"
#include "builder.h"
#include <cstdio>
#include <string>

int main(int argc, char **argv) {
  std::string out, sg;
  std::vector<char *> args;
  for (int i = 0; i < argc; i++) {
    if (std::string(argv[i]) == "-o" && i + 1 < argc) { out = argv[++i]; continue; }
    if (std::string(argv[i]) == "-S" && i + 1 < argc) { sg = argv[++i]; continue; }
    args.push_back(argv[i]);
  }
  if (out.empty()) { fprintf(stderr, "usage: ref_gen (-g s | -u s) [-k d] -o prefix\n"); return 2; }
  Graph g;
  if (!sg.empty()) {
    Reader<VertexID, VertexID, WeightT, true> r(sg);
    r.ReadSerializedGraph(g);
    g.SetupRowptr();                 // the reader fills the pointer index only (include/graph.h:272-282)
  } else {
    CLBase cli((int)args.size(), args.data(), "ref_gen");
    if (!cli.ParseArgs()) return 2;
    Builder b(cli);
    b.MakeGraph(g);
  }
  int64_t m = g.num_vertices();
  const int *rowptr = g.out_rowptr();
  const int *col = g.out_colidx();
  int64_t nnz = rowptr[m];
  std::vector<uint64_t> off(m + 1);
  int maxdeg = 0;
  for (int64_t i = 0; i <= m; i++) off[i] = (uint64_t)rowptr[i];
  for (int64_t i = 0; i < m; i++) maxdeg = std::max(maxdeg, rowptr[i + 1] - rowptr[i]);
  FILE *f = fopen((out + ".meta.txt").c_str(), "w");
  if (!f) { perror("meta"); return 2; }
  fprintf(f, "%ld\n%ld\n%d\n%d\n", (long)m, (long)nnz, 4, maxdeg);
  fclose(f);
  f = fopen((out + ".vertex.bin").c_str(), "wb");
  fwrite(off.data(), 8, m + 1, f);
  fclose(f);
  f = fopen((out + ".edge.bin").c_str(), "wb");
  fwrite(col, 4, nnz, f);
  fclose(f);
  printf("m %ld nnz %ld maxdeg %d\n", (long)m, (long)nnz, maxdeg);
  return 0;
}
