// ref_driver.cc -- thin command-line harness AROUND the unmodified reference.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; it contains no reference code.
// oracle/Makefile compiles it together with the reference's own sources where
// they lie under /root/reference (src/bfs/omp_beamer.cc, src/pr/omp_base.cc,
// src/spmv/omp_base.cc, include/*.h) into oracle/_ref/ref_driver.  It exists so
// that (a) tools/make_golden.py can record the reference's outputs as fixtures
// and (b) bench.py --impl reference can time the reference's own OpenMP
// solvers on the GPU box's host cores (the reference prints its own
// "runtime [omp_*] = ... ms." line, which is what we parse).
//
// usage:
//   ref_driver csr  <filetype> <prefix> <symmetrize> <reverse> <out_prefix>
//   ref_driver bfs  <filetype> <prefix> <symmetrize> <reverse> <source> <out.i32> [repeat]
//   ref_driver bfs  <filetype> <prefix> <symmetrize> <reverse> <s0,s1,...> <out.i8>     (one BFS per source, one graph load;
//                   depths of all sources concatenated as int8, -1 = unreached)
//   ref_driver pr   <filetype> <prefix> <symmetrize> <out.f32> [repeat]
//   ref_driver spmv <filetype> <prefix> <symmetrize> <reverse> <seed> <out.f32> [repeat]
#include "bfs.h"     // src/bfs/bfs.h  -> common.h, csr_graph.h
#include "pr.h"      // src/pr/pr.h
#include "spmv.h"    // src/spmv/spmv.h
#include <random>
#include <cstdio>
#include <cstring>
#include <string>

template <typename T>
static void dump(const std::string &path, const T *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  if (n && fwrite(p, sizeof(T), n, f) != n) { perror("fwrite"); exit(2); }
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: see header of ref_driver.cc\n"); return 2; }
  std::string cmd = argv[1];
  if (cmd == "csr") {
    Graph g(argv[3], argv[2], atoi(argv[4]), atoi(argv[5]));
    std::string out = argv[6];
    size_t m = g.V(), nnz = g.E();
    dump(out + ".out_rowptr.u64", g.out_rowptr(), m + 1);
    dump(out + ".out_colidx.i32", g.out_colidx(), nnz);
    if (g.has_reverse_graph()) {
      dump(out + ".in_rowptr.u64", g.in_rowptr(), m + 1);
      dump(out + ".in_colidx.i32", g.in_colidx(), nnz);
    }
    printf("m %zu nnz %zu\n", m, nnz);
    return 0;
  }
  if (cmd == "bfs") {
    Graph g(argv[3], argv[2], atoi(argv[4]), atoi(argv[5]));
    std::vector<DistT> dist(g.V(), MYINFINITY);
    if (strchr(argv[6], ',')) {
      std::vector<int> sources;
      for (const char *p = argv[6]; *p;) { sources.push_back(atoi(p)); p = strchr(p, ','); if (!p) break; p++; }
      FILE *f = fopen(argv[7], "wb");
      if (!f) { perror(argv[7]); return 2; }
      std::vector<signed char> d8(g.V());
      for (int source : sources) {
        std::fill(dist.begin(), dist.end(), MYINFINITY);   // src/bfs/main.cc:21
        BFSSolver(g, source, &dist[0]);
        for (size_t i = 0; i < dist.size(); i++) {
          if (dist[i] != MYINFINITY && dist[i] > 126) { fprintf(stderr, "depth %d does not fit int8\n", dist[i]); return 3; }
          d8[i] = dist[i] == MYINFINITY ? -1 : (signed char)dist[i];
        }
        if (fwrite(&d8[0], 1, d8.size(), f) != d8.size()) { perror("fwrite"); return 2; }
      }
      fclose(f);
      return 0;
    }
    int source = atoi(argv[6]);
    int repeat = argc > 8 ? atoi(argv[8]) : 1;
    for (int r = 0; r < repeat; r++) {
      std::fill(dist.begin(), dist.end(), MYINFINITY);   // src/bfs/main.cc:21
      BFSSolver(g, source, &dist[0]);
    }
    dump(argv[7], &dist[0], dist.size());
    return 0;
  }
  if (cmd == "pr") {
    Graph g(argv[3], argv[2], atoi(argv[4]), 1);         // src/pr/main.cc:15
    int repeat = argc > 6 ? atoi(argv[6]) : 1;
    auto m = g.V();
    std::vector<ScoreT> scores(m);
    for (int r = 0; r < repeat; r++) {
      std::fill(scores.begin(), scores.end(), 1.0f / m);  // src/pr/main.cc:17-18
      PRSolver(g, &scores[0]);
    }
    dump(argv[5], &scores[0], scores.size());
    return 0;
  }
  if (cmd == "spmv") {
    Graph g(argv[3], argv[2], atoi(argv[4]), atoi(argv[5]));
    unsigned seed = (unsigned)atoi(argv[6]);
    int repeat = argc > 8 ? atoi(argv[8]) : 1;
    size_t m = g.V(), nnz = g.E();
    // seeded inputs instead of the constants of src/spmv/main.cc:27-37
    // (its rand() lines are commented out); same stream as gardenia_b200.
    std::mt19937 rng(seed);
    std::vector<ValueT> Ax(nnz), x(m), y(m);
    for (size_t i = 0; i < nnz; i++) Ax[i] = (rng() >> 8) * (1.0f / 16777216.0f);
    for (size_t i = 0; i < m; i++) x[i] = (rng() >> 8) * (1.0f / 16777216.0f);
    for (int r = 0; r < repeat; r++) {
      std::fill(y.begin(), y.end(), 0.0f);
      SpmvSolver(g, &Ax[0], &x[0], &y[0]);
    }
    dump(argv[7], &y[0], y.size());
    return 0;
  }
  fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 2;
}
