// Drop-in BFS variant object for the reference tree: defines BFSSolver (src/bfs/bfs.h:43)
// and forwards to the C-ABI.  Links with the UNMODIFIED src/bfs/main.cc + verifier.cc.
#include "bfs.h"
#include "gdn_b200.h"
void BFSSolver(Graph &g, int source, DistT *dist) {
  if (!g.has_reverse_graph()) {   // same refusal as src/bfs/omp_beamer.cc:98-102
    std::cout << "This algorithm requires the reverse graph constructed for directed graph\n";
    std::cout << "Please set reverse to 1 in the command line\n";
    exit(1);
  }
  printf("Launching CUDA BFS solver (sm_100a, direction-optimizing) ...\n");
  gdn_stats st;
  int rc = gdn_bfs(g.V(), g.E(), g.out_rowptr(), g.out_colidx(), g.in_rowptr(), g.in_colidx(), source, dist, nullptr, &st);
  if (rc != GDN_OK) { fprintf(stderr, "%s\n", gdn_last_error()); exit(EXIT_FAILURE); }   // include/cutil_subset.h:4-12
  printf("\titerations = %d.\n", st.iterations);
  printf("\truntime [b200_hybrid] = %f ms.\n", st.solve_ms);
}
