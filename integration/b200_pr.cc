// Drop-in PageRank variant object: defines PRSolver (src/pr/pr.h:31).
#include "pr.h"
#include "gdn_b200.h"
void PRSolver(Graph &g, ScoreT *scores) {
  auto m = g.V();
  std::vector<int32_t> out_degree(m);
  for (VertexId v = 0; v < m; v++) out_degree[v] = g.get_degree(v);   // src/pr/omp_base.cc:25
  printf("Launching CUDA PR solver (sm_100a, pull) ...\n");
  gdn_stats st;
  int rc = gdn_pagerank_pull(m, g.E(), g.in_rowptr(), g.in_colidx(), out_degree.data(), scores, kDamp, EPSILON, MAX_ITER, &st);
  if (rc != GDN_OK) { fprintf(stderr, "%s\n", gdn_last_error()); exit(EXIT_FAILURE); }
  for (int i = 0; i < st.iterations && i < MAX_ITER; i++) printf(" %2d    %lf\n", i + 1, st.pr_err[i]);   // src/pr/omp_base.cc:35
  printf("\titerations = %d.\n", st.iterations);
  printf("\truntime [b200_pull] = %f ms.\n", st.solve_ms);
}
