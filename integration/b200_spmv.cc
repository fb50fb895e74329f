// Drop-in SpMV variant object: defines SpmvSolver (src/spmv/spmv.h:29).
#include "spmv.h"
#include "spmv_util.h"
#include "gdn_b200.h"
void SpmvSolver(Graph &g, const ValueT *Ax, const ValueT *x, ValueT *y) {
  printf("Launching CUDA SpMV solver (sm_100a) ...\n");
  gdn_stats st;
  int rc = gdn_spmv_csr(g.V(), g.E(), g.in_rowptr(), g.in_colidx(), Ax, x, y, &st);   // in-CSR, src/spmv/omp_base.cc:10-11
  if (rc != GDN_OK) { fprintf(stderr, "%s\n", gdn_last_error()); exit(EXIT_FAILURE); }
  double t = st.solve_ms;
  printf("\truntime [b200_csr] = %.4f ms ( %5.2f GFLOP/s %5.1f GB/s)\n", t, 2.0 * g.E() / t / 1e6,
         bytes_per_spmv(g.V(), g.E()) / t / 1e6);   // src/spmv/base.cu:69
}
