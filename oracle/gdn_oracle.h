/*
 * gdn_oracle.h -- CPU restatement of the Gardenia CSR-traversal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gardenia_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg use it, and only as the checker.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against
 *   - the reference's golden PR trace  test/reference/graph-pr.mtx.out:13-28
 *   - outputs of the reference's own OpenMP sources compiled from
 *     /root/reference by oracle/Makefile (oracle/_ref/ref_driver), committed
 *     as fixtures under tests/golden/ by tools/make_golden.py.
 *
 * All file:line citations are relative to the reference tree root.
 */
#ifndef GDN_ORACLE_H_
#define GDN_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_INFINITY 1000000000 /* MYINFINITY, include/common.h:66 */

/* One entry per BFS step taken by the direction-optimizing controller. */
typedef struct {
  int32_t dir;          /* 0 = top-down step, 1 = bottom-up step */
  int32_t pad;
  int64_t frontier;     /* |frontier| expanded by this step */
  int64_t edges;        /* TD: sum of out-degree over the frontier; BU: in-edges examined until first hit */
  int64_t scanned;      /* BU: number of unvisited vertices scanned (0 for TD) */
  int64_t discovered;   /* vertices that received a depth in this step */
  int64_t scout;        /* TD: scout_count returned; BU: awake_count returned */
} oracle_bfs_step;

/* src/bfs/omp_beamer.cc:97-171 (BFSSolver), :35-58 (TDStep), :13-32 (BUStep),
 * :60-79 (QueueToBitmap / BitmapToQueue), :89-95 (InitDepth).
 * Returns the reference's `iterations` counter, or -1 on bad arguments.
 * steps may be NULL; at most max_steps entries are written, *n_steps gets the
 * number of steps taken. */
int oracle_bfs_do(int64_t m,
                  const uint64_t *out_rowptr, const int32_t *out_colidx,
                  const uint64_t *in_rowptr, const int32_t *in_colidx,
                  int32_t source, int32_t *dist,
                  oracle_bfs_step *steps, int max_steps, int *n_steps);

/* src/bfs/omp_base.cc:11-65 -- level-synchronous top-down BFS. Returns `iterations`. */
int oracle_bfs_td(int64_t m, const uint64_t *rowptr, const int32_t *colidx,
                  int32_t source, int32_t *dist);

/* src/bfs/verifier.cc:8-40 -- serial queue BFS + exact compare.
 * Returns the number of mismatching vertices (0 == "Correct"). */
int64_t oracle_bfs_verify(int64_t m, const uint64_t *rowptr, const int32_t *colidx,
                          int32_t source, const int32_t *dist_to_test);

/* Graph500-style parent-tree check (extension; the reference keeps parents
 * only in comments, src/bfs/omp_beamer.cc:12,18,22,44,47).
 * Returns number of violations. */
int64_t oracle_bfs_check_parents(int64_t m, const uint64_t *in_rowptr, const int32_t *in_colidx,
                                 int32_t source, const int32_t *dist, const int32_t *parent);

/* src/pr/omp_base.cc:8-42 -- Jacobi pull PageRank.
 * scores is in/out (pre-filled by the caller, src/pr/main.cc:17-18).
 * err_trace (nullable, max_iter doubles) receives the per-iteration L1 delta
 * printed by the reference.  Returns the reference's `iterations` (iter+1). */
int oracle_pr_pull(int64_t m, const uint64_t *in_rowptr, const int32_t *in_colidx,
                   const int32_t *out_degree, float *scores,
                   float damp, double eps, int max_iter, double *err_trace);

/* src/pr/verifier.cc:40-54 -- one push iteration from scores_to_test; returns
 * the L1 residual that the reference compares against target_error. */
double oracle_pr_residual(int64_t m, const uint64_t *out_rowptr, const int32_t *out_colidx,
                          const float *scores_to_test, float damp);

/* src/spmv/omp_base.cc:22-33 == src/spmv/spmv_util.h:31-42 -- y += A*x, fp32,
 * sequential per row. */
void oracle_spmv(int64_t m, const uint64_t *Ap, const int32_t *Aj, const float *Ax,
                 const float *x, float *y);

/* src/spmv/spmv_util.h:15-29 -- maximum symmetric relative error. */
float oracle_max_relative_error(const float *a, const float *b, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
