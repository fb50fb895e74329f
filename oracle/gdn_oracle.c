/*
 * gdn_oracle.c -- CPU restatement of the Gardenia CSR-traversal hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see gdn_oracle.h).  Plain C; compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared
 * -ffp-contract=off matters: the reference is built as ISO C++
 * (`g++ -std=c++11`), which disables FMA contraction, so
 * `base + damp*sum` and `sum += x*a` round twice.
 *
 * Everything that the reference does in parallel with a result that does not
 * depend on the interleaving (row-independent gathers) may run under OpenMP
 * here; everything order-dependent is restated serially, i.e. this file is
 * the reference at OMP_NUM_THREADS=1.
 */
#include "gdn_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- 64-bit-word bitmap, include/bitmap.h:21-77 ------------------------- */
typedef struct { uint64_t *w; int64_t nwords; } bitmap_t;
static int bm_init(bitmap_t *b, int64_t n) {
  b->nwords = (n + 63) / 64;
  b->w = (uint64_t *)calloc((size_t)(b->nwords > 0 ? b->nwords : 1), sizeof(uint64_t));
  return b->w != NULL;
}
static void bm_reset(bitmap_t *b) { memset(b->w, 0, (size_t)b->nwords * sizeof(uint64_t)); }
static void bm_set(bitmap_t *b, int64_t pos) { b->w[pos >> 6] |= (uint64_t)1 << (pos & 63); }
static int bm_get(const bitmap_t *b, int64_t pos) { return (int)((b->w[pos >> 6] >> (pos & 63)) & 1); }
static void bm_swap(bitmap_t *a, bitmap_t *b) { bitmap_t t = *a; *a = *b; *b = t; }

/* ---- sliding queue, include/sliding_queue.h:28-77 ----------------------- */
typedef struct { int32_t *buf; int64_t in, out_start, out_end; } squeue_t;
static void sq_slide(squeue_t *q) { q->out_start = q->out_end; q->out_end = q->in; }
static int sq_empty(const squeue_t *q) { return q->out_start == q->out_end; }
static int64_t sq_size(const squeue_t *q) { return q->out_end - q->out_start; }

/* BUStep, src/bfs/omp_beamer.cc:13-32 */
static int64_t bu_step(int64_t m, const uint64_t *in_rowptr, const int32_t *in_colidx,
                       int32_t *depths, const bitmap_t *front, bitmap_t *next,
                       int64_t *edges_examined, int64_t *scanned) {
  int64_t awake = 0, ex = 0, sc = 0;
  bm_reset(next);
  for (int64_t dst = 0; dst < m; dst++) {
    if (depths[dst] < 0) {
      sc++;
      for (uint64_t e = in_rowptr[dst]; e < in_rowptr[dst + 1]; e++) {
        int32_t src = in_colidx[e];
        ex++;
        if (bm_get(front, src)) {
          depths[dst] = depths[src] + 1;
          awake++;
          bm_set(next, dst);
          break;
        }
      }
    }
  }
  *edges_examined = ex;
  *scanned = sc;
  return awake;
}

/* TDStep, src/bfs/omp_beamer.cc:35-58 (serial: the CAS always succeeds) */
static int64_t td_step(const uint64_t *out_rowptr, const int32_t *out_colidx,
                       int32_t *depths, squeue_t *q, int64_t *edges_examined,
                       int64_t *discovered) {
  int64_t scout = 0, ex = 0, nd = 0;
  for (int64_t i = q->out_start; i < q->out_end; i++) {
    int32_t src = q->buf[i];
    for (uint64_t e = out_rowptr[src]; e < out_rowptr[src + 1]; e++) {
      int32_t dst = out_colidx[e];
      int32_t cur = depths[dst];
      ex++;
      if (cur < 0) {
        depths[dst] = depths[src] + 1;
        q->buf[q->in++] = dst;
        scout += -(int64_t)cur;
        nd++;
      }
    }
  }
  *edges_examined = ex;
  *discovered = nd;
  return scout;
}

int oracle_bfs_do(int64_t m,
                  const uint64_t *out_rowptr, const int32_t *out_colidx,
                  const uint64_t *in_rowptr, const int32_t *in_colidx,
                  int32_t source, int32_t *dist,
                  oracle_bfs_step *steps, int max_steps, int *n_steps) {
  if (m <= 0 || source < 0 || source >= m) return -1;
  const int alpha = 15, beta = 18;                 /* omp_beamer.cc:111 */
  int32_t *depths = (int32_t *)malloc((size_t)m * sizeof(int32_t));
  squeue_t q;
  q.buf = (int32_t *)malloc((size_t)m * sizeof(int32_t));
  q.in = q.out_start = q.out_end = 0;
  bitmap_t curr, front;
  if (!depths || !q.buf || !bm_init(&curr, m) || !bm_init(&front, m)) return -1;
  /* InitDepth, omp_beamer.cc:89-95: -out_degree, or -1 for degree 0 */
  for (int64_t n = 0; n < m; n++) {
    int32_t deg = (int32_t)(out_rowptr[n + 1] - out_rowptr[n]);
    depths[n] = deg != 0 ? -deg : -1;
  }
  int64_t src_deg = (int64_t)(out_rowptr[source + 1] - out_rowptr[source]);
  depths[source] = 0;
  q.buf[q.in++] = source;
  sq_slide(&q);
  int64_t edges_to_check = (int64_t)out_rowptr[m];  /* g.E(), :129 */
  int64_t scout_count = src_deg;                    /* degrees[source], :130 */
  int iter = 0, ns = 0;
  while (!sq_empty(&q)) {
    if (scout_count > edges_to_check / alpha) {     /* :136 */
      int64_t awake, old_awake;
      for (int64_t i = q.out_start; i < q.out_end; i++) bm_set(&front, q.buf[i]);
      awake = sq_size(&q);
      sq_slide(&q);
      do {
        int64_t ex = 0, sc = 0;
        ++iter;
        old_awake = awake;
        awake = bu_step(m, in_rowptr, in_colidx, depths, &front, &curr, &ex, &sc);
        bm_swap(&front, &curr);
        if (steps && ns < max_steps) {
          oracle_bfs_step s = {1, 0, old_awake, ex, sc, awake, awake};
          steps[ns] = s;
        }
        ns++;
      } while ((awake >= old_awake) || (awake > m / beta));   /* :148-149 */
      /* BitmapToQueue, :69-79 */
      for (int64_t n = 0; n < m; n++)
        if (bm_get(&front, n)) q.buf[q.in++] = (int32_t)n;
      sq_slide(&q);
      scout_count = 1;                              /* :151 */
    } else {
      int64_t ex = 0, nd = 0, fsz = sq_size(&q);
      ++iter;
      edges_to_check -= scout_count;                /* :154 */
      scout_count = td_step(out_rowptr, out_colidx, depths, &q, &ex, &nd);
      sq_slide(&q);
      if (steps && ns < max_steps) {
        oracle_bfs_step s = {0, 0, fsz, ex, 0, nd, scout_count};
        steps[ns] = s;
      }
      ns++;
    }
  }
  for (int64_t i = 0; i < m; i++)                   /* :166-169 */
    dist[i] = depths[i] >= 0 ? depths[i] : ORACLE_INFINITY;
  if (n_steps) *n_steps = ns;
  free(depths); free(q.buf); free(curr.w); free(front.w);
  return iter;
}

int oracle_bfs_td(int64_t m, const uint64_t *rowptr, const int32_t *colidx,
                  int32_t source, int32_t *dist) {
  if (m <= 0 || source < 0 || source >= m) return -1;
  int32_t *queue = (int32_t *)malloc((size_t)m * sizeof(int32_t));
  if (!queue) return -1;
  for (int64_t i = 0; i < m; i++) dist[i] = ORACLE_INFINITY;   /* omp_base.cc:42 */
  dist[source] = 0;
  int64_t in = 0, start = 0, end;
  queue[in++] = source;
  end = in;
  int iter = 0;
  while (start != end) {                                       /* :52-57 */
    ++iter;
    for (int64_t i = start; i < end; i++) {                    /* bfs_step :11-31 */
      int32_t src = queue[i];
      for (uint64_t e = rowptr[src]; e < rowptr[src + 1]; e++) {
        int32_t dst = colidx[e];
        if (dist[dst] == ORACLE_INFINITY) {
          dist[dst] = dist[src] + 1;
          queue[in++] = dst;
        }
      }
    }
    start = end;
    end = in;
  }
  free(queue);
  return iter;
}

int64_t oracle_bfs_verify(int64_t m, const uint64_t *rowptr, const int32_t *colidx,
                          int32_t source, const int32_t *dist_to_test) {
  int32_t *depth = (int32_t *)malloc((size_t)m * sizeof(int32_t));
  if (!depth) return -1;
  oracle_bfs_td(m, rowptr, colidx, source, depth);   /* verifier.cc:14-27 is the same serial BFS */
  int64_t bad = 0;
  for (int64_t n = 0; n < m; n++)
    if (dist_to_test[n] != depth[n]) bad++;           /* verifier.cc:32-37 */
  free(depth);
  return bad;
}

int64_t oracle_bfs_check_parents(int64_t m, const uint64_t *in_rowptr, const int32_t *in_colidx,
                                 int32_t source, const int32_t *dist, const int32_t *parent) {
  int64_t bad = 0;
  for (int64_t v = 0; v < m; v++) {
    if (dist[v] == ORACLE_INFINITY) { if (parent[v] != -1) bad++; continue; }
    if (v == source) { if (parent[v] != source || dist[v] != 0) bad++; continue; }
    int32_t p = parent[v];
    if (p < 0 || p >= m || dist[p] != dist[v] - 1) { bad++; continue; }
    int found = 0;   /* (p -> v) must be an edge: p is in v's in-row */
    for (uint64_t e = in_rowptr[v]; e < in_rowptr[v + 1]; e++)
      if (in_colidx[e] == p) { found = 1; break; }
    if (!found) bad++;
  }
  return bad;
}

int oracle_pr_pull(int64_t m, const uint64_t *in_rowptr, const int32_t *in_colidx,
                   const int32_t *out_degree, float *scores,
                   float damp, double eps, int max_iter, double *err_trace) {
  const float base_score = (1.0f - damp) / (float)(int32_t)m;   /* omp_base.cc:16 */
  float *contrib = (float *)malloc((size_t)m * sizeof(float));
  float *delta = (float *)malloc((size_t)m * sizeof(float));
  if (!contrib || !delta) return -1;
  int iter;
  for (iter = 0; iter < max_iter; iter++) {                     /* :21 */
    double error = 0;
#pragma omp parallel for
    for (int64_t n = 0; n < m; n++)
      contrib[n] = scores[n] / (float)out_degree[n];            /* :24-25 */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t dst = 0; dst < m; dst++) {                     /* :27-33 */
      float incoming_total = 0;
      for (uint64_t e = in_rowptr[dst]; e < in_rowptr[dst + 1]; e++)
        incoming_total += contrib[in_colidx[e]];
      float old_score = scores[dst];
      float t = damp * incoming_total;
      scores[dst] = base_score + t;
      delta[dst] = fabsf(scores[dst] - old_score);
    }
    for (int64_t dst = 0; dst < m; dst++) error += (double)delta[dst];  /* :33, 1-thread order */
    if (err_trace) err_trace[iter] = error;
    if (error < eps) break;                                     /* :36 */
  }
  free(contrib); free(delta);
  return iter < max_iter ? iter + 1 : max_iter + 1;             /* printf("iterations = %d", iter+1) :38 */
}

double oracle_pr_residual(int64_t m, const uint64_t *out_rowptr, const int32_t *out_colidx,
                          const float *scores_to_test, float damp) {
  const float base_score = (1.0f - damp) / (float)(int32_t)m;   /* verifier.cc:11 */
  float *sums = (float *)calloc((size_t)m, sizeof(float));
  if (!sums) return -1.0;
  for (int64_t src = 0; src < m; src++) {                       /* :43-47 */
    int32_t deg = (int32_t)(out_rowptr[src + 1] - out_rowptr[src]);
    float oc = scores_to_test[src] / (float)deg;
    for (uint64_t e = out_rowptr[src]; e < out_rowptr[src + 1]; e++)
      sums[out_colidx[e]] += oc;
  }
  double error = 0;
  for (int64_t i = 0; i < m; i++) {                             /* :48-52 */
    float t = damp * sums[i];
    float new_score = base_score + t;
    error += (double)fabsf(new_score - scores_to_test[i]);
  }
  free(sums);
  return error;
}

void oracle_spmv(int64_t m, const uint64_t *Ap, const int32_t *Aj, const float *Ax,
                 const float *x, float *y) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < m; i++) {
    float sum = y[i];
    for (uint64_t jj = Ap[i]; jj < Ap[i + 1]; jj++) {
      float p = x[Aj[jj]] * Ax[jj];
      sum += p;
    }
    y[i] = sum;
  }
}

float oracle_max_relative_error(const float *a, const float *b, int64_t n) {
  float max_error = 0;
  const float eps = sqrtf(1.1920929e-07f);           /* sqrt(FLT_EPSILON), spmv_util.h:18 */
  for (int64_t i = 0; i < n; i++) {
    float error = fabsf(a[i] - b[i]);
    if (error != 0) {
      float r = error / (fabsf(a[i]) + fabsf(b[i]) + eps);
      if (r > max_error) max_error = r;
    }
  }
  return max_error;
}
