#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>
// histogram of (row, band) pair counts, weighted by entries: out[t] = entries in pairs with count >= thr[t]
void dense_stats(int64_t m, const uint64_t *rp, const int32_t *col, const int32_t *newid, const int32_t *perm, int64_t nrows,
                 int B, int band, int nthr, const int *thr, int64_t *out_entries, int64_t *out_pairs) {
#pragma omp parallel
  {
    uint32_t *cnt = calloc(B, 4);
    int64_t *le = calloc(nthr, 8), *lp = calloc(nthr, 8);
#pragma omp for schedule(dynamic, 256)
    for (int64_t j = 0; j < nrows; j++) {
      int32_t r = perm[j];
      memset(cnt, 0, 4 * B);
      for (uint64_t e = rp[r]; e < rp[r + 1]; e++) { int b = newid[col[e]] / band; if (b < B) cnt[b]++; }
      for (int b = 0; b < B; b++) for (int t = 0; t < nthr; t++) if (cnt[b] >= (uint32_t)thr[t]) { le[t] += cnt[b]; lp[t]++; }
    }
#pragma omp critical
    for (int t = 0; t < nthr; t++) { out_entries[t] += le[t]; out_pairs[t] += lp[t]; }
    free(cnt); free(le); free(lp);
  }
}
