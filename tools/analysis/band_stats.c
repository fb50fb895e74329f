// offline analysis: what a banded shared-memory layout would move (not product code)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>
#define CAP 65536
// out[b*4+0]=pairs, +1 = moved edges, +2 = padded u16 entries (per-band sort, slices of 32, width ceil8), +3 = padded without sort
void band_stats(int64_t m, const uint64_t *rp, const int32_t *col, const int32_t *newid, const int32_t *perm,
                int B, int band, int cmin, int64_t *out, int64_t *main_pad /* [2]: remaining entries, padded remaining entries (ceil4, slices in perm order) */) {
  int T = omp_get_max_threads();
  uint32_t **hist = malloc(sizeof(void *) * T);
  int64_t *nosort = calloc((size_t)B, 8);
  int64_t rem_tot = 0, rem_pad = 0;
  int64_t nsl = (m + 31) / 32;
#pragma omp parallel reduction(+ : rem_tot, rem_pad)
  {
    int t = omp_get_thread_num();
    hist[t] = calloc((size_t)B * CAP, 4);
    uint32_t *cnt = malloc(4 * 32 * (size_t)B);
    int64_t *ns = calloc((size_t)B, 8);
#pragma omp for schedule(dynamic, 64)
    for (int64_t s = 0; s < nsl; s++) {
      memset(cnt, 0, 4 * 32 * (size_t)B);
      uint32_t remmax = 0;
      for (int l = 0; l < 32; l++) {
        int64_t j = s * 32 + l;
        if (j >= m) break;
        int32_t r = perm[j];
        uint32_t rem = 0;
        for (uint64_t e = rp[r]; e < rp[r + 1]; e++) {
          int32_t c = newid[col[e]];
          int b = c / band;
          if (b < B) cnt[b * 32 + l]++; else rem++;
        }
        for (int b = 0; b < B; b++) if (cnt[b * 32 + l] < (uint32_t)cmin) { rem += cnt[b * 32 + l]; cnt[b * 32 + l] = 0; }
        rem_tot += rem;
        if (rem > remmax) remmax = rem;
      }
      rem_pad += 32ll * ((remmax + 3) / 4 * 4);
      for (int b = 0; b < B; b++) {
        uint32_t mx = 0;
        for (int l = 0; l < 32; l++) { uint32_t c = cnt[b * 32 + l]; if (c) { hist[t][(size_t)b * CAP + (c < CAP ? c : CAP - 1)]++; if (c > mx) mx = c; } }
        ns[b] += 32ll * ((mx + 7) / 8 * 8);
      }
    }
#pragma omp critical
    for (int b = 0; b < B; b++) nosort[b] += ns[b];
    free(cnt); free(ns);
  }
  for (int b = 0; b < B; b++) {
    int64_t pairs = 0, moved = 0, padded = 0, inslice = 0; uint32_t wmax = 0;
    for (int c = CAP - 1; c >= 1; c--) {
      int64_t n = 0;
      for (int t = 0; t < T; t++) n += hist[t][(size_t)b * CAP + c];
      pairs += n; moved += n * c;
      while (n > 0) {
        if (inslice == 0) wmax = (c + 7) / 8 * 8;
        int64_t take = 32 - inslice < n ? 32 - inslice : n;
        // full slices of this count
        inslice += take; n -= take;
        if (inslice == 32) { padded += 32ll * wmax; inslice = 0; if (n >= 32) { int64_t full = n / 32; padded += full * 32ll * ((c + 7) / 8 * 8); n -= full * 32; } }
      }
    }
    if (inslice) padded += 32ll * wmax;
    out[b * 4 + 0] = pairs; out[b * 4 + 1] = moved; out[b * 4 + 2] = padded; out[b * 4 + 3] = nosort[b];
  }
  main_pad[0] = rem_tot; main_pad[1] = rem_pad;
  for (int t = 0; t < T; t++) free(hist[t]);
  free(hist); free(nosort);
}
