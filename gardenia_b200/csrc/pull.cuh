// pull.cuh -- what pull.cu (SELL PageRank iteration) and band.cu (banded shared-memory part of it) share.
#pragma once
#include "common.cuh"

namespace gdn {

constexpr int kHotMax = 49152;          // fp32 entries in the shared-memory table (192 KB)
constexpr int kGroupCh = 1024;          // int4 groups per work item (= 4096 column ids)
constexpr int kSellThreads = 1024;      // one CTA per SM
constexpr int kExactCols = 262144;       // slices wider than this keep the reference's summation order (pull.cu pull_prepare)
constexpr int kExactColsStrict = 8192;   // ... in exact-order mode, where narrower slices are whole items of one warp
constexpr int kMaxPeers = 7;            // other GPUs of one box whose vectors a row epilogue writes

struct SellArgs {
  const int4 *sell;
  const uint32_t *slice_ptr;
  const int32_t *chunk_slice;
  int32_t n_chunks;
  const int2 *heavy_seg;
  const int32_t *heavy_slice, *heavy_first;
  int32_t n_heavy_segs, n_heavy_slices;
  float *partial;
  const float *contrib_in;     // indexed by NEW global id, length Mp
  float *contrib_out;
  float *scores;               // sorted local order
  const int32_t *sdeg;         // row length of sorted row j
  const int32_t *sout;         // out-degree of sorted row j (nullptr: = sdeg)
  const int32_t *rowid;        // new global id of sorted row j (nullptr: formula below)
  int64_t n_nz_rows, rows;
  int32_t H;                   // hot ids of the id space (row_newid)
  int32_t hot_n;               // entries of the shared-memory table (<= H)
  int64_t Hp, Wc;
  int32_t rank;
  float base, damp;
  double *err_partial;
  const int32_t *done;
  int32_t err_slot0;
  int32_t warm;                // ids whose tier_id is below this are kept L2-resident; colder ids are gathered evict-first
  int32_t P;                   // ranks of the row partition the id space was laid out for
  float inv_wc;                // 1 / Wc
  uint32_t group_ch;           // int4 groups per work item; slices wider than this are cut into segments (0xffffffff: never, exact order)
  int32_t n_exact;             // the first n_exact slices are exact slices: one item each, never cut, never banded
  int32_t *work_counter;       // batches of work items drawn so far by the warps of this launch
  int32_t pf_trips;            // index groups of a slice requested into L2 ahead of the trip being cut (0 = none; GDN_PR_MAIN_PF)
  int32_t strict_chain;        // exact-order mode: exact slices by a true sequential chain (bit-exact) instead of the ordered-sum emulation
  const float4 *exact_vals;    // gathered values of the exact slices (pr_exact_gather), laid out like their index groups
  // banded layout (band.cu): sorted rows below n_band_rows (a multiple of 32) only deposit the sum over the columns left
  // in the main array; pr_band_finalize adds their band partials and runs the row epilogue
  int64_t n_band_rows;
  float *acc_main;
  // row partition with peer-mapped vectors: contrib_out / the other buffer of the n_peers OTHER GPUs
  int32_t n_peers;
  float *peer_out[kMaxPeers];
  float *peer_other[kMaxPeers];
};

__device__ __forceinline__ int64_t row_newid(const SellArgs &a, int64_t j) {
  if (a.rowid) return a.rowid[j];
  return j < a.Hp ? (int64_t)a.rank * a.Hp + j : (int64_t)a.H + (int64_t)a.rank * a.Wc + (j - a.Hp);
}

// Hotness rank of a new id for the L2 residency tiers.  One GPU: the id itself (ids are sorted hottest first).  Row
// partition: the id space is [hot prefix | cold slice of rank 0 | ... ], every slice hottest first, so the rank is the
// position INSIDE the slice.  The float quotient may be off by one at a slice boundary: that only mislabels the cache
// policy of a few ids, never an address.
__device__ __forceinline__ int32_t tier_id(const SellArgs &a, int32_t c) {
  if (a.P > 1 && c >= a.H) {
    const uint32_t cw = (uint32_t)(c - a.H);
    const uint32_t q = (uint32_t)__fmul_rz((float)cw, a.inv_wc);
    return a.H + (int32_t)(cw - q * (uint32_t)a.Wc);
  }
  return c;
}

// New contrib of the row that owns new id `id`: into this GPU's vector and, in a row partition with peer-mapped
// vectors (comm.cu pull_peer_setup), straight into every other GPU's copy over NVLink -- the exchange of SURVEY 8(e)
// fused into the row epilogue, no collective afterwards.  Lanes of a warp own consecutive ids: 128-byte stores.
__device__ __forceinline__ void contrib_store(const SellArgs &a, int64_t id, float cv, bool warm) {
  if (warm) a.contrib_out[id] = cv; else __stcs(a.contrib_out + id, cv);
#pragma unroll
  for (int p = 0; p < kMaxPeers; p++)
    if (p < a.n_peers) a.peer_out[p][id] = cv;
}
// (rows settled once per solve also seed the OTHER buffer of the double-buffered vector, on every GPU)
__device__ __forceinline__ void contrib_store_other(const SellArgs &a, float *contrib_other, int64_t id, float cv) {
  __stcs(contrib_other + id, cv);
#pragma unroll
  for (int p = 0; p < kMaxPeers; p++)
    if (p < a.n_peers) a.peer_other[p][id] = cv;
}

// Last statement of every kernel that may have stored into the peers' vectors: each thread makes ITS remote stores
// performed system-wide before the kernel can end, so that the flag the barrier kernel raises afterwards (comm.cu
// peer_sync_kernel) can never overtake them on the way to another GPU.
__device__ __forceinline__ void contrib_flush(const SellArgs &a) {
  if (a.n_peers > 0) __threadfence_system();
}

// scores[dst] = base + damp * sum; error += |new - old|; next contrib   (src/pr/omp_base.cc:24-25,31-33)
__device__ __forceinline__ void pr_epilogue_pre(const SellArgs &a, int64_t j, float acc, double &err, float old_score, int32_t deg) {
  // scores / sdeg are touched once per iteration: streaming (evict-first) accesses keep them from
  // displacing the warm part of contrib in L2; the new contrib of a warm id is stored normally (it is
  // gathered in the next iteration), a cold one streaming.
  const float nw = __fadd_rn(a.base, __fmul_rn(a.damp, acc));
  __stcs(a.scores + j, nw);
  err += (double)fabsf(__fsub_rn(nw, old_score));
  const int64_t id = row_newid(a, j);
  const float cv = __fdiv_rn(nw, (float)deg);
  contrib_store(a, id, cv, tier_id(a, (int32_t)id) < a.warm);
}
__device__ __forceinline__ void pr_epilogue_core(const SellArgs &a, int64_t j, float acc, double &err) {
  const float old_score = __ldcs(a.scores + j);
  const int32_t deg = a.sout ? __ldcs(a.sout + j) : __ldcs(a.sdeg + j);
  pr_epilogue_pre(a, j, acc, err, old_score, deg);
}
__device__ __forceinline__ void pr_epilogue(const SellArgs &a, int64_t j, float acc, double &err) {
  if (j < a.n_band_rows) { a.acc_main[j] = acc; return; }
  pr_epilogue_core(a, j, acc, err);
}

// band.cu
int band_build(gdn_graph *g);
void band_free(BandLayout &b);
int band_launch(gdn_graph *g, const SellArgs &a, double fix_scale, cudaStream_t s);    // the band partial sums of one iteration
int band_finalize_launch(gdn_graph *g, const SellArgs &a, double fix_scale, int grid, cudaStream_t s);    // + main sums -> row epilogue
int band_finalize_grid(const gdn_graph *g);
int band_launches(const gdn_graph *g);
int band_solve_begin(gdn_graph *g, cudaStream_t s);

}  // namespace gdn
