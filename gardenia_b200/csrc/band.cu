// band.cu -- banded shared-memory gather for the heavy rows of the PageRank pull layout.
//
// Why (profiles/r1_pr_tier_probe.txt, r1_ncu_pr_pipe_kron26.txt): pr_sell_pipe sits on the L1TEX miss path -- one
// 32-byte sector per clock and SM for every contrib[src] (src/pr/omp_base.cc:29-30) that is not in the 48 K-entry
// shared-memory table, 1.5 G sectors per iteration at Kronecker scale 26 -- while a gather out of shared memory costs
// a fifth of that.  Only 28 % of the column ids fall into the ONE table every SM holds; 73 % fall into the hottest
// 3 M ids.  So the id space is cut into bands of 48 K ids and each band's table is held by the CTAs that own it:
//
//   * for every sorted row of length >= dmin and every band b in which the row has >= cmin column ids, those ids
//     leave the row's main SELL slice and are stored as 16-bit band-local ids in band b's own SELL-32 array, whose
//     rows are sorted by their count IN THAT BAND (13 % padding instead of 85 %);
//   * pr_band_kernel: a CTA loads contrib[band b] into shared memory once and streams a contiguous run of band b's
//     index groups (8 ids per 128-bit load, lane = row, sequential fp32 sum per row in column order); one partial
//     sum per (band slice segment, row);
//   * what is left (cold ids, sparse (row, band) pairs) is a compacted main SELL array that the unchanged
//     pr_sell_pipe walks; its epilogue for these rows only deposits the main sum;
//   * the band partials of a row meet in a 64-bit FIXED-POINT accumulator (integer atomics: order-free, exact, hence
//     bit-reproducible from run to run); pr_band_finalize_fix adds main sum + band sum and runs the row epilogue
//     (score, L1 delta, next contrib).  The scale 2^e of the accumulators follows sum |scores_0| of the solve (pull.cu
//     fix_scale_for), so any caller-supplied start vector stays in range.
//
// The layout is built once per resident graph (untimed, like include/segmenting.h preprocessing of the reference):
// two device passes over the existing SELL array (count, fill) around a host pass that sorts each band's rows.
// Every rank of a row partition bands ITS rows over the shared id space (band_of); one-shot calls keep the plain layout.
//
// SEGMENTED mode (graphs WITHOUT a hot set whose gathered vector does not fit L2, e.g. uniform-random scale >= 25: every
// gather of the plain layout is then an HBM sector miss, 33 ms per iteration at urand-26): the same machinery with
// bands the size of an L2-resident slice (40 MB), every row taking part with every id, 32-bit ids, one launch per band
// (pr_seg_kernel) so that only ONE slice of contrib is hot in L2 at a time.  The gathers become L2 hits; the per-row
// partials of the passes meet in the fixed-point accumulators.
#include "pull.cuh"
#include <omp.h>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace gdn {

constexpr int kBandSeg = 64;            // index groups (8 ids each) per lane and item: 512 ids of a row
constexpr int kBandTab = kHotMax + 256; // table entries in shared memory: [band, kBandTab) is zero
constexpr uint32_t kBandPadId = 0xC0C0; // padding id = a zero table entry; byte-uniform so cudaMemset can write it
constexpr uint32_t kNone = 0xffffffffu;
static_assert(kBandPadId >= (uint32_t)kHotMax && kBandPadId < (uint32_t)kBandTab, "padding id must hit the zero tail");

// Which band holds new id c.  The id space is [hot prefix of H ids | cold slice of rank 0 | ... | cold slice of rank
// P-1] (pull.cu), every slice sorted hottest first: the hot prefix is cut into n0 bands, then the cold slices are cut
// into pieces of `band` ids that are taken round-robin over the ranks, so that a low band index always means hot ids.
// On one GPU this is simply b = c / band when band divides H.
struct BandMap {
  int32_t band, n0, P, B;
  int64_t H, Wc;
};
__host__ __device__ __forceinline__ uint32_t band_of(const BandMap &mp, int64_t c, uint32_t &local) {
  if (c < mp.H || mp.Wc <= 0) { const uint32_t b = (uint32_t)(c / mp.band); local = (uint32_t)(c - (int64_t)b * mp.band); return b; }
  const int64_t off = c - mp.H, q = off / mp.Wc, r = off - q * mp.Wc, t = r / mp.band;
  local = (uint32_t)(r - t * mp.band);
  const int64_t b = mp.n0 + t * mp.P + q;
  return b < mp.B ? (uint32_t)b : 0xffffffffu;
}
static void band_range(const BandMap &mp, int b, int64_t &start, int32_t &len) {
  if (b < mp.n0) { start = (int64_t)b * mp.band; len = (int32_t)std::min<int64_t>(mp.band, mp.H - start); return; }
  const int64_t k = b - mp.n0, t = k / mp.P, q = k % mp.P;
  start = mp.H + q * mp.Wc + t * mp.band;
  len = (int32_t)std::max<int64_t>(0, std::min<int64_t>(mp.band, mp.Wc - t * mp.band));
}

// ------------------------------------------------------------------ build pass 1: count[b][j]
// One warp per slice of the existing SELL array (lane = row): how many ids of row j fall into band b.
__global__ void __launch_bounds__(128)
band_count(const int4 *__restrict__ sell, const uint32_t *__restrict__ slice_ptr, int32_t nb_slices, BandMap mp,
           uint32_t *__restrict__ cnt, int64_t n_rows) {
  const int B = mp.B;
  extern __shared__ uint32_t s_u32[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *c = s_u32 + (size_t)wib * B * 32;
  for (int32_t s = blockIdx.x * 4 + wib; s < nb_slices; s += gridDim.x * 4) {
    for (int b = 0; b < B; b++) c[b * 32 + lane] = 0;
    const uint32_t g0 = slice_ptr[s], g1 = slice_ptr[s + 1];
    const int4 *p = sell + g0 + lane;
    const uint32_t n = (g1 - g0) >> 5;
    for (uint32_t k = 0; k < n; k += 4) {
      int4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) q[u] = k + u < n ? p[(size_t)(k + u) * 32] : make_int4(-1, -1, -1, -1);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (v[t] >= 0) { uint32_t loc; const uint32_t b = band_of(mp, v[t], loc); if (b < (uint32_t)B) c[b * 32 + lane]++; }
      }
    }
    for (int b = 0; b < B; b++) cnt[(size_t)b * n_rows + (size_t)s * 32 + lane] = c[b * 32 + lane];
  }
}

// (row, band) pairs with fewer than cmin ids stay in the main array; rem_w[s] = widest remainder of slice s.
__global__ void __launch_bounds__(256)
band_select(uint32_t *__restrict__ cnt, int B, int64_t n_rows, uint32_t cmin, const int32_t *__restrict__ sdeg,
            uint32_t *__restrict__ rem_w, int64_t n_exact_rows) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_rows; j += (int64_t)gridDim.x * blockDim.x) {
    uint32_t tot = 0;
    const uint32_t need = j < n_exact_rows ? 0xffffffffu : cmin;       // rows of exact slices keep every id in the main array
    for (int b = 0; b < B; b++) {
      const uint32_t v = cnt[(size_t)b * n_rows + j];
      if (v < need) { if (v) cnt[(size_t)b * n_rows + j] = 0; } else tot += v;
    }
    uint32_t rem = (uint32_t)sdeg[j] - tot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rem = max(rem, __shfl_xor_sync(kFull, rem, o));
    if ((threadIdx.x & 31) == 0) rem_w[j >> 5] = rem;       // n_rows is a multiple of 32 and j is warp-aligned
  }
}

// ------------------------------------------------------------------ build pass 2: fill both arrays
// One warp per slice again.  rank[b][j] = position of row j among band b's rows (kNone: the pair stays in the main
// array).  A lane walks ITS row in column order, so the order of a row's ids inside a band section and inside the
// compacted main slice is the order they had before.
template <bool WIDE>
__global__ void __launch_bounds__(128)
band_fill(const int4 *__restrict__ sell, const uint32_t *__restrict__ slice_ptr, int32_t nb_slices, BandMap mp,
          const uint32_t *__restrict__ rank, int64_t n_rows, const uint32_t *__restrict__ bslice_ptr,
          const int32_t *__restrict__ bslice_first, uint16_t *__restrict__ bsell16, int4 *__restrict__ sell2,
          const uint32_t *__restrict__ slice_ptr2) {
  extern __shared__ uint32_t s_u32[];
  const int B = mp.B;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t *base = s_u32 + (size_t)wib * 2 * B * 32;     // unit of group 0 of this row in band b (kNone: not there)
  uint32_t *kcnt = base + (size_t)B * 32;
  for (int32_t s = blockIdx.x * 4 + wib; s < nb_slices; s += gridDim.x * 4) {
    const int64_t j = (int64_t)s * 32 + lane;
    for (int b = 0; b < B; b++) {
      const uint32_t r = rank[(size_t)b * n_rows + j];
      base[b * 32 + lane] = r == kNone ? kNone : bslice_ptr[bslice_first[b] + (r >> 5)] + (r & 31);
      kcnt[b * 32 + lane] = 0;
    }
    const uint32_t g0 = slice_ptr[s], g1 = slice_ptr[s + 1];
    const int4 *p = sell + g0 + lane;
    const uint32_t n = (g1 - g0) >> 5;
    int4 *dst = sell2 + slice_ptr2[s] + lane;
    int4 pend = make_int4(-1, -1, -1, -1);
    int np = 0;
    for (uint32_t k = 0; k < n; k += 4) {
      int4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) q[u] = k + u < n ? p[(size_t)(k + u) * 32] : make_int4(-1, -1, -1, -1);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int c = v[t];
          if (c < 0) continue;
          uint32_t loc;
          const uint32_t b = band_of(mp, c, loc);
          uint32_t bs = kNone;
          if (b < (uint32_t)B) bs = base[b * 32 + lane];
          if (bs != kNone) {
            const uint32_t kk = kcnt[b * 32 + lane]++;
            if (WIDE) reinterpret_cast<int32_t *>(bsell16)[((size_t)bs + (size_t)(kk >> 2) * 32) * 4 + (kk & 3)] = c;   // 4 global ids per unit
            else bsell16[((size_t)bs + (size_t)(kk >> 3) * 32) * 8 + (kk & 7)] = (uint16_t)loc;
          } else {
            if (np == 0) pend.x = c; else if (np == 1) pend.y = c; else if (np == 2) pend.z = c; else pend.w = c;
            if (++np == 4) { *dst = pend; dst += 32; pend = make_int4(-1, -1, -1, -1); np = 0; }
          }
        }
      }
    }
    if (np) *dst = pend;
  }
}

// ------------------------------------------------------------------ build pass 3: spread a row's ids over the banks
// pr_band_kernel reads tab[id] for 32 ROWS at a time (lane = row, same position in every row): 32 random words of a
// 192 KB table fall into the 32 banks like balls into bins -- 3.2 wavefronts per LDS measured (ncu: 113 M of 164 M
// shared wavefronts are bank conflicts, profiles/r2_ncu_pr.txt).  The order of a row's ids INSIDE an item is free (the
// item's partial is a band-local sum already), so one warp per item re-orders every row (lane) in chunks of 64 positions:
// at position p lane l prefers bank (l + p) mod 32 -- a Latin square, conflict-free by construction -- and takes the
// nearest bank it still has an id in; lanes that collide are settled lowest-lane-first and the losers look for a bank
// nobody holds (a few match_any rounds).  Sums change by a re-ordering of <= 512 fp32 addends per (row, band, item).
// stats[0], stats[1] = bank-conflict degree (max lanes on one bank) summed over all positions before / after.
constexpr int kSpreadPos = 64;                               // positions (ids per row) per chunk
__global__ void __launch_bounds__(256)
band_spread(uint16_t *__restrict__ bsell16, const uint32_t *__restrict__ item_ptr, int32_t n_items, unsigned long long *stats) {
  __shared__ uint16_t s_ids[8][kSpreadPos * 32];             // [position][lane]: the chunk's ids of a lane, sorted by bank
  __shared__ uint8_t s_ptr[8][32 * 32], s_rem[8][32 * 32];   // [bank][lane]: next unused position of the bank / ids left in it
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint16_t *ids = s_ids[wib];
  uint8_t *ptr = s_ptr[wib], *rem = s_rem[wib];
  unsigned long long before = 0, after = 0;                  // (warp-uniform)
  auto degree = [&](uint32_t bank, bool real) -> uint32_t {  // most lanes on one bank at this position
    const unsigned peers = __match_any_sync(kFull, real ? bank : 64u + lane);
    return __reduce_max_sync(kFull, real ? (uint32_t)__popc(peers) : 0u);
  };
  for (int32_t it = (blockIdx.x * 256 + threadIdx.x) >> 5; it < n_items; it += (gridDim.x * 256) >> 5) {
    const uint32_t u0 = item_ptr[it], ng = (item_ptr[it + 1] - u0) >> 5;           // index groups (8 ids) per lane
    for (uint32_t g0 = 0; g0 < ng; g0 += kSpreadPos / 8) {
      const uint32_t cg = min((uint32_t)(kSpreadPos / 8), ng - g0), np = cg * 8;
      uint16_t *base = bsell16 + ((size_t)u0 + (size_t)g0 * 32 + lane) * 8;        // group g of this lane: base + g * 256
      // counting sort of the lane's ids by bank (everything of a lane lives in its own shared-memory column)
      for (int b = 0; b < 32; b++) rem[b * 32 + lane] = 0;
      for (uint32_t g = 0; g < cg; g++) {
        const uint4 q = *reinterpret_cast<const uint4 *>(base + (size_t)g * 256);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint32_t id = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
          const bool real = id != kBandPadId;
          before += degree(id & 31u, real);
          if (real) rem[(id & 31u) * 32 + lane]++;
        }
      }
      uint32_t avail = 0, left = 0;
      for (int b = 0; b < 32; b++) {
        const uint32_t c = rem[b * 32 + lane];
        if (c) avail |= 1u << b;
        ptr[b * 32 + lane] = (uint8_t)left;
        left += c;
      }
      for (uint32_t g = 0; g < cg; g++) {
        const uint4 q = *reinterpret_cast<const uint4 *>(base + (size_t)g * 256);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint32_t id = (w[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
          if (id != kBandPadId) ids[(uint32_t)(ptr[(id & 31u) * 32 + lane]++) * 32 + lane] = (uint16_t)id;
        }
      }
      for (int b = 0; b < 32; b++) ptr[b * 32 + lane] -= rem[b * 32 + lane];      // back to the start of every bank
      uint32_t out[4] = {0, 0, 0, 0};
      for (uint32_t p = 0; p < np; p++) {
        const bool real = left > 0;
        const uint32_t pref = (lane + p) & 31u;
        auto nearest = [&](uint32_t m) -> uint32_t {         // first bank of mask m at or after pref, cyclically
          const uint32_t r = __funnelshift_r(m, m, pref);
          return ((uint32_t)(__ffs(r) - 1) + pref) & 31u;
        };
        uint32_t want = real ? nearest(avail) : 0u;
        bool fixed = !real;
        for (int round = 0; round < 4; round++) {
          const unsigned fixed_mask = __ballot_sync(kFull, fixed && real);
          const unsigned peers = __match_any_sync(kFull, real ? want : 64u + lane);
          if (real && !fixed && !(peers & fixed_mask) && lane == __ffs(peers) - 1) fixed = true;    // lowest lane of a free bank wins
          const uint32_t held = __reduce_or_sync(kFull, fixed && real ? 1u << want : 0u);
          if (__all_sync(kFull, fixed)) break;
          if (!fixed) { const uint32_t alt = avail & ~held; if (alt) want = nearest(alt); }
        }
        after += degree(want, real);
        uint32_t id = kBandPadId;
        if (real) {
          const uint32_t pos = ptr[want * 32 + lane];
          id = ids[pos * 32 + lane];
          ptr[want * 32 + lane] = (uint8_t)(pos + 1);
          if (--rem[want * 32 + lane] == 0) avail &= ~(1u << want);
          left--;
        }
        out[(p & 7) >> 1] |= id << ((p & 1) * 16);
        if ((p & 7) == 7) {
          *reinterpret_cast<uint4 *>(base + (size_t)(p >> 3) * 256) = make_uint4(out[0], out[1], out[2], out[3]);
          out[0] = out[1] = out[2] = out[3] = 0;
        }
      }
    }
  }
  if (lane == 0 && stats) { atomicAdd(stats, before); atomicAdd(stats + 1, after); }
}

// ------------------------------------------------------------------ the iteration: band partial sums
struct BandArgs {
  const uint4 *bsell;
  const uint32_t *item_ptr;
  const int4 *job;
  const int32_t *job_first;
  const int32_t *wrun;
  const int64_t *band_start;   // first new id of band b ...
  const int32_t *band_len;     // ... and how many ids it holds (<= kHotMax)
  const float *contrib_in;
  const int32_t *irow;         // sorted row of (item, lane), -1 = no row ...
  unsigned long long *acc_fix; // ... whose fixed-point accumulator takes the partial
  double fix_scale;            // 2^e of the accumulators for this solve
  const int32_t *done;
  int32_t pf_groups;           // index groups requested into L2 ahead of the loads of a warp (0 = no prefetch)
};

__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}

// PD index groups (8 ids each) requested ahead per lane.  A warp owns a contiguous run of items = one contiguous
// piece of the band's array: it streams through it without a bubble at item boundaries; the item lengths come 32 at a
// time (lane i holds the length of item ibase + i) with the next batch requested one batch ahead.
template <int PD>
__global__ void __launch_bounds__(kSellThreads, 1)
pr_band_kernel(BandArgs a) {
  extern __shared__ float tab[];
  if (*a.done) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t pol = l2_policy_evict_first();
  const uint4 padq = make_uint4(kBandPadId * 0x10001u, kBandPadId * 0x10001u, kBandPadId * 0x10001u, kBandPadId * 0x10001u);
  for (int32_t jb = a.job_first[blockIdx.x]; jb < a.job_first[blockIdx.x + 1]; jb++) {
    const int4 J = a.job[jb];
    __syncthreads();                                       // the previous band's readers are done
    const int64_t id0 = a.band_start[J.x];
    const int32_t len = a.band_len[J.x];
    for (int i = threadIdx.x; i < kBandTab; i += kSellThreads) tab[i] = i < len ? a.contrib_in[id0 + i] : 0.f;
    __syncthreads();
    const int32_t i0 = a.wrun[J.y + w], i1 = a.wrun[J.y + w + 1];
    if (i0 >= i1) continue;
    const uint32_t u0 = a.item_ptr[i0], u1 = a.item_ptr[i1];
    const uint4 *p = a.bsell + u0 + lane;
    const uint32_t nrows = (u1 - u0) >> 5;
    auto len_batch = [&](int32_t ib) -> uint32_t {
      const int32_t i = ib + lane;
      return i < i1 ? (a.item_ptr[i + 1] - a.item_ptr[i]) >> 5 : 0u;
    };
    int32_t ibase = i0, item = i0;
    uint32_t lens = len_batch(ibase), lens_next = len_batch(ibase + 32);
    uint32_t left = __shfl_sync(kFull, lens, 0);
    // the rows of the next two items are requested ahead (an item is ~4 index groups long)
    int32_t jrow = a.irow[(size_t)i0 * 32 + lane], jrow1 = -1, jrow2 = -1;
    if (i0 + 1 < i1) jrow1 = a.irow[(size_t)(i0 + 1) * 32 + lane];
    if (i0 + 2 < i1) jrow2 = a.irow[(size_t)(i0 + 2) * 32 + lane];
    // the warp's run is one contiguous piece of the band's array: ask L2 for it a few KB ahead of the register loads (one
    // bulk prefetch per 8 index groups = 4 KB, issued by lane 0), so that the PD loads in flight per lane wait for L2,
    // not for HBM -- the kernel was latency-bound on exactly these loads (ncu: long scoreboard, DRAM at 39 %)
    const uint32_t pf = (uint32_t)a.pf_groups;
    if (pf && lane == 0) bulk_prefetch_l2(p - lane, min(pf, nrows) * 512u);
    uint4 q[PD];
#pragma unroll
    for (int d = 0; d < PD; d++) q[d] = (uint32_t)d < nrows ? ld_stream_u4(p + 32 * d, pol) : padq;
    float acc = 0.f;
    for (uint32_t r = 0; r < nrows; r += PD) {
      if (pf && (r & 7u) == 0 && lane == 0 && r + pf < nrows) bulk_prefetch_l2(p - lane + (size_t)(r + pf) * 32, min(8u, nrows - (r + pf)) * 512u);
#pragma unroll
      for (int d = 0; d < PD; d++) {
        if (r + d < nrows) {                               // warp-uniform
          const uint4 c = q[d];
          q[d] = r + d + PD < nrows ? ld_stream_u4(p + (size_t)(r + d + PD) * 32, pol) : padq;
          const float v0 = tab[c.x & 0xffffu], v1 = tab[c.x >> 16], v2 = tab[c.y & 0xffffu], v3 = tab[c.y >> 16];
          const float v4 = tab[c.z & 0xffffu], v5 = tab[c.z >> 16], v6 = tab[c.w & 0xffffu], v7 = tab[c.w >> 16];
          acc = __fadd_rn(acc, v0); acc = __fadd_rn(acc, v1); acc = __fadd_rn(acc, v2); acc = __fadd_rn(acc, v3);
          acc = __fadd_rn(acc, v4); acc = __fadd_rn(acc, v5); acc = __fadd_rn(acc, v6); acc = __fadd_rn(acc, v7);
          if (--left == 0) {
            if (jrow >= 0) atomicAdd(a.acc_fix + jrow, (unsigned long long)__double2ll_rn((double)acc * a.fix_scale));
            jrow = jrow1; jrow1 = jrow2;
            jrow2 = item + 3 < i1 ? a.irow[(size_t)(item + 3) * 32 + lane] : -1;
            acc = 0.f;
            item++;
            if (item - ibase == 32) { ibase += 32; lens = lens_next; lens_next = len_batch(ibase + 32); }
            left = __shfl_sync(kFull, lens, item - ibase);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the iteration, SEGMENTED mode: one band per launch
// Same item / job / warp-run structure as pr_band_kernel, but the band is an L2-resident slice of contrib (not a
// shared-memory table): 4 global ids per unit, predicated evict-last gathers, the gathers of index group r in flight
// while group r-1 is added; every item's per-row partial goes to the row's fixed-point accumulator.
template <int PD>
__global__ void __launch_bounds__(kSellThreads, 1)
pr_seg_kernel(BandArgs a, int32_t job0) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t pol = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  const int4 J = a.job[job0 + blockIdx.x];
  const int32_t i0 = a.wrun[J.y + w], i1 = a.wrun[J.y + w + 1];
  if (i0 >= i1) return;
  const uint32_t u0 = a.item_ptr[i0], u1 = a.item_ptr[i1];
  const int4 *p = reinterpret_cast<const int4 *>(a.bsell) + u0 + lane;
  const uint32_t nrows = (u1 - u0) >> 5;
  auto len_batch = [&](int32_t ib) -> uint32_t {
    const int32_t i = ib + lane;
    return i < i1 ? (a.item_ptr[i + 1] - a.item_ptr[i]) >> 5 : 0u;
  };
  int32_t ibase = i0, item = i0;
  uint32_t lens = len_batch(ibase), lens_next = len_batch(ibase + 32);
  uint32_t left = __shfl_sync(kFull, lens, 0);
  int32_t jrow = a.irow[(size_t)i0 * 32 + lane], jrow1 = -1, jrow2 = -1;
  if (i0 + 1 < i1) jrow1 = a.irow[(size_t)(i0 + 1) * 32 + lane];
  if (i0 + 2 < i1) jrow2 = a.irow[(size_t)(i0 + 2) * 32 + lane];
  const int4 none = make_int4(-1, -1, -1, -1);
  int4 q[PD];
#pragma unroll
  for (int d = 0; d < PD; d++) q[d] = (uint32_t)d < nrows ? ld_stream_v4(p + 32 * d, pol) : none;
  auto pull = [&](int c) -> float {
    float v = 0.f;
    if (c >= 0) v = ld_gather_f32(a.contrib_in + c, pol_last);
    return v;
  };
  float acc = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
  for (uint32_t r = 0; r <= nrows; r += PD) {
#pragma unroll
    for (int d = 0; d < PD; d++) {
      const uint32_t idx = r + d;
      float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
      if (idx < nrows) {                                   // warp-uniform: issue the gathers of group idx
        const int4 c = q[d];
        q[d] = idx + PD < nrows ? ld_stream_v4(p + (size_t)(idx + PD) * 32, pol) : none;
        n0 = pull(c.x); n1 = pull(c.y); n2 = pull(c.z); n3 = pull(c.w);
      }
      if (idx > 0 && idx <= nrows) {                       // add group idx - 1
        acc = __fadd_rn(acc, v0); acc = __fadd_rn(acc, v1); acc = __fadd_rn(acc, v2); acc = __fadd_rn(acc, v3);
        if (--left == 0) {
          if (jrow >= 0) atomicAdd(a.acc_fix + jrow, (unsigned long long)__double2ll_rn((double)acc * a.fix_scale));
          jrow = jrow1; jrow1 = jrow2;
          jrow2 = item + 3 < i1 ? a.irow[(size_t)(item + 3) * 32 + lane] : -1;
          acc = 0.f;
          item++;
          if (item - ibase == 32) { ibase += 32; lens = lens_next; lens_next = len_batch(ibase + 32); }
          left = __shfl_sync(kFull, lens, item - ibase);
        }
      }
      v0 = n0; v1 = n1; v2 = n2; v3 = n3;
    }
  }
}

// ------------------------------------------------------------------ the iteration: main sum + band partials -> epilogue
// pr_band_kernel has already added every band partial of row j into acc_fix[j] as a fixed-point integer (integer
// addition commutes: the order of the atomics does not matter, the sum is exact and bit-reproducible), so this is one
// coalesced pass: main sum + band sum, row epilogue, accumulator back to zero.  (A per-row gather of per-item partial
// slots, and cooperative warp / CTA versions of it, were measured slower: profiles/r1_pr_band_ab.txt.)
__global__ void __launch_bounds__(256, 4)
pr_band_finalize_fix(SellArgs a, long long *__restrict__ acc_fix, double inv_scale) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double err = 0.0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n_band_rows; j += (int64_t)gridDim.x * blockDim.x) {
    const long long f = acc_fix[j];
    acc_fix[j] = 0;
    const float acc = (float)((double)__ldcs(a.acc_main + j) + (double)f * inv_scale);
    pr_epilogue_core(a, j, acc, err);
  }
  err = warp_sum(err);
  if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  contrib_flush(a);
}

// ------------------------------------------------------------------ host: tables
// Work tables of a SELL array from its slice pointers (same meaning as in pull_prepare; widths need not be monotone).
static void make_work_tables(const std::vector<uint32_t> &sptr, int32_t n_slices, int32_t n_exact, std::vector<int32_t> &chunk,
                             std::vector<int32_t> &hslice, std::vector<int32_t> &hfirst, std::vector<int2> &hseg) {
  const uint64_t tot = sptr[n_slices];
  const int32_t n_chunks = (int32_t)(tot / kGroupCh + 1);
  chunk.assign((size_t)n_chunks + 1, 0);
  int32_t s = 0;
  for (int32_t k = 0; k < n_chunks; k++) {
    while (s < n_slices && sptr[s] < (uint64_t)k * kGroupCh) s++;
    // empty slices (rows whose ids all moved into bands) at the head of a chunk are never visited: in the segmented
    // mode that is EVERY band slice, and one warp walking a million empty slices costs 100 ms
    while (s < n_slices && sptr[s + 1] == sptr[s]) s++;
    chunk[k] = s;
  }
  chunk[n_chunks] = n_slices;
  hslice.clear(); hfirst.clear(); hseg.clear();
  for (int32_t t = n_exact; t < n_slices; t++) {          // (exact slices are items of their own, never cut)
    const uint32_t sz = sptr[t + 1] - sptr[t];
    if (sz <= (uint32_t)kGroupCh) continue;
    hslice.push_back(t);
    hfirst.push_back((int32_t)hseg.size());
    for (uint32_t q = 0; q < (sz + kGroupCh - 1) / kGroupCh; q++) hseg.push_back(make_int2(t, (int)q));
  }
  hfirst.push_back((int32_t)hseg.size());
}

template <typename T>
static int up(gdn_graph *g, T **dptr, const T *h, size_t n) {
  GDN_CUDA(cudaMalloc((void **)dptr, std::max<size_t>(n, 4) * sizeof(T)));
  g->device_bytes += n * sizeof(T);
  if (n) GDN_CUDA(cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, lib().stream));
  return GDN_OK;
}

// ------------------------------------------------------------------ host: tables of the band layout (device-free)
// Everything the host derives from the count matrix: each band's row order and ranks, band slices, items, the
// (item, lane) -> row map, jobs and warp runs, and the compacted main array's slice pointers and work tables.
// No CUDA call in here: tests/test_host.py drives it through gdn_band_host_probe with synthetic counts.
struct BandHost {
  bool ok = false;
  std::vector<uint32_t> rank, item_ptr, bslice_ptr, bslice_item, sp2;
  std::vector<int32_t> bslice_first, item_band, irow, job_first, wrun, chunk, hslice, hfirst;
  std::vector<uint64_t> band_item0;
  std::vector<int4> job;
  std::vector<int2> hseg;
  uint64_t units = 0, moved = 0, pairs = 0, tot2 = 0;
  int32_t n_items = 0;
};

static void band_host_tables(int B, int64_t n_rows, int32_t nb, uint32_t W, bool seg, int n_cta, const std::vector<uint32_t> &cnt,
                             const std::vector<uint32_t> &remw, const std::vector<uint32_t> &sp, int32_t n_slices, int32_t n_exact, BandHost &H) {
  // host: every band's rows sorted by their count in it; slices, items, ranks
  struct PerBand {
    std::vector<uint32_t> order;           // rows by (count desc, row asc)
    std::vector<uint32_t> slice_units;     // unit offset of each band slice inside the band
    std::vector<uint32_t> slice_item;      // first item of each band slice inside the band
    std::vector<uint32_t> item_ng;         // index groups per lane of each item
    uint64_t units = 0, moved = 0;
  };
  std::vector<PerBand> pb((size_t)B);
  std::vector<uint32_t> rank(cnt.size());
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; b++) {
    PerBand &q = pb[b];
    const uint32_t *c = cnt.data() + (size_t)b * n_rows;
    uint32_t *rk = rank.data() + (size_t)b * n_rows;
    for (int64_t j = 0; j < n_rows; j++) { rk[j] = kNone; if (c[j]) { q.order.push_back((uint32_t)j); q.moved += c[j]; } }
    if (!seg) {
      std::stable_sort(q.order.begin(), q.order.end(), [&](uint32_t x, uint32_t y) { return c[x] > c[y]; });
    } else {
      // segmented mode: every row is here, so sort by count only inside windows of 64 K consecutive entries -- the rows
      // of an item stay close together and the accumulator atomics of a pass sweep through memory instead of scattering
      // (counting sort: the counts of a graph without hubs are small)
      const size_t win = 65536;
      std::vector<uint32_t> tmp(win), start;
      for (size_t a0 = 0; a0 < q.order.size(); a0 += win) {
        const size_t a1 = std::min(q.order.size(), a0 + win);
        uint32_t cmax = 0;
        for (size_t i = a0; i < a1; i++) cmax = std::max(cmax, c[q.order[i]]);
        if (cmax > (1u << 20)) {
          std::stable_sort(q.order.begin() + a0, q.order.begin() + a1, [&](uint32_t x, uint32_t y) { return c[x] > c[y]; });
          continue;
        }
        start.assign((size_t)cmax + 2, 0);
        for (size_t i = a0; i < a1; i++) start[cmax - c[q.order[i]] + 1]++;          // bucket = cmax - count: descending
        for (size_t k = 1; k < start.size(); k++) start[k] += start[k - 1];
        for (size_t i = a0; i < a1; i++) tmp[start[cmax - c[q.order[i]]]++] = q.order[i];
        std::copy(tmp.begin(), tmp.begin() + (a1 - a0), q.order.begin() + a0);
      }
    }
    const size_t n = q.order.size();
    for (size_t pos = 0; pos < n; pos++) rk[q.order[pos]] = (uint32_t)pos;
    for (size_t s0 = 0; s0 < n; s0 += 32) {
      uint32_t cmx = c[q.order[s0]];                                   // (windowed order: not necessarily the first row)
      if (seg) for (size_t i = s0; i < std::min(n, s0 + 32); i++) cmx = std::max(cmx, c[q.order[i]]);
      const uint32_t ng = (cmx + W - 1) / W;
      q.slice_units.push_back((uint32_t)q.units);
      q.slice_item.push_back((uint32_t)q.item_ng.size());
      for (uint32_t t = 0; t < ng; t += kBandSeg) q.item_ng.push_back(std::min<uint32_t>(kBandSeg, ng - t));
      q.units += 32ull * ng;
    }
  }
  uint64_t units = 0, n_items64 = 0, n_bslices = 0, moved = 0, pairs = 0;
  std::vector<uint64_t> band_unit0((size_t)B + 1), band_item0((size_t)B + 1);
  std::vector<int32_t> bslice_first((size_t)B + 1);
  for (int b = 0; b < B; b++) {
    band_unit0[b] = units; band_item0[b] = n_items64; bslice_first[b] = (int32_t)n_bslices;
    units += pb[b].units; n_items64 += pb[b].item_ng.size(); n_bslices += pb[b].slice_units.size();
    moved += pb[b].moved; pairs += pb[b].order.size();
  }
  band_unit0[B] = units; band_item0[B] = n_items64; bslice_first[B] = (int32_t)n_bslices;
  if (n_items64 == 0 || units >= 0xfffffff0ull / 8 * 8 || n_items64 * 32 >= 0xfffffff0ull) {
    return;                                          // nothing qualifies (or 32-bit unit offsets would overflow): plain layout
  }
  const int32_t n_items = (int32_t)n_items64;
  std::vector<uint32_t> item_ptr((size_t)n_items + 1), bslice_ptr(n_bslices), bslice_item(n_bslices);
  std::vector<int32_t> item_band((size_t)n_items);
  for (int b = 0; b < B; b++) {
    const PerBand &q = pb[b];
    uint64_t u = band_unit0[b];
    for (size_t i = 0; i < q.item_ng.size(); i++) {
      item_ptr[band_item0[b] + i] = (uint32_t)u;
      item_band[band_item0[b] + i] = b;
      u += 32ull * q.item_ng[i];
    }
    for (size_t s0 = 0; s0 < q.slice_units.size(); s0++) {
      bslice_ptr[bslice_first[b] + s0] = (uint32_t)(band_unit0[b] + q.slice_units[s0]);
      bslice_item[bslice_first[b] + s0] = (uint32_t)(band_item0[b] + q.slice_item[s0]);
    }
  }
  item_ptr[n_items] = (uint32_t)units;

  // sorted row of every (item, lane): the rows of a band slice, repeated for each of its segments
  std::vector<int32_t> irow((size_t)n_items * 32);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; b++) {
    const PerBand &q = pb[b];
    const size_t n = q.order.size();
    for (size_t sl = 0; sl < q.slice_item.size(); sl++) {
      const size_t it0 = band_item0[b] + q.slice_item[sl];
      const size_t it1 = sl + 1 < q.slice_item.size() ? band_item0[b] + q.slice_item[sl + 1] : band_item0[b + 1];
      for (size_t it = it0; it < it1; it++)
        for (int l = 0; l < 32; l++) irow[it * 32 + l] = sl * 32 + l < n ? (int32_t)q.order[sl * 32 + l] : -1;
    }
  }
  // jobs: equal-cost contiguous runs of items per CTA, cut at band boundaries, each cut into 32 warp runs
  std::vector<uint64_t> pc((size_t)n_items + 1, 0);
  for (int32_t i = 0; i < n_items; i++) pc[i + 1] = pc[i] + ((item_ptr[i + 1] - item_ptr[i]) >> 5) + 2;
  auto first_at = [&](uint64_t cost) -> int32_t { return (int32_t)(std::lower_bound(pc.begin(), pc.end(), cost) - pc.begin()); };
  std::vector<int4> job;
  std::vector<int32_t> job_first((size_t)n_cta + 1, 0), wrun;
  auto push_job = [&](int b, int32_t lo, int32_t e) {
    const int32_t w0 = (int32_t)wrun.size();
    for (int w = 0; w <= 32; w++) {
      int32_t x = w == 32 ? e : first_at(pc[lo] + (pc[e] - pc[lo]) * w / 32);
      x = std::max(lo, std::min(x, e));
      wrun.push_back(x);
    }
    job.push_back(make_int4(b, w0, lo, e));
  };
  if (seg) {
    // one launch per band: job b * n_cta + c = CTA c's equal-cost share of band b's items (possibly empty)
    for (int b = 0; b < B; b++) {
      const int32_t bl = (int32_t)band_item0[b], bh = (int32_t)band_item0[b + 1];
      for (int c = 0; c < n_cta; c++) {
        int32_t lo = bl, hi = bh;
        if (bh > bl) {
          lo = std::max(bl, std::min(first_at(pc[bl] + (pc[bh] - pc[bl]) * c / n_cta), bh));
          hi = c == n_cta - 1 ? bh : std::max(bl, std::min(first_at(pc[bl] + (pc[bh] - pc[bl]) * (c + 1) / n_cta), bh));
        }
        push_job(b, lo, std::max(lo, hi));
      }
    }
  }
  for (int c = 0; c < n_cta && !seg; c++) {
    job_first[c] = (int32_t)job.size();
    int32_t lo = std::min(first_at(pc[n_items] * c / n_cta), n_items), hi = std::min(first_at(pc[n_items] * (c + 1) / n_cta), n_items);
    if (c == n_cta - 1) hi = n_items;
    while (lo < hi) {
      const int b = item_band[lo];
      const int32_t e = (int32_t)std::min<uint64_t>((uint64_t)hi, band_item0[b + 1]);
      push_job(b, lo, e);
      lo = e;
    }
  }
  job_first[n_cta] = seg ? 0 : (int32_t)job.size();

  // compacted main array: band slices get their remaining width, the others keep theirs
  std::vector<uint32_t> sp2((size_t)n_slices + 1);
  uint64_t tot2 = 0;
  for (int32_t s = 0; s < n_slices; s++) {
    sp2[s] = (uint32_t)tot2;
    tot2 += s < nb ? 32ull * ((remw[s] + 3) / 4) : (uint64_t)(sp[s + 1] - sp[s]);
  }
  sp2[n_slices] = (uint32_t)tot2;
  std::vector<int32_t> chunk, hslice, hfirst;
  std::vector<int2> hseg;
  make_work_tables(sp2, n_slices, n_exact, chunk, hslice, hfirst, hseg);
  H.rank = std::move(rank); H.item_ptr = std::move(item_ptr); H.bslice_ptr = std::move(bslice_ptr); H.bslice_item = std::move(bslice_item);
  H.sp2 = std::move(sp2);
  H.bslice_first = std::move(bslice_first); H.item_band = std::move(item_band); H.irow = std::move(irow);
  H.job_first = std::move(job_first); H.wrun = std::move(wrun); H.chunk = std::move(chunk); H.hslice = std::move(hslice); H.hfirst = std::move(hfirst);
  H.band_item0 = std::move(band_item0); H.job = std::move(job); H.hseg = std::move(hseg);
  H.units = units; H.moved = moved; H.pairs = pairs; H.tot2 = tot2; H.n_items = n_items;
  H.ok = true;
}

// Invariants of the host tables (what the kernels rely on); returns 0 or the negative number of the first one broken.
static int band_host_check(int B, int64_t n_rows, int32_t nb, uint32_t W, bool seg, int n_cta, const std::vector<uint32_t> &cnt,
                           const std::vector<uint32_t> &remw, const std::vector<uint32_t> &sp, int32_t n_slices, int32_t n_exact, const BandHost &H) {
  const int32_t n_items = H.n_items;
  if ((int64_t)H.item_ptr.size() != (int64_t)n_items + 1 || H.item_ptr[0] != 0 || H.item_ptr[n_items] != H.units) return -1;
  // 1. ranks of a band are a bijection between its rows and 0 .. n_b - 1; slices are wide enough for every row in them
  uint64_t moved = 0, pairs = 0;
  for (int b = 0; b < B; b++) {
    const uint32_t *c = cnt.data() + (size_t)b * n_rows, *rk = H.rank.data() + (size_t)b * n_rows;
    size_t n = 0;
    for (int64_t j = 0; j < n_rows; j++) n += c[j] != 0;
    std::vector<uint8_t> seen(n, 0);
    const int32_t bs0 = H.bslice_first[b], bs1 = H.bslice_first[b + 1];
    if ((size_t)(bs1 - bs0) != (n + 31) / 32) return -2;
    for (int64_t j = 0; j < n_rows; j++) {
      if (!c[j]) { if (rk[j] != kNone) return -3; continue; }
      const uint32_t r = rk[j];
      if (r >= n || seen[r]) return -4;
      seen[r] = 1; moved += c[j]; pairs++;
      // the slice of rank r: its items hold >= c[j] ids per lane, and (item, lane) maps back to row j
      const uint32_t sl = bs0 + (r >> 5);
      const uint32_t it0 = H.bslice_item[sl], it1 = sl + 1 < (uint32_t)bs1 ? H.bslice_item[sl + 1] : (uint32_t)H.band_item0[b + 1];
      if (it0 >= it1 || H.item_ptr[it0] != H.bslice_ptr[sl]) return -5;
      const uint64_t cap = (uint64_t)(H.item_ptr[it1] - H.item_ptr[it0]) / 32 * W;
      if (cap < c[j]) return -6;
      for (uint32_t it = it0; it < it1; it++) {
        if (H.item_band[it] != b) return -7;
        if (H.irow[(size_t)it * 32 + (r & 31)] != (int32_t)j) return -8;
        const uint32_t ng = (H.item_ptr[it + 1] - H.item_ptr[it]) / 32;
        if (ng == 0 || ng > (uint32_t)kBandSeg || (H.item_ptr[it + 1] - H.item_ptr[it]) % 32) return -9;
      }
    }
    // lanes beyond the band's last row map to no row
    if (n % 32) {
      const uint32_t sl = bs1 - 1;
      for (uint32_t it = H.bslice_item[sl]; it < (uint32_t)H.band_item0[b + 1]; it++)
        for (uint32_t l = n % 32; l < 32; l++) if (H.irow[(size_t)it * 32 + l] != -1) return -10;
    }
  }
  if (moved != H.moved || pairs != H.pairs) return -11;
  // 3. jobs: every item belongs to exactly one warp run of one job; a job stays inside one band; runs are ordered
  std::vector<uint8_t> cover((size_t)n_items, 0);
  const size_t n_jobs = H.job.size();
  if (seg && n_jobs != (size_t)B * n_cta) return -16;
  if (!seg && ((int)H.job_first.size() != n_cta + 1 || H.job_first[n_cta] != (int32_t)n_jobs)) return -17;
  for (size_t q = 0; q < n_jobs; q++) {
    const int4 J = H.job[q];
    if (seg && J.x != (int)(q / n_cta)) return -18;
    if (J.y < 0 || (size_t)J.y + 33 > H.wrun.size() || H.wrun[J.y] != J.z || H.wrun[J.y + 32] != J.w) return -19;
    for (int w = 0; w < 32; w++) {
      const int32_t i0 = H.wrun[J.y + w], i1 = H.wrun[J.y + w + 1];
      if (i0 > i1 || i0 < 0 || i1 > n_items) return -20;
      for (int32_t i = i0; i < i1; i++) { if (cover[i]++ || H.item_band[i] != J.x) return -21; }
    }
  }
  for (int32_t i = 0; i < n_items; i++) if (!cover[i]) return -22;
  // 4. the compacted main array: band slices are exactly as wide as their longest remainder, the others keep their width
  if ((int32_t)H.sp2.size() != n_slices + 1 || H.sp2[0] != 0 || H.sp2[n_slices] != H.tot2) return -23;
  for (int32_t t = 0; t < n_slices; t++) {
    const uint32_t w2 = H.sp2[t + 1] - H.sp2[t];
    if (t < nb ? w2 != 32 * ((remw[t] + 3) / 4) : w2 != sp[t + 1] - sp[t]) return -24;
  }
  // 5. work tables of the main array: chunks are ordered, every non-empty light slice is reachable, wide slices are segmented
  const int32_t n_chunks = (int32_t)H.chunk.size() - 1;
  if (n_chunks < 1 || H.chunk[n_chunks] != n_slices) return -25;
  for (int32_t k = 0; k < n_chunks; k++) if (H.chunk[k] > H.chunk[k + 1]) return -26;
  for (int32_t t = 0; t < H.chunk[0]; t++) if (H.sp2[t + 1] != H.sp2[t]) return -27;      // skipped slices must be empty
  size_t hs = 0;
  for (int32_t t = 0; t < n_slices; t++) {
    const uint32_t sz = H.sp2[t + 1] - H.sp2[t];
    if (t >= n_exact && sz > (uint32_t)kGroupCh) {
      if (hs >= H.hslice.size() || H.hslice[hs] != t) return -28;
      if (H.hfirst[hs + 1] - H.hfirst[hs] != (int32_t)((sz + kGroupCh - 1) / kGroupCh)) return -29;
      hs++;
    }
  }
  if (hs != H.hslice.size() || H.hfirst[hs] != (int32_t)H.hseg.size()) return -30;
  return 0;
}

void band_free(BandLayout &b) {
  cudaFree(b.bsell); cudaFree(b.band_start); cudaFree(b.band_len); cudaFree(b.item_ptr); cudaFree(b.job); cudaFree(b.job_first); cudaFree(b.wrun); cudaFree(b.irow); cudaFree(b.acc_fix);
  cudaFree(b.acc_main); cudaFree(b.sell); cudaFree(b.slice_ptr);
  cudaFree(b.chunk_slice); cudaFree(b.heavy_slice); cudaFree(b.heavy_first); cudaFree(b.heavy_seg); cudaFree(b.partial);
  b = BandLayout();
}

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Build the banded layout from the (already built) SELL array.  Leaves b.built = false -- and the plain layout in
// use -- when the graph is too small or too flat for any (row, band) pair to qualify, and ALSO when the build itself
// fails (out of device / host memory at the scales this layout targets): it is an optimisation, so the temporaries are
// released, the CUDA error is cleared and the solve goes on with the plain layout.
namespace {
struct Tmp {                 // device temporaries of the build: freed on every way out
  void *p = nullptr;
  ~Tmp() { if (p) cudaFree(p); }
  template <typename T> T *as() { return (T *)p; }
};
}  // namespace

static int band_build_body(gdn_graph *g) {
  PullLayout &L = g->pull;
  BandLayout &bd = L.band;
  const int B_want = env_int("GDN_PR_BANDS", 64);
  if (B_want <= 0 || !L.prepared || !L.sell || L.n_slices < 2 || L.h_slice_ptr.empty()) return GDN_OK;
  const std::vector<uint32_t> &sp = L.h_slice_ptr;
  // Is there a hot set?  Share of the SELL array held by the rows of the 64 hottest bands' worth of ids (rows and
  // columns are ranked by the same degrees on a symmetric graph): 0.87 at Kronecker scale 26, 0.05 at urand-26.
  const int64_t hot_slices = std::min<int64_t>(L.n_slices, (int64_t)64 * kHotMax / 32 / std::max(L.P, 1));   // this rank's share of them
  const double hot_share = (double)sp[hot_slices] / (double)std::max<uint32_t>(sp[L.n_slices], 1u);
  const int e_seg = env_int("GDN_PR_SEGMENT", -1);       // -1: decide here, 0: never, 1: always
  const bool seg = e_seg == 1 || (e_seg < 0 && hot_share < 0.4 && L.Mp * 4 > ((int64_t)64 << 20));
  const int seg_ids = std::max(env_int("GDN_PR_SEG_IDS", (40 << 20) / 4), 64);
  const int band = seg ? seg_ids : std::min(std::max(env_int("GDN_PR_BAND_SIZE", kHotMax), 32), kHotMax);
  const int cmin = seg ? 1 : std::max(env_int("GDN_PR_BAND_CMIN", 4), 1);
  const int dmin = seg ? 1 : std::max(env_int("GDN_PR_BAND_DMIN", 64), 1);
  const uint32_t W = seg ? 4 : 8;                        // ids per 16-byte unit
  BandMap mp;
  mp.band = band; mp.P = L.P; mp.H = L.H; mp.Wc = L.Wc;
  mp.n0 = (int32_t)((L.H + band - 1) / band);
  const int B = (int)std::min<int64_t>(std::min(B_want, 96), (int64_t)mp.n0 + (L.Wc + band - 1) / band * L.P);
  mp.B = B;
  if (B <= 0) return GDN_OK;
  if (seg && (int64_t)mp.n0 + (L.Wc + band - 1) / band * L.P > B) return GDN_OK;   // more than 96 slices: plain layout
  // slices whose rows are all at least dmin long: the first row of the NEXT slice is (rows sorted by length, descending)
  int32_t nb = 0;
  while (nb + 1 < L.n_slices && (int64_t)(sp[nb + 2] - sp[nb + 1]) / 32 * 4 >= dmin) nb++;
  if (nb == 0) return GDN_OK;
  const int64_t n_rows = (int64_t)nb * 32;
  cudaStream_t st = lib().stream;
  const int sm = lib().sm_count;
  trace("band_build: begin");

  // pass 1 on the device
  Tmp t_cnt, t_remw, t_bslice_ptr, t_bslice_first;
  GDN_CUDA(cudaMalloc(&t_cnt.p, sizeof(uint32_t) * (size_t)B * n_rows));
  GDN_CUDA(cudaMalloc(&t_remw.p, sizeof(uint32_t) * (size_t)nb));
  uint32_t *d_cnt = t_cnt.as<uint32_t>(), *d_remw = t_remw.as<uint32_t>();
  const size_t smem1 = sizeof(uint32_t) * 4 * (size_t)B * 32, smem2 = 2 * smem1;
  GDN_CUDA(cudaFuncSetAttribute(band_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  GDN_CUDA(cudaFuncSetAttribute(band_fill<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  GDN_CUDA(cudaFuncSetAttribute(band_fill<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  const int grid = (int)std::min<int64_t>((nb + 3) / 4, (int64_t)sm * 16);
  band_count<<<grid, 128, smem1, st>>>(L.sell, L.slice_ptr, nb, mp, d_cnt, n_rows);
  band_select<<<(int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)sm * 8), 256, 0, st>>>(d_cnt, B, n_rows, (uint32_t)cmin, L.sdeg, d_remw, (int64_t)L.n_exact * 32);
  std::vector<uint32_t> cnt, remw;
  try { cnt.resize((size_t)B * n_rows); remw.resize((size_t)nb); }
  catch (const std::bad_alloc &) { set_error("band layout: out of host memory"); return GDN_ERR_NOMEM; }
  GDN_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(uint32_t) * cnt.size(), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaMemcpyAsync(remw.data(), d_remw, sizeof(uint32_t) * remw.size(), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  trace("band_build: counted");

  // host: every band's rows sorted by their count in it; slices, items, ranks, jobs, main array tables
  const int n_cta = sm;
  BandHost H;
  try { band_host_tables(B, n_rows, nb, W, seg, n_cta, cnt, remw, sp, L.n_slices, L.n_exact, H); }
  catch (const std::bad_alloc &) { set_error("band layout: out of host memory"); return GDN_ERR_NOMEM; }
  if (!H.ok) return GDN_OK;
  if (env_int("GDN_BAND_CHECK", 0)) {              // the same invariants the CPU test checks, on the real counts
    const int bad = band_host_check(B, n_rows, nb, W, seg, n_cta, cnt, remw, sp, L.n_slices, L.n_exact, H);
    if (bad) { set_error("band layout: host table invariant %d broken", bad); return GDN_ERR_GRAPH; }
  }
  trace("band_build: host tables");
  auto &rank = H.rank; auto &item_ptr = H.item_ptr; auto &bslice_ptr = H.bslice_ptr;
  auto &sp2 = H.sp2; auto &bslice_first = H.bslice_first; auto &irow = H.irow; auto &job_first = H.job_first; auto &wrun = H.wrun;
  auto &chunk = H.chunk; auto &hslice = H.hslice; auto &hfirst = H.hfirst; auto &job = H.job; auto &hseg = H.hseg;
  const uint64_t units = H.units, moved = H.moved, pairs = H.pairs, tot2 = H.tot2;
  const int32_t n_items = H.n_items;

  bd.seg = seg;
  bd.B = B; bd.band = band; bd.cmin = cmin; bd.dmin = dmin; bd.n_rows = n_rows;
  bd.n_units = units; bd.n_items = n_items; bd.n_jobs = (int32_t)job.size(); bd.n_cta = n_cta;
  bd.n_groups = tot2; bd.n_chunks = (int32_t)chunk.size() - 1;
  bd.n_heavy_slices = (int32_t)hslice.size(); bd.n_heavy_segs = (int32_t)hseg.size();
  bd.moved = moved; bd.pairs = pairs;
  uint32_t *d_rank = d_cnt;                         // the counts are not needed on the device any more
  GDN_CUDA(cudaMemcpyAsync(d_rank, rank.data(), sizeof(uint32_t) * rank.size(), cudaMemcpyHostToDevice, st));
  GDN_CHECK(up(g, (uint32_t **)&t_bslice_ptr.p, bslice_ptr.data(), bslice_ptr.size()));
  GDN_CHECK(up(g, (int32_t **)&t_bslice_first.p, bslice_first.data(), bslice_first.size()));
  std::vector<int64_t> band_start((size_t)B);
  std::vector<int32_t> band_len((size_t)B);
  for (int b = 0; b < B; b++) band_range(mp, b, band_start[b], band_len[b]);
  GDN_CHECK(up(g, &bd.band_start, band_start.data(), band_start.size()));
  GDN_CHECK(up(g, &bd.band_len, band_len.data(), band_len.size()));
  GDN_CHECK(up(g, &bd.item_ptr, item_ptr.data(), item_ptr.size()));
  GDN_CHECK(up(g, &bd.job, job.data(), job.size()));
  GDN_CHECK(up(g, &bd.job_first, job_first.data(), job_first.size()));
  GDN_CHECK(up(g, &bd.wrun, wrun.data(), wrun.size()));
  GDN_CHECK(up(g, &bd.irow, irow.data(), irow.size()));
  GDN_CUDA(cudaMalloc((void **)&bd.acc_fix, sizeof(long long) * (size_t)n_rows));
  GDN_CUDA(cudaMemsetAsync(bd.acc_fix, 0, sizeof(long long) * (size_t)n_rows, st));
  GDN_CHECK(up(g, &bd.slice_ptr, sp2.data(), sp2.size()));
  GDN_CHECK(up(g, &bd.chunk_slice, chunk.data(), chunk.size()));
  if (bd.n_heavy_slices) {
    GDN_CHECK(up(g, &bd.heavy_slice, hslice.data(), hslice.size()));
    GDN_CHECK(up(g, &bd.heavy_first, hfirst.data(), hfirst.size()));
    GDN_CHECK(up(g, &bd.heavy_seg, hseg.data(), hseg.size()));
    GDN_CUDA(cudaMalloc((void **)&bd.partial, sizeof(float) * 32 * (size_t)bd.n_heavy_segs));
  }
  GDN_CUDA(cudaMalloc((void **)&bd.bsell, sizeof(uint4) * units + 256));
  GDN_CUDA(cudaMalloc((void **)&bd.sell, sizeof(int4) * std::max<uint64_t>(tot2, 1) + 256));
  GDN_CUDA(cudaMalloc((void **)&bd.acc_main, sizeof(float) * (size_t)n_rows));
  GDN_CUDA(cudaMemsetAsync(bd.bsell, seg ? 0xff : (int)(kBandPadId & 0xff), sizeof(uint4) * units + 256, st));
  GDN_CUDA(cudaMemsetAsync(bd.sell, 0xff, sizeof(int4) * (size_t)sp2[nb] + (tot2 == sp2[nb] ? 256 : 0), st));
  GDN_CUDA(cudaMemsetAsync(bd.acc_main, 0, sizeof(float) * (size_t)n_rows, st));
  if (tot2 > sp2[nb])
    GDN_CUDA(cudaMemcpyAsync(bd.sell + sp2[nb], L.sell + sp[nb], sizeof(int4) * (size_t)(tot2 - sp2[nb]) + 256, cudaMemcpyDeviceToDevice, st));
  if (seg)
    band_fill<true><<<grid, 128, smem2, st>>>(L.sell, L.slice_ptr, nb, mp, d_rank, n_rows, t_bslice_ptr.as<uint32_t>(), t_bslice_first.as<int32_t>(),
                                              (uint16_t *)bd.bsell, bd.sell, bd.slice_ptr);
  else
    band_fill<false><<<grid, 128, smem2, st>>>(L.sell, L.slice_ptr, nb, mp, d_rank, n_rows, t_bslice_ptr.as<uint32_t>(), t_bslice_first.as<int32_t>(),
                                               (uint16_t *)bd.bsell, bd.sell, bd.slice_ptr);
  double spread[2] = {0, 0};
  if (!seg && env_int("GDN_PR_BAND_SPREAD", 1)) {   // pass 3: every row's ids re-ordered inside its items against bank conflicts
    unsigned long long *d_stats = nullptr, h_stats[2] = {0, 0};
    GDN_CUDA(cudaMalloc((void **)&d_stats, sizeof(h_stats)));
    GDN_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(h_stats), st));
    band_spread<<<sm * 8, 256, 0, st>>>((uint16_t *)bd.bsell, bd.item_ptr, n_items, d_stats);
    GDN_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    GDN_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_stats);
    const double positions = (double)units / 4.0;            // warp-wide table reads of pr_band_kernel: 8 per 32 units
    spread[0] = (double)h_stats[0] / std::max(positions, 1.0); spread[1] = (double)h_stats[1] / std::max(positions, 1.0);
  }
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  GDN_CUDA(cudaFuncSetAttribute(pr_band_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kBandTab)));
  g->device_bytes += sizeof(uint4) * units + sizeof(int4) * tot2 + (sizeof(float) + sizeof(long long)) * (size_t)n_rows;
  bd.built = true;
  if (getenv("GDN_TRACE"))
    fprintf(stderr, "[gdn] %s layout: B=%d band=%d cmin=%d rows=%lld  moved=%llu ids (%.1f %% of nnz) in %llu pairs, %d items, "
                    "%llu padded ids (x%.2f), %d jobs; main array %llu -> %llu groups; lanes per bank and table read %.2f -> %.2f\n",
            seg ? "segmented" : "band", B, band, cmin, (long long)n_rows, (unsigned long long)moved, 100.0 * (double)moved / (double)std::max<uint64_t>(g->in.nnz, 1),
            (unsigned long long)pairs, n_items, (unsigned long long)(units * W), (double)(units * W) / (double)std::max<uint64_t>(moved, 1),
            bd.n_jobs, (unsigned long long)L.n_groups, (unsigned long long)tot2, spread[0], spread[1]);
  trace("band_build: done");
  return GDN_OK;
}

int band_build(gdn_graph *g) {
  BandLayout &bd = g->pull.band;
  if (bd.tried) return GDN_OK;
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = band_build_body(g);
  if (rc != GDN_OK || !bd.built) {
    // not built (nothing qualifies) or failed half-way: drop whatever was allocated and stay on the plain layout
    band_free(bd);
    cudaGetLastError();
    if (rc != GDN_OK && getenv("GDN_TRACE")) fprintf(stderr, "[gdn] band layout not built (%s): plain layout in use\n", gdn_last_error());
  }
  bd.tried = true;
  bd.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return GDN_OK;
}

}  // namespace gdn

// Host-only probe of the id -> band map (tests/test_host.py): no device needed.
extern "C" int gdn_band_map_probe(int64_t H, int64_t Wc, int32_t P, int32_t band, int32_t B, int64_t id, int32_t *band_out,
                                  int32_t *local_out, int64_t *start_out, int32_t *len_out) {
  if (H < 0 || Wc < 0 || P < 1 || band < 1 || B < 1 || id < 0 || !band_out || !local_out || !start_out || !len_out) return GDN_ERR_ARG;
  gdn::BandMap mp;
  mp.band = band; mp.P = P; mp.H = H; mp.Wc = Wc; mp.B = B;
  mp.n0 = (int32_t)((H + band - 1) / band);
  uint32_t loc = 0;
  const uint32_t b = gdn::band_of(mp, id, loc);
  *band_out = (b < (uint32_t)B) ? (int32_t)b : -1;
  *local_out = (int32_t)loc;
  *start_out = 0; *len_out = 0;
  if (*band_out >= 0) gdn::band_range(mp, *band_out, *start_out, *len_out);
  return GDN_OK;
}

// Host-only probe of the band layout's host tables (tests/test_host.py): builds them from a caller-supplied count
// matrix cnt[B][n_rows] (n_rows = 32 * number of band slices), the remaining widths of the band slices and the slice
// pointers of the plain array, and checks the invariants the kernels rely on.  Returns 0, 1 (nothing qualifies) or the
// negative number of the broken invariant.  stats[0..5] = items, units, pairs, moved ids, jobs, main groups.
extern "C" int gdn_band_host_probe(int32_t B, int64_t n_rows, int32_t ids_per_unit, int32_t segmented, int32_t n_cta,
                                   const uint32_t *cnt, const uint32_t *rem_w, const uint32_t *slice_ptr, int32_t n_slices,
                                   int64_t *stats) {
  if (B < 1 || n_rows < 32 || n_rows % 32 || (ids_per_unit != 4 && ids_per_unit != 8) || n_cta < 1 || !cnt || !rem_w || !slice_ptr ||
      n_slices < n_rows / 32) return GDN_ERR_ARG;
  const int32_t nb = (int32_t)(n_rows / 32);
  std::vector<uint32_t> c(cnt, cnt + (size_t)B * n_rows), rw(rem_w, rem_w + nb), sp(slice_ptr, slice_ptr + n_slices + 1);
  gdn::BandHost H;
  gdn::band_host_tables(B, n_rows, nb, (uint32_t)ids_per_unit, segmented != 0, n_cta, c, rw, sp, n_slices, 0, H);
  if (!H.ok) return 1;
  if (stats) {
    stats[0] = H.n_items; stats[1] = (int64_t)H.units; stats[2] = (int64_t)H.pairs; stats[3] = (int64_t)H.moved;
    stats[4] = (int64_t)H.job.size(); stats[5] = (int64_t)H.tot2;
  }
  return gdn::band_host_check(B, n_rows, nb, (uint32_t)ids_per_unit, segmented != 0, n_cta, c, rw, sp, n_slices, 0, H);
}

namespace gdn {

int band_launch(gdn_graph *g, const SellArgs &sa, double fix_scale, cudaStream_t s) {
  const BandLayout &bd = g->pull.band;
  BandArgs a;
  a.bsell = bd.bsell; a.item_ptr = bd.item_ptr; a.job = bd.job; a.job_first = bd.job_first; a.wrun = bd.wrun;
  a.pf_groups = env_int("GDN_PR_BAND_PF", 8);      // 8 index groups = 4 KB ahead (0.84 ms; 16: 0.87, 24: 0.98, none: 1.10)
  a.band_start = bd.band_start; a.band_len = bd.band_len; a.contrib_in = sa.contrib_in; a.done = sa.done;
  a.irow = bd.irow; a.acc_fix = (unsigned long long *)bd.acc_fix; a.fix_scale = fix_scale;
  if (bd.seg) {
    for (int b = 0; b < bd.B; b++) pr_seg_kernel<4><<<bd.n_cta, kSellThreads, 0, s>>>(a, b * bd.n_cta);
  } else {
    pr_band_kernel<4><<<bd.n_cta, kSellThreads, sizeof(float) * kBandTab, s>>>(a);
  }
  return GDN_OK;
}

int band_launches(const gdn_graph *g) { return g->pull.band.seg ? g->pull.band.B : 1; }

int band_finalize_grid(const gdn_graph *g) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((g->pull.band.n_rows + 255) / 256, (int64_t)lib().sm_count * 8));
}

int band_finalize_launch(gdn_graph *g, const SellArgs &a, double fix_scale, int grid, cudaStream_t s) {
  pr_band_finalize_fix<<<grid, 256, 0, s>>>(a, g->pull.band.acc_fix, 1.0 / fix_scale);
  return GDN_OK;
}

// start of a solve: the fixed-point accumulators must be zero (pr_band_finalize_fix leaves them so; an aborted solve may not)
int band_solve_begin(gdn_graph *g, cudaStream_t s) {
  const BandLayout &bd = g->pull.band;
  GDN_CUDA(cudaMemsetAsync(bd.acc_fix, 0, sizeof(long long) * (size_t)bd.n_rows, s));
  return GDN_OK;
}

}  // namespace gdn
