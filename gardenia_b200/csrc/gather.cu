// gather.cu -- the pull-direction CSR gather shared by PageRank and SpMV.
//
// Replaces (not ports) the reference's src/pr/{base,warp,vector,lb}.cu and
// src/spmv/{base,warp,vector}.cu.  Semantics follow the OpenMP variants the
// results are checked against: src/pr/omp_base.cc:21-37, src/spmv/omp_base.cc:22-33.
//
// Load balancing is chosen from the degree distribution once, at upload
// (build_schedule): the row space is cut at every kChunk-th non-zero into
// "light blocks" of whole rows (< 2*kChunk non-zeros each); rows longer than
// kChunk are "heavy" and are cut into kSeg-non-zero segments.  Every work item
// (light block or heavy segment) is about the same size and is processed by
// one warp:
//   phase 1  stream column indices (and SpMV values) with 256-bit coalesced
//            L1-bypassing loads, gather the vector element for each, and
//            stage the products in shared memory -- 8 independent gathers in
//            flight per lane;
//   phase 2  (light blocks) one lane per row sums its staged products in row
//            order, i.e. in exactly the sequential fp32 order of the reference,
//            so light rows are bit-identical to the OpenMP result;
//            (heavy segments) warp-shuffle reduction into a partial, summed
//            per row by finalize_heavy in a fixed order.
// No atomics, deterministic run to run.
#include "common.cuh"
#include <cstdlib>

namespace gdn {

constexpr int kChunk = 512;            // non-zeros per light block (target)
constexpr int kCap = 2 * kChunk + 8;   // staged products per warp
constexpr int kSeg = 1024;             // non-zeros per heavy-row segment
constexpr int kWarps = 8;              // warps per CTA
constexpr int kThreads = kWarps * 32;

enum { kModeSpmv = 0, kModePr = 1 };

struct GatherArgs {
  // schedule
  const int32_t *chunk_row;
  int32_t n_chunks;
  const int2 *heavy_seg;
  int32_t n_heavy_segs;
  float *heavy_partial;
  const int32_t *heavy_row;
  const int32_t *heavy_first;
  int32_t n_heavy_rows;
  int64_t rows;
  // SpMV:  y[r] += sum Ax[j] * x[col[j]]
  const float *Ax;
  uint64_t nnz;           // bound for guarded Ax tail loads
  const float *vec;       // x (SpMV) or contrib_in (PR), GLOBAL length m
  float *y;               // SpMV y (local rows)
  // PR
  float *scores;          // local rows
  float *contrib_out;     // GLOBAL length m; this GPU writes [row_lo, row_lo+rows)
  const int32_t *out_degree;  // local rows; nullptr -> degree = row length (symmetric graph)
  int64_t row_lo;
  float base, damp;
  double *err_partial;
  const int32_t *done;    // device flag set once converged: later launches are no-ops
  int err_slot0;          // first err_partial slot of this kernel
  // column window of this pass (SpMV column passes, see spmv_t): entries outside [col_lo, col_hi) add +0.0f
  int32_t col_lo, col_hi;
};

template <int MODE>
__device__ __forceinline__ void row_epilogue(const GatherArgs &a, int64_t r, float acc, int32_t row_len,
                                             double &err) {
  if (MODE == kModeSpmv) {
    __stcs(a.y + r, acc);                              // acc started from y[r], omp_base.cc:25,32; streaming: y is touched once
  } else {
    const float old_score = a.scores[r];
    const float nw = __fadd_rn(a.base, __fmul_rn(a.damp, acc));   // pr/omp_base.cc:32, no FMA contraction
    a.scores[r] = nw;
    err += (double)fabsf(__fsub_rn(nw, old_score));                // :33, double accumulator :22
    const int32_t deg = a.out_degree ? a.out_degree[r] : row_len;
    a.contrib_out[a.row_lo + r] = __fdiv_rn(nw, (float)deg);       // next iteration's :24-25
  }
}

// Stream [b,e) of this CSR, gather, and either stage the products (STAGE) or
// accumulate them per lane.  a0 = b rounded down to a 32-byte boundary.
template <typename OffT, int MODE, bool STAGE>
__device__ __forceinline__ float stream_gather(const GatherArgs &a, const int32_t *__restrict__ col, OffT b,
                                               OffT e, OffT a0, float *sv, uint64_t pol, int lane) {
  float acc0 = 0.f, acc1 = 0.f;
  for (OffT i = a0 + (OffT)lane * 8; i < e; i += 256) {
    int q[8];
    float ax[8];
    float v[8];
    ld_stream_v8(col + i, q);
    if (MODE == kModeSpmv) {
      if ((uint64_t)i + 8 <= a.nnz) {
        ld_stream_v8(a.Ax + i, ax);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) ax[j] = ((uint64_t)i + j < a.nnz) ? a.Ax[i + j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const bool ok = (i + j >= b) && (i + j < e) && (MODE != kModeSpmv || (q[j] >= a.col_lo && q[j] < a.col_hi));
      float g = 0.f;
      if (ok) g = ld_gather_f32(a.vec + q[j], pol);
      v[j] = (MODE == kModeSpmv) ? (ok ? __fmul_rn(g, ax[j]) : 0.f) : g;
    }
    if (STAGE) {
      float4 *dst = reinterpret_cast<float4 *>(sv + (i - a0));
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      acc0 += (v[0] + v[1]) + (v[2] + v[3]);
      acc1 += (v[4] + v[5]) + (v[6] + v[7]);
    }
  }
  return acc0 + acc1;
}

template <typename OffT, int MODE>
__global__ void __launch_bounds__(kThreads, 4)
gather_kernel(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, GatherArgs a) {
  __shared__ __align__(32) float s_vals[kWarps][kCap];
  if (MODE == kModePr && *a.done) return;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int64_t n_items = (int64_t)a.n_chunks + a.n_heavy_segs;
  float *sv = s_vals[wib];
  const uint64_t pol = l2_policy_evict_last();
  double err = 0.0;

  for (int64_t item = warp; item < n_items; item += nwarps) {
    if (item < a.n_chunks) {
      // ---- light block: whole rows [r0, r1)
      const int32_t r0 = __ldcs(a.chunk_row + item);
      int32_t r1 = __ldcs(a.chunk_row + item + 1);
      if (r1 > r0) {
        const OffT lb = rowptr[r1 - 1], le = rowptr[r1];
        if (le - lb > (OffT)kChunk) r1--;               // heavy last row: handled as segments
      }
      if (r1 <= r0) continue;                           // warp-uniform
      const OffT b = rowptr[r0], e = rowptr[r1];
      const OffT a0 = b & ~(OffT)7;
      stream_gather<OffT, MODE, true>(a, col, b, e, a0, sv, pol, lane);
      __syncwarp();
      for (int32_t r = r0 + lane; r < r1; r += 32) {
        const int32_t s = (int32_t)(rowptr[r] - a0), t = (int32_t)(rowptr[r + 1] - a0);
        float acc = (MODE == kModeSpmv) ? __ldcs(a.y + r) : 0.f;
        for (int32_t j = s; j < t; j++) acc = __fadd_rn(acc, sv[j]);
        row_epilogue<MODE>(a, r, acc, t - s, err);
      }
      __syncwarp();
    } else {
      // ---- heavy segment: kSeg non-zeros of one long row -> one partial
      const int32_t h = (int32_t)(item - a.n_chunks);
      const int2 hs = a.heavy_seg[h];
      const OffT rb = rowptr[hs.x], re = rowptr[hs.x + 1];
      const OffT b = rb + (OffT)hs.y * kSeg;
      const OffT e = (re - b > (OffT)kSeg) ? b + kSeg : re;
      const OffT a0 = b & ~(OffT)7;
      float acc = stream_gather<OffT, MODE, false>(a, col, b, e, a0, nullptr, pol, lane);
      acc = warp_sum(acc);
      if (lane == 0) a.heavy_partial[h] = acc;
    }
  }
  if (MODE == kModePr) {
    err = warp_sum(err);
    if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  }
}

// One warp per heavy row: sum its segment partials in a fixed order, epilogue.
template <typename OffT, int MODE>
__global__ void __launch_bounds__(kThreads, 4)
finalize_heavy(const OffT *__restrict__ rowptr, GatherArgs a) {
  if (MODE == kModePr && *a.done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  double err = 0.0;
  for (int64_t hr = warp; hr < a.n_heavy_rows; hr += nwarps) {
    const int32_t row = a.heavy_row[hr];
    const int32_t first = a.heavy_first[hr];
    const OffT len = rowptr[row + 1] - rowptr[row];
    const int32_t nseg = (int32_t)((len + kSeg - 1) / kSeg);
    float acc = 0.f;
    for (int32_t s = lane; s < nseg; s += 32) acc += a.heavy_partial[first + s];
    acc = warp_sum(acc);
    if (lane == 0) {
      if (MODE == kModeSpmv) acc = __fadd_rn(a.y[row], acc);
      row_epilogue<MODE>(a, row, acc, (int32_t)len, err);
    }
  }
  if (MODE == kModePr) {
    err = warp_sum(err);
    if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  }
}

// ---------------------------------------------------------------- SpMV with TMA-staged row segments
// Same schedule, same arithmetic, different data movement: the column indices and matrix values of a
// work item are brought into shared memory by ONE elected lane with two 1-D bulk copies
// (cp.async.bulk -> SASS UBLKCP, completion on a per-warp mbarrier), double-buffered so the copy of
// item k+1 overlaps the gathers of item k.  The LSU then carries only the x gathers: ncu r1 showed the
// register-staged kernel latency-bound with a third of its in-flight sectors being the (always
// HBM-latency) stream.  Products are staged in place over the values.
constexpr int kTmaWarps = 4;
constexpr int kTmaCap = 1040;            // >= 2*kChunk + alignment slack, multiple of 4

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

struct TmaStage {
  int32_t col[kTmaCap];
  float val[kTmaCap];
};

template <typename OffT>
struct ItemDesc {
  OffT b, e;          // non-zero range
  int32_t r0, r1;     // light block rows, or {row, -1 - segment} for a heavy segment
};

// Describe work item `item` (warp-uniform loads; cheap next to the copy they precede).
template <typename OffT>
__device__ __forceinline__ bool describe(const OffT *__restrict__ rowptr, const GatherArgs &a, int64_t item, ItemDesc<OffT> &d) {
  if (item < a.n_chunks) {
    int32_t r0 = a.chunk_row[item], r1 = a.chunk_row[item + 1];
    if (r1 > r0) {
      const OffT lb = rowptr[r1 - 1], le = rowptr[r1];
      if (le - lb > (OffT)kChunk) r1--;
    }
    if (r1 <= r0) return false;
    d.r0 = r0; d.r1 = r1; d.b = rowptr[r0]; d.e = rowptr[r1];
  } else {
    const int32_t h = (int32_t)(item - a.n_chunks);
    const int2 hs = a.heavy_seg[h];
    const OffT rb = rowptr[hs.x], re = rowptr[hs.x + 1];
    d.b = rb + (OffT)hs.y * kSeg;
    d.e = (re - d.b > (OffT)kSeg) ? d.b + kSeg : re;
    d.r0 = h; d.r1 = -1;
  }
  return true;
}

template <typename OffT>
__global__ void __launch_bounds__(kTmaWarps * 32, 3)
spmv_tma_kernel(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, GatherArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  TmaStage *stages = reinterpret_cast<TmaStage *>(smem_raw) + 2 * wib;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(TmaStage) * 2 * kTmaWarps) + 2 * wib;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t warp = (int64_t)blockIdx.x * kTmaWarps + wib, nwarps = (int64_t)gridDim.x * kTmaWarps;
  const int64_t n_items = (int64_t)a.n_chunks + a.n_heavy_segs;
  const uint64_t pol_stream = l2_policy_evict_first(), pol = l2_policy_evict_last();
  const uint64_t nnz4 = a.nnz & ~3ull;

  // issue the copies of one item into stage st (lane 0); a0 = first staged non-zero
  auto issue = [&](const ItemDesc<OffT> &d, int st) {
    if (lane == 0) {
      const uint64_t a0 = (uint64_t)d.b & ~3ull, a1 = ((uint64_t)d.e + 3) & ~3ull;
      const uint64_t v1 = a1 < nnz4 ? a1 : (nnz4 > a0 ? nnz4 : a0);      // Ax has no slack past nnz: tail is patched by hand
      const uint32_t cb = (uint32_t)(a1 - a0) * 4, vb = (uint32_t)(v1 - a0) * 4;
      mbar_expect_tx(&bars[st], cb + vb);
      if (cb) tma_load_1d(stages[st].col, col + a0, cb, &bars[st], pol_stream);
      if (vb) tma_load_1d(stages[st].val, a.Ax + a0, vb, &bars[st], pol_stream);
    }
  };

  ItemDesc<OffT> cur, nxt;
  int64_t item = warp;
  bool have = false;
  while (item < n_items && !(have = describe(rowptr, a, item, cur))) item += nwarps;
  uint32_t ph0 = 0, ph1 = 0;
  int st = 0;
  if (have) issue(cur, 0);
  while (have) {
    // find and launch the next non-empty item into the other stage
    int64_t nitem = item + nwarps;
    bool have_n = false;
    while (nitem < n_items && !(have_n = describe(rowptr, a, nitem, nxt))) nitem += nwarps;
    if (have_n) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy accesses to that stage are done
      __syncwarp();
      issue(nxt, st ^ 1);
    }
    mbar_wait(&bars[st], st ? ph1 : ph0);
    if (st) ph1 ^= 1; else ph0 ^= 1;
    int32_t *sc = stages[st].col;
    float *sv = stages[st].val;
    const uint64_t a0 = (uint64_t)cur.b & ~3ull;
    const int32_t s_b = (int32_t)((uint64_t)cur.b - a0), s_e = (int32_t)((uint64_t)cur.e - a0);
    if ((uint64_t)cur.e > nnz4) {                     // the last <= 3 values of the matrix
      const uint64_t i = nnz4 + lane;
      if (lane < 3 && i < (uint64_t)cur.e && i >= (uint64_t)cur.b) sv[i - a0] = a.Ax[i];
      __syncwarp();
    }
    if (cur.r1 >= 0) {
      // ---- light block: products in place, then one lane per row sums in the reference's order
      for (int32_t j0 = s_b; j0 < s_e; j0 += 256) {
        float g[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int32_t j = j0 + u * 32 + lane;
          g[u] = (j < s_e) ? ld_gather_f32(a.vec + sc[j], pol) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int32_t j = j0 + u * 32 + lane;
          if (j < s_e) sv[j] = __fmul_rn(g[u], sv[j]);
        }
      }
      __syncwarp();
      for (int32_t r = cur.r0 + lane; r < cur.r1; r += 32) {
        const int32_t s = (int32_t)((uint64_t)rowptr[r] - a0), t = (int32_t)((uint64_t)rowptr[r + 1] - a0);
        float acc = a.y[r];
        for (int32_t j = s; j < t; j++) acc = __fadd_rn(acc, sv[j]);
        a.y[r] = acc;
      }
    } else {
      // ---- heavy segment -> one partial
      float acc0 = 0.f, acc1 = 0.f;
      for (int32_t j0 = s_b; j0 < s_e; j0 += 256) {
        float g[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int32_t j = j0 + u * 32 + lane;
          g[u] = (j < s_e) ? __fmul_rn(ld_gather_f32(a.vec + sc[j], pol), sv[j]) : 0.f;
        }
        acc0 += (g[0] + g[1]) + (g[2] + g[3]);
        acc1 += (g[4] + g[5]) + (g[6] + g[7]);
      }
      const float acc = warp_sum(acc0 + acc1);
      if (lane == 0) a.heavy_partial[cur.r0] = acc;
    }
    __syncwarp();
    cur = nxt; item = nitem; have = have_n; st ^= 1;
  }
}


// ---------------------------------------------------------------- SpMV, software-pipelined (default: 384 threads x 2 CTAs per SM)
// Same schedule and the same arithmetic as gather_kernel<., kModeSpmv> (products staged per warp, one lane per light
// row adds them in the reference's order; heavy segments -> partials), different issue order.  gather_kernel lets a
// warp wait for its column/value loads, then for its gathers, then sum rows: with every warp of an SM in the same
// phase the L1TEX miss path idles between bursts (the convoy measured on PageRank, profiles/r1_pr_tier_probe.txt).
// Here a warp's work is one flat sequence of 128-entry trips that crosses item boundaries, three trips deep:
//      stream loads of trip n+1  |  x gathers of trip n  |  products of trip n-1 -> shared memory (+ row sums at item end)
// and the descriptors of the next two items (chunk_row -> rowptr chains) are prefetched as raw values, so an item
// boundary costs no dependent-load stall.
template <typename OffT>
struct SpTrip {
  OffT tb;              // first entry of the trip (multiple of 4)
  OffT b, e;            // entry range of the item it belongs to
  int32_t r0, r1;       // light block rows [r0, r1), or {heavy segment index, -1}
  int32_t flags;        // 1 valid, 2 last trip of its item
};

template <typename OffT>
struct SpGen {
  const OffT *__restrict__ rowptr;
  const GatherArgs &a;
  int64_t n_items, nwarps;
  // item being cut into trips
  OffT b, e, tb;
  int32_t r0, r1;
  bool have;
  // level B: item k+1, offsets requested;  level A: item k+2, ids requested
  int64_t itB, itA;
  int32_t xB0, xB1, xA0, xA1;
  OffT vB0, vB1, vB2;

  __device__ __forceinline__ SpGen(const OffT *rp, const GatherArgs &a_, int64_t warp, int64_t nwarps_)
      : rowptr(rp), a(a_), n_items((int64_t)a_.n_chunks + a_.n_heavy_segs), nwarps(nwarps_), b(0), e(0), tb(0), r0(0), r1(0),
        have(false), itB(warp - nwarps_), itA(warp - nwarps_), xB0(0), xB1(0), xA0(0), xA1(0), vB0(0), vB1(0), vB2(0) {
    loadA(warp);             // A <- first item
    shift(warp + nwarps_);   // B <- first item (offsets requested), A <- second
    take();                  // current <- first non-empty item
  }
  __device__ __forceinline__ void loadA(int64_t it) {
    itA = it;
    if (it < a.n_chunks) { xA0 = __ldcs(a.chunk_row + it); xA1 = __ldcs(a.chunk_row + it + 1); }
    else if (it < n_items) { const int2 hs = a.heavy_seg[it - a.n_chunks]; xA0 = hs.x; xA1 = hs.y; }
  }
  // B <- A (request its offsets), A <- item `it_next`
  __device__ __forceinline__ void shift(int64_t it_next) {
    itB = itA; xB0 = xA0; xB1 = xA1;
    if (itB < a.n_chunks) {
      if (xB1 > xB0) { vB0 = rowptr[xB0]; vB1 = rowptr[xB1 - 1]; vB2 = rowptr[xB1]; }
    } else if (itB < n_items) {
      vB0 = rowptr[xB0]; vB1 = rowptr[xB0 + 1];
    }
    loadA(it_next);
  }
  // current <- B, skipping empty light blocks
  __device__ __forceinline__ void take() {
    for (;;) {
      if (itB >= n_items) { have = false; return; }
      bool ok;
      if (itB < a.n_chunks) {
        r0 = xB0; r1 = xB1;
        ok = r1 > r0;
        if (ok) {
          const bool dec = (vB2 - vB1) > (OffT)kChunk;          // heavy last row: handled as segments
          if (dec) r1--;
          ok = r1 > r0;
          b = vB0; e = dec ? vB1 : vB2;
        }
      } else {
        r0 = (int32_t)(itB - a.n_chunks); r1 = -1;
        b = vB0 + (OffT)xB1 * (OffT)kSeg;
        e = (vB1 - b > (OffT)kSeg) ? b + (OffT)kSeg : vB1;
        ok = true;
      }
      shift(itA + nwarps);
      if (ok && e > b) { tb = b & ~(OffT)3; have = true; return; }
    }
  }
  __device__ __forceinline__ SpTrip<OffT> next() {
    SpTrip<OffT> t;
    t.tb = tb; t.b = b; t.e = e; t.r0 = r0; t.r1 = r1; t.flags = have ? 1 : 0;
    if (have) {
      tb += 128;
      if (tb >= e) { t.flags |= 2; take(); }
    }
    return t;
  }
};

template <typename OffT, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
spmv_pipe(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, GatherArgs a) {
  extern __shared__ __align__(16) float s_stage[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (THREADS / 32) + wib, nwarps = (int64_t)gridDim.x * (THREADS / 32);
  float *sv = s_stage + (size_t)wib * kCap;
  const uint64_t pol_x = l2_policy_evict_last(), pol_s = l2_policy_evict_first();
  SpGen<OffT> gen(rowptr, a, warp, nwarps);

  int4 q1; float4 x1;                 // trip n+1: column ids, matrix values
  float4 x0; float g0[4];             // trip n  : matrix values, gathered x
  float4 xm; float gm[4];             // trip n-1
  SpTrip<OffT> t1, t0, tm;
  float hacc = 0.f;
  t0.flags = 0; tm.flags = 0;
  x0 = make_float4(0.f, 0.f, 0.f, 0.f); xm = x0; x1 = x0; q1 = make_int4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < 4; u++) { g0[u] = 0.f; gm[u] = 0.f; }

  auto stream = [&](const SpTrip<OffT> &t) {
    if (!(t.flags & 1)) return;
    const uint64_t i = (uint64_t)t.tb + 4ull * lane;
    if (i < (uint64_t)t.e) {
      q1 = ld_stream_v4(reinterpret_cast<const int4 *>(col + i), pol_s);       // col has 256 B of slack past nnz
      if (i + 4 <= a.nnz) {
        const int4 r = ld_stream_v4(reinterpret_cast<const int4 *>(a.Ax + i), pol_s);
        x1 = make_float4(__int_as_float(r.x), __int_as_float(r.y), __int_as_float(r.z), __int_as_float(r.w));
      } else {
        x1.x = i < a.nnz ? a.Ax[i] : 0.f; x1.y = i + 1 < a.nnz ? a.Ax[i + 1] : 0.f;
        x1.z = i + 2 < a.nnz ? a.Ax[i + 2] : 0.f; x1.w = 0.f;
      }
    }
  };
  auto gather = [&](const SpTrip<OffT> &t, const int4 &q) {
    const OffT i = t.tb + (OffT)(4 * lane);
    const int c[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      g0[u] = 0.f;
      if ((t.flags & 1) && i + u >= t.b && i + u < t.e) g0[u] = ld_gather_f32(a.vec + c[u], pol_x);
    }
  };
  auto consume = [&](const SpTrip<OffT> &t) {
    if (!(t.flags & 1)) return;
    const OffT a0 = t.b & ~(OffT)3;
    const OffT i = t.tb + (OffT)(4 * lane);
    const float p0 = __fmul_rn(gm[0], xm.x), p1 = __fmul_rn(gm[1], xm.y), p2 = __fmul_rn(gm[2], xm.z), p3 = __fmul_rn(gm[3], xm.w);
    if (t.r1 >= 0) {
      if (i < t.e) *reinterpret_cast<float4 *>(sv + (i - a0)) = make_float4(p0, p1, p2, p3);
      if (t.flags & 2) {
        __syncwarp();
        for (int32_t r = t.r0 + lane; r < t.r1; r += 32) {
          const int32_t s = (int32_t)(rowptr[r] - a0), e = (int32_t)(rowptr[r + 1] - a0);
          float acc = __ldcs(a.y + r);
          for (int32_t j = s; j < e; j++) acc = __fadd_rn(acc, sv[j]);
          __stcs(a.y + r, acc);
        }
        __syncwarp();
      }
    } else {
      // entries outside [b, e) add +0.0f (their value slots hold a neighbouring row's entries)
      const float m0 = (i >= t.b && i < t.e) ? p0 : 0.f, m1 = (i + 1 >= t.b && i + 1 < t.e) ? p1 : 0.f;
      const float m2 = (i + 2 >= t.b && i + 2 < t.e) ? p2 : 0.f, m3 = (i + 3 >= t.b && i + 3 < t.e) ? p3 : 0.f;
      hacc += (m0 + m1) + (m2 + m3);
      if (t.flags & 2) {
        const float tot = warp_sum(hacc);
        if (lane == 0) a.heavy_partial[t.r0] = tot;
        hacc = 0.f;
      }
    }
  };

  t1 = gen.next();
  stream(t1);
  for (;;) {
    // rotate: n+1 -> n -> n-1
    tm = t0; xm = x0;
#pragma unroll
    for (int u = 0; u < 4; u++) gm[u] = g0[u];
    t0 = t1; x0 = x1;
    const int4 q0 = q1;
    t1 = gen.next();
    stream(t1);
    gather(t0, q0);
    consume(tm);
    if (!(t0.flags & 1) && !(tm.flags & 1)) break;
  }
}

// contrib[v] = scores[v] / out_degree(v)  (src/pr/omp_base.cc:24-25) for the local rows.
template <typename OffT>
__global__ void pr_init_contrib(const OffT *__restrict__ rowptr, const int32_t *__restrict__ out_degree,
                                const float *__restrict__ scores, float *__restrict__ contrib, int64_t rows,
                                int64_t row_lo) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t deg = out_degree ? out_degree[r] : (int32_t)(rowptr[r + 1] - rowptr[r]);
    contrib[row_lo + r] = __fdiv_rn(scores[r], (float)deg);
  }
}

// Fixed-order reduction of the per-warp partial L1 deltas -> err_trace[iter];
// raises the done flag when the total drops below eps (src/pr/omp_base.cc:36).
__global__ void __launch_bounds__(256)
pr_reduce_err(const double *__restrict__ partial, int n, double *err_trace, int iter, double eps, int32_t *done) {
  __shared__ double s[256];
  if (*done) return;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    err_trace[iter] = s[0];
    if (s[0] < eps) *done = iter + 1;
  }
}

// ---------------------------------------------------------------- schedule build
// chunk_row[k] = first row whose first non-zero is at or after k*kChunk.
template <typename OffT>
__global__ void build_chunk_rows(const OffT *__restrict__ rowptr, int64_t rows, int32_t n_chunks,
                                 int32_t *__restrict__ chunk_row) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_chunks) return;
  if (k == n_chunks) { chunk_row[k] = (int32_t)rows; return; }
  const OffT target = (OffT)k * kChunk;
  int64_t lo = 0, hi = rows;            // lower_bound over rowptr[0..rows]
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  chunk_row[k] = (int32_t)lo;
}

// counters: one 64-bit word, high 32 = heavy rows, low 32 = heavy segments, so
// that a row's slot and its first segment are allocated by ONE atomic.
template <typename OffT, bool FILL>
__global__ void scan_heavy(const OffT *__restrict__ rowptr, int64_t rows, unsigned long long *counter,
                           int32_t *heavy_row, int32_t *heavy_first, int2 *heavy_seg) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const OffT len = rowptr[r + 1] - rowptr[r];
    if (len > (OffT)kChunk) {
      const unsigned nseg = (unsigned)((len + kSeg - 1) / kSeg);
      const unsigned long long old = atomicAdd(counter, (1ull << 32) | nseg);
      if (FILL) {
        const int32_t slot = (int32_t)(old >> 32), first = (int32_t)(old & 0xffffffffu);
        heavy_row[slot] = (int32_t)r;
        heavy_first[slot] = first;
        for (unsigned s = 0; s < nseg; s++) heavy_seg[first + s] = make_int2((int)r, (int)s);
      }
    }
  }
}

static int dev_alloc(gdn_graph *g, void **p, size_t bytes) {
  GDN_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  g->device_bytes += bytes;
  return GDN_OK;
}

template <typename OffT>
static int build_schedule_t(gdn_graph *g, DevCsr &c) {
  cudaStream_t st = lib().stream;
  const uint64_t nch = c.nnz / kChunk + 1;
  if (nch > 0x7ffffff0ull) { set_error("too many row blocks"); return GDN_ERR_ARG; }
  c.n_chunks = (int32_t)nch;
  GDN_CHECK(dev_alloc(g, (void **)&c.chunk_row, sizeof(int32_t) * (nch + 1)));
  const OffT *rp = (const OffT *)c.rowptr;
  build_chunk_rows<OffT><<<(unsigned)((nch + 1 + 255) / 256), 256, 0, st>>>(rp, c.rows, c.n_chunks, c.chunk_row);
  unsigned long long *counter;
  GDN_CUDA(cudaMalloc((void **)&counter, 8));
  GDN_CUDA(cudaMemsetAsync(counter, 0, 8, st));
  const int grid = (int)std::min<int64_t>((c.rows + 255) / 256 + 1, 148 * 16);
  scan_heavy<OffT, false><<<grid, 256, 0, st>>>(rp, c.rows, counter, nullptr, nullptr, nullptr);
  unsigned long long h = 0;
  GDN_CUDA(cudaMemcpyAsync(&h, counter, 8, cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  c.n_heavy_rows = (int32_t)(h >> 32);
  c.n_heavy_segs = (int32_t)(h & 0xffffffffu);
  if (c.n_heavy_rows > 0) {
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_row, sizeof(int32_t) * c.n_heavy_rows));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_first, sizeof(int32_t) * c.n_heavy_rows));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_seg, sizeof(int2) * c.n_heavy_segs));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_partial, sizeof(float) * c.n_heavy_segs));
    GDN_CUDA(cudaMemsetAsync(counter, 0, 8, st));
    scan_heavy<OffT, true><<<grid, 256, 0, st>>>(rp, c.rows, counter, c.heavy_row, c.heavy_first, c.heavy_seg);
  }
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaFree(counter));
  GDN_CUDA(cudaGetLastError());
  return GDN_OK;
}

int build_schedule(gdn_graph *g, DevCsr &c) {
  return c.off64 ? build_schedule_t<uint64_t>(g, c) : build_schedule_t<uint32_t>(g, c);
}

// ---------------------------------------------------------------- launch helpers
static int gather_grid(const DevCsr &c) {
  const int64_t items = (int64_t)c.n_chunks + c.n_heavy_segs;
  const int64_t want = (items + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)lib().sm_count * 4 * 2;      // 2 waves of 4 resident CTAs per SM
  return (int)std::max<int64_t>(1, std::min(want, cap));
}
static int heavy_grid(const DevCsr &c) {
  const int64_t want = ((int64_t)c.n_heavy_rows + kWarps - 1) / kWarps;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)lib().sm_count * 4));
}

static void fill_sched(const DevCsr &c, GatherArgs &a) {
  a.chunk_row = c.chunk_row; a.n_chunks = c.n_chunks;
  a.heavy_seg = c.heavy_seg; a.n_heavy_segs = c.n_heavy_segs;
  a.heavy_partial = c.heavy_partial; a.heavy_row = c.heavy_row;
  a.heavy_first = c.heavy_first; a.n_heavy_rows = c.n_heavy_rows;
  a.rows = c.rows; a.nnz = c.nnz;
}

template <typename OffT>
static int spmv_t(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y, gdn_stats *st) {
  const DevCsr &c = g->in;
  GatherArgs a = {};
  fill_sched(c, a);
  a.Ax = d_Ax; a.vec = d_x; a.y = d_y;
  cudaStream_t s = lib().stream;
  const OffT *rp = (const OffT *)c.rowptr;
  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  // The TMA-staged variant is correct but measured 2.6x SLOWER than the register-staged kernel on urand-24
  // (8.4 vs 3.2 ms, profiles/r1_ncu_summary.md): 12 warps/SM of per-warp 1 K-entry items cannot cover the
  // describe -> copy -> gather -> row-sum chain.  Kept selectable for the next round's rework.
  const bool legacy = getenv("GDN_SPMV_TMA") == nullptr;
  const size_t tma_smem = sizeof(TmaStage) * 2 * kTmaWarps + sizeof(uint64_t) * 2 * kTmaWarps;
  // Column passes (experiment, GDN_SPMV_PASSES=P): pass p gathers only the columns of window p (a prefix / middle /
  // suffix of every sorted row) so that the window stays L2-resident; the other entries add +0.0f and y carries the
  // running sum from pass to pass -- the fp32 addition order of a light row is still exactly the reference's
  // (src/spmv/omp_base.cc:26-31).  Measured (profiles/r1_spmv_column_passes.txt): urand-24 3.16 -> 3.03 ms at P=2,
  // slower everywhere else (col/Ax are streamed once per pass): the kernel is not bound by the gather's L2 misses.
  // Default: one pass.
  const char *e_pass = getenv("GDN_SPMV_PASSES"), *e_win = getenv("GDN_SPMV_WINDOW_MB");
  const int64_t win_ids = (int64_t)(e_win ? atoi(e_win) : 40) * (1 << 20) / 4;
  int passes = e_pass ? atoi(e_pass) : (e_win ? (int)std::min<int64_t>(4, (g->m + win_ids - 1) / win_ids) : 1);
  if (passes < 1 || !legacy) passes = 1;
  a.col_lo = 0; a.col_hi = 0x7fffffff;
  // GDN_SPMV_PIPE: 0 = gather_kernel (one trip at a time); 10*T + C = spmv_pipe with T threads per CTA, C CTAs per SM
  const char *e_pipe = getenv("GDN_SPMV_PIPE");
  const int pipe = (legacy && passes == 1) ? (e_pipe ? atoi(e_pipe) : 3842) : 0;
  kev_begin();
  if (pipe) {
    // measured on urand-24 (profiles/r1_spmv_pipe_sweep.txt): 24 warps per SM at 76-80 registers win; 32 warps force 64
    // registers and spill (4.2 ms), 16 warps are short of gathers in flight (2.8 ms)
    void (*kern)(const OffT *, const int32_t *, GatherArgs) = spmv_pipe<OffT, 384, 2>;
    int threads = 384, ctas = 2;
    switch (pipe) {
      case 10241: kern = spmv_pipe<OffT, 1024, 1>; threads = 1024; ctas = 1; break;
      case 5122: kern = spmv_pipe<OffT, 512, 2>; threads = 512; ctas = 2; break;
      case 2563: kern = spmv_pipe<OffT, 256, 3>; threads = 256; ctas = 3; break;
      case 2562: kern = spmv_pipe<OffT, 256, 2>; threads = 256; ctas = 2; break;     // 128 registers
      case 2564: kern = spmv_pipe<OffT, 256, 4>; threads = 256; ctas = 4; break;
      default: break;
    }
    const size_t smem = sizeof(float) * (size_t)kCap * (threads / 32);
    GDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<lib().sm_count * ctas, threads, smem, s>>>(rp, c.col, a);
  } else if (legacy) {
    for (int p = 0; p < passes; p++) {
      a.col_lo = (int32_t)(g->m * p / passes);
      a.col_hi = p + 1 == passes ? 0x7fffffff : (int32_t)(g->m * (p + 1) / passes);
      gather_kernel<OffT, kModeSpmv><<<gather_grid(c), kThreads, 0, s>>>(rp, c.col, a);
      if (c.n_heavy_rows > 0 && p + 1 < passes) finalize_heavy<OffT, kModeSpmv><<<heavy_grid(c), kThreads, 0, s>>>(rp, a);
    }
  } else {
    GDN_CUDA(cudaFuncSetAttribute(spmv_tma_kernel<OffT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
    const int64_t items = (int64_t)c.n_chunks + c.n_heavy_segs;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((items + kTmaWarps - 1) / kTmaWarps, (int64_t)lib().sm_count * 3));
    spmv_tma_kernel<OffT><<<grid, kTmaWarps * 32, tma_smem, s>>>(rp, c.col, a);
  }
  kev_end();
  int launches = passes;
  if (c.n_heavy_rows > 0) {
    finalize_heavy<OffT, kModeSpmv><<<heavy_grid(c), kThreads, 0, s>>>(rp, a);
    launches += passes;
  }
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms; st->kernel_launches = launches; st->iterations = 1;
    kev_collect(st);
  }
  return GDN_OK;
}

int spmv_run(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y, gdn_stats *st) {
  if (((uintptr_t)d_Ax & 31) != 0) { set_error("d_Ax must be 32-byte aligned"); return GDN_ERR_ARG; }
  return g->in.off64 ? spmv_t<uint64_t>(g, d_Ax, d_x, d_y, st) : spmv_t<uint32_t>(g, d_Ax, d_x, d_y, st);
}

int pr_exchange(gdn_graph *g, float *contrib, double *err_slot);   // comm.cu: allgather over NVLink (no-op at 1 GPU)
int comm_size();
int64_t partition_width(int64_t m, int nparts);

template <typename OffT>
static int pr_t(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st) {
  const DevCsr &c = g->in;
  cudaStream_t s = lib().stream;
  const OffT *rp = (const OffT *)c.rowptr;
  const int ggrid = gather_grid(c), hgrid = heavy_grid(c);
  const int n_partial = ggrid * kWarps + (c.n_heavy_rows > 0 ? hgrid * kWarps : 0);
  if (!g->contrib[0]) {
    // full-length vector, padded so that every rank's allgather slice has equal width
    const int64_t len = std::max<int64_t>(g->m, partition_width(g->m, comm_size()) * comm_size());
    GDN_CHECK(dev_alloc(g, (void **)&g->contrib[0], sizeof(float) * len));
    GDN_CHECK(dev_alloc(g, (void **)&g->contrib[1], sizeof(float) * len));
    GDN_CHECK(dev_alloc(g, (void **)&g->err_trace, sizeof(double) * GDN_MAX_PR_ITER));
    GDN_CHECK(dev_alloc(g, (void **)&g->pr_done, sizeof(int32_t)));
  }
  if (g->n_err_partial < n_partial) {
    if (g->err_partial) GDN_CUDA(cudaFree(g->err_partial));
    GDN_CHECK(dev_alloc(g, (void **)&g->err_partial, sizeof(double) * n_partial));
    g->n_err_partial = n_partial;
  }
  if (max_iter > GDN_MAX_PR_ITER - 1) max_iter = GDN_MAX_PR_ITER - 1;
  GatherArgs a = {};
  fill_sched(c, a);
  a.scores = d_scores; a.out_degree = g->out_degree; a.row_lo = g->row_lo;
  a.base = (1.0f - damp) / (float)(int32_t)g->m;            // pr/omp_base.cc:16
  a.damp = damp; a.err_partial = g->err_partial; a.done = g->pr_done;
  double *h_err = (double *)lib().pinned;
  int64_t launches = 0;

  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  GDN_CUDA(cudaMemsetAsync(g->pr_done, 0, sizeof(int32_t), s));
  const int igrid = (int)std::min<int64_t>((c.rows + 255) / 256 + 1, (int64_t)lib().sm_count * 8);
  pr_init_contrib<OffT><<<igrid, 256, 0, s>>>(rp, g->out_degree, d_scores, g->contrib[0], c.rows, g->row_lo);
  launches++;
  GDN_CHECK(pr_exchange(g, g->contrib[0], nullptr));
  int iter, cur = 0;
  for (iter = 0; iter < max_iter; iter++) {
    a.vec = g->contrib[cur];
    a.contrib_out = g->contrib[cur ^ 1];
    a.err_slot0 = 0;
    kev_begin();
    gather_kernel<OffT, kModePr><<<ggrid, kThreads, 0, s>>>(rp, c.col, a);
    kev_end();
    launches++;
    if (c.n_heavy_rows > 0) {
      a.err_slot0 = ggrid * kWarps;
      finalize_heavy<OffT, kModePr><<<hgrid, kThreads, 0, s>>>(rp, a);
      launches++;
    }
    // multi-GPU: the stop test needs the all-reduced delta, so the host decides
    pr_reduce_err<<<1, 256, 0, s>>>(g->err_partial, n_partial, g->err_trace, iter, comm_size() > 1 ? -1.0 : eps, g->pr_done);
    launches++;
    GDN_CHECK(pr_exchange(g, g->contrib[cur ^ 1], g->err_trace + iter));
    GDN_CUDA(cudaMemcpyAsync(h_err, g->err_trace + iter, sizeof(double), cudaMemcpyDeviceToHost, s));
    GDN_CUDA(cudaStreamSynchronize(s));
    if (st) st->pr_err[iter] = *h_err;
    cur ^= 1;
    if (*h_err < eps) break;                                 // pr/omp_base.cc:36
  }
  GDN_CUDA(cudaEventRecord(lib().ev1, s));
  GDN_CUDA(cudaStreamSynchronize(s));
  GDN_CUDA(cudaGetLastError());
  if (st) {
    float ms = 0;
    GDN_CUDA(cudaEventElapsedTime(&ms, lib().ev0, lib().ev1));
    st->solve_ms = ms;
    st->kernel_launches = launches;
    st->iterations = iter + 1;                               // printf("iterations = %d", iter+1), :38
    kev_collect(st);
  }
  return GDN_OK;
}

int pr_run_sell(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st);   // pull.cu

int pr_run(gdn_graph *g, float *d_scores, float damp, double eps, int max_iter, gdn_stats *st) {
  if (g->pull.prepared && !getenv("GDN_PR_CSR")) return pr_run_sell(g, d_scores, damp, eps, max_iter, st);
  return g->in.off64 ? pr_t<uint64_t>(g, d_scores, damp, eps, max_iter, st)
                     : pr_t<uint32_t>(g, d_scores, damp, eps, max_iter, st);
}

}  // namespace gdn
