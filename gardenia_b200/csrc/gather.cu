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
#include "ordered_core.cuh"
#include <chrono>
#include <vector>
#include <cstdlib>

namespace gdn {

constexpr int kChunk = 512;            // non-zeros per light block (target)
constexpr int kCap = 2 * kChunk + 8;   // staged products per warp
constexpr int kSeg = 1024;             // non-zeros per heavy-row segment
constexpr int kWarps = 8;              // warps per CTA
// SpMV rows longer than this are summed in the reference's order (SpmvExact below); GDN_SPMV_EXACT_LEN overrides it when a
// graph is created (0 = never)
constexpr int64_t kSpmvExactLen = 8192;
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
      const bool ok = (i + j >= b) && (i + j < e);
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

// counters: one 64-bit word per class, high 32 = heavy rows, low 32 = heavy segments, so that a row's slot and its first
// segment are allocated by ONE atomic.  Class 1 = rows longer than xlen (listed after all rows of class 0).
template <typename OffT, bool FILL>
__global__ void scan_heavy(const OffT *__restrict__ rowptr, int64_t rows, unsigned long long *counter, OffT xlen, int32_t rows_lt,
                           int32_t segs_lt, int32_t *heavy_row, int32_t *heavy_first, int2 *heavy_seg) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const OffT len = rowptr[r + 1] - rowptr[r];
    if (len > (OffT)kChunk) {
      const int cls = len > xlen ? 1 : 0;
      const unsigned nseg = (unsigned)((len + kSeg - 1) / kSeg);
      const unsigned long long old = atomicAdd(counter + cls, (1ull << 32) | nseg);
      if (FILL) {
        const int32_t slot = (int32_t)(old >> 32) + (cls ? rows_lt : 0), first = (int32_t)(old & 0xffffffffu) + (cls ? segs_lt : 0);
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
  GDN_CUDA(cudaMalloc((void **)&counter, 16));
  GDN_CUDA(cudaMemsetAsync(counter, 0, 16, st));
  const int grid = (int)std::min<int64_t>((c.rows + 255) / 256 + 1, 148 * 16);
  const char *xe = getenv("GDN_SPMV_EXACT_LEN");
  const int64_t xl = xe ? atoll(xe) : kSpmvExactLen;
  const OffT xlen = xl <= 0 ? (OffT)~(OffT)0 : (OffT)std::max<int64_t>(xl, kChunk);
  scan_heavy<OffT, false><<<grid, 256, 0, st>>>(rp, c.rows, counter, xlen, 0, 0, nullptr, nullptr, nullptr);
  unsigned long long h[2] = {0, 0};
  GDN_CUDA(cudaMemcpyAsync(h, counter, 16, cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  c.n_heavy_rows_lt = (int32_t)(h[0] >> 32);
  c.n_heavy_segs_lt = (int32_t)(h[0] & 0xffffffffu);
  c.n_heavy_rows = c.n_heavy_rows_lt + (int32_t)(h[1] >> 32);
  c.n_heavy_segs = c.n_heavy_segs_lt + (int32_t)(h[1] & 0xffffffffu);
  if (c.n_heavy_rows > 0) {
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_row, sizeof(int32_t) * c.n_heavy_rows));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_first, sizeof(int32_t) * c.n_heavy_rows));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_seg, sizeof(int2) * c.n_heavy_segs));
    GDN_CHECK(dev_alloc(g, (void **)&c.heavy_partial, sizeof(float) * c.n_heavy_segs));
    GDN_CUDA(cudaMemsetAsync(counter, 0, 16, st));
    scan_heavy<OffT, true><<<grid, 256, 0, st>>>(rp, c.rows, counter, xlen, c.n_heavy_rows_lt, c.n_heavy_segs_lt, c.heavy_row, c.heavy_first, c.heavy_seg);
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

// ---- hot-first columns for SpMV on skewed graphs whose vector does not fit L2 (Kronecker scale >= 25) --------------
// With the generator's random ids every x[Aj] of such a graph is an HBM sector miss: 19.8 ms per SpMV at Kronecker scale
// 26 where PageRank's iteration over the same edges takes 4.5.  The degree ranking that the PageRank layout already holds
// (newid: hottest vertex first) is applied to the COLUMN ids only -- the rows, their order of addition and Ax stay as they
// are, so every y[r] keeps its bits -- and x is scattered into that order at the start of each call (0.4 ms): three
// quarters of the gathers then land in a dozen L2-resident megabytes.  Built once per resident graph (4 bytes per edge).
__global__ void __launch_bounds__(256)
spmv_renumber_cols(const int32_t *__restrict__ col, const int32_t *__restrict__ newid, int32_t *__restrict__ out, uint64_t nnz) {
  for (uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x; e < nnz; e += (uint64_t)gridDim.x * 256) out[e] = newid[col[e]];
}
__global__ void __launch_bounds__(256)
spmv_scatter_x(const float *__restrict__ x, const int32_t *__restrict__ newid, float *__restrict__ xp, int64_t m) {
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < m; v += (int64_t)gridDim.x * 256) xp[__ldcs(newid + v)] = __ldcs(x + v);
}
double pull_hot_share(const PullLayout &L);          // pull.cu

static int spmv_hot_columns(gdn_graph *g) {
  if (g->spmv_tried) return GDN_OK;
  g->spmv_tried = true;
  const PullLayout &L = g->pull;
  const char *e = getenv("GDN_SPMV_HOT");           // 0: never, 1: whenever the ranking exists
  const int force = e ? atoi(e) : -1;
  if (force == 0 || g->one_shot || !L.prepared || L.P != 1 || !L.newid || !L.symmetric_order || L.rows != g->m) return GDN_OK;
  if (force < 0 && (g->m * 4 <= ((int64_t)48 << 20) || pull_hot_share(L) < 0.4)) return GDN_OK;
  const DevCsr &c = g->in;
  const auto t0 = std::chrono::steady_clock::now();
  int32_t *colr = nullptr;
  float *xp = nullptr;
  if (cudaMalloc((void **)&colr, sizeof(int32_t) * c.nnz + 256) != cudaSuccess || cudaMalloc((void **)&xp, sizeof(float) * (g->m + 64)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(colr);
    return GDN_OK;                                   // an optimisation: without the memory the plain column ids do
  }
  GDN_CUDA(cudaMemsetAsync((char *)colr + sizeof(int32_t) * c.nnz, 0, 256, lib().stream));
  spmv_renumber_cols<<<lib().sm_count * 16, 256, 0, lib().stream>>>(c.col, L.newid, colr, c.nnz);
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  GDN_CUDA(cudaGetLastError());
  g->spmv_col = colr; g->spmv_x = xp;
  g->device_bytes += sizeof(int32_t) * c.nnz + sizeof(float) * g->m;
  g->spmv_prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return GDN_OK;
}

// ---- rows longer than kSpmvExactLen: summed in the reference's order ------------------------------------------------
// A sequential fp32 sum (src/spmv/omp_base.cc:27-31) of n addends drifts from the exact sum by an amount that grows with n
// (measured against spmv_omp_base: 1.2e-5 relative on the longest row of Kronecker scale 22, 1.4e-4 at scale 26, where
// segment partials + a tree stay within 1e-7) -- so to meet the reference within 1e-5 per row the long rows have to be
// ROUNDED like the reference, i.e. summed in its order.  That is the ordered sum of ordered_core.cuh: products of a row
// written row-major in 512-entry blocks (pass 1, the only pass that touches the matrix), per-block plans and integer sums,
// and one warp per row walking its blocks in order from y[row].  Bit-identical to the reference for these rows.
template <typename OffT>
__global__ void __launch_bounds__(256, 4)
spmv_exact_gather(const OffT *__restrict__ rowptr, const int32_t *__restrict__ col, const float *__restrict__ Ax,
                  const float *__restrict__ vec, const int32_t *__restrict__ xrow, ExactArgs x) {
  const int lane = threadIdx.x & 31;
  const uint64_t pol_x = l2_policy_evict_last();
  const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
  for (uint32_t k = warp; k < (uint32_t)x.n_blocks_total; k += nwarps) {
    int lo = 0, hi = x.n_exact;                                            // last row i with blk_base[i] <= k
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x.blk_base[mid] <= k) lo = mid; else hi = mid; }
    const int32_t row = xrow[lo];
    const OffT rb = rowptr[row] + (OffT)(k - x.blk_base[lo]) * 512u, re = rowptr[row + 1];
    // lane l owns entries 16 l .. 16 l + 15 of the block (the layout ordered_block wants): 64 contiguous bytes per lane
    int32_t c[16];
    float w[16], v[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const OffT j = rb + (OffT)(lane * 16 + u);
      c[u] = j < re ? __ldg(col + j) : -1;
      w[u] = j < re ? __ldg(Ax + j) : 0.f;
    }
    double s = 0.0;
    uint32_t mbits = 0;
#pragma unroll
    for (int u = 0; u < 16; u++) {
      v[u] = c[u] >= 0 ? __fmul_rn(ld_gather_f32(vec + c[u], pol_x), w[u]) : 0.f;   // x[j] * Ax[jj], rounded like the reference's
      s += (double)v[u];
      mbits = max(mbits, __float_as_uint(v[u]));
    }
    float4 *dst = x.vals + (size_t)k * kOrdBlockGroups + lane * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mbits = max(mbits, __shfl_xor_sync(kFull, mbits, o));
    if (lane == 0) { x.S[k] = s; x.mx[k] = mbits; }
  }
}

__global__ void __launch_bounds__(256, 4)
spmv_exact_plan(const int32_t *__restrict__ xrow, const float *__restrict__ y, ExactArgs x) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (i >= x.n_exact) return;
  exact_plan_row(x, x.blk_base[i], x.blk_base[i + 1] - x.blk_base[i], (double)y[xrow[i]], lane);
}

__global__ void __launch_bounds__(256, 4)
spmv_exact_combine(const int32_t *__restrict__ xrow, float *__restrict__ y, ExactArgs x) {
  __shared__ float s_stage[8][512];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (i >= x.n_exact) return;
  const int32_t row = xrow[i];
  const uint32_t ab = exact_combine_row(x, x.blk_base[i], x.blk_base[i + 1] - x.blk_base[i], __float_as_uint(y[row]), lane, s_stage[wib]);
  if (lane == 0) y[row] = __uint_as_float(ab);
}

template <typename OffT>
__global__ void spmv_exact_blocks(const OffT *__restrict__ rowptr, const int32_t *__restrict__ xrow, int32_t n, uint32_t *__restrict__ nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) nb[i] = (uint32_t)((rowptr[xrow[i] + 1] - rowptr[xrow[i]] + 511) / 512);
}

// Tables of the exact rows, once per graph.  An optimisation for parity only: if the memory is not there the long rows
// stay with the segment sums (n_heavy_*_lt are reset to the full list).
template <typename OffT>
static int spmv_exact_setup(gdn_graph *g) {
  auto &X = g->spmv_exact;
  DevCsr &c = g->in;
  if (X.tried) return GDN_OK;
  X.tried = true;
  const int32_t n = c.n_heavy_rows - c.n_heavy_rows_lt;
  if (n <= 0) return GDN_OK;
  cudaStream_t st = lib().stream;
  const int32_t *xrow = c.heavy_row + c.n_heavy_rows_lt;
  auto give_up = [&]() {
    cudaGetLastError();
    cudaFree(X.blk_base); cudaFree(X.vals); cudaFree(X.S); cudaFree(X.mx); cudaFree(X.Q); cudaFree(X.plan);
    X = gdn_graph::SpmvExact();
    X.tried = true;
    c.n_heavy_rows_lt = c.n_heavy_rows; c.n_heavy_segs_lt = c.n_heavy_segs;
    return GDN_OK;
  };
  if (cudaMalloc((void **)&X.blk_base, sizeof(uint32_t) * ((size_t)n + 1)) != cudaSuccess) return give_up();
  spmv_exact_blocks<OffT><<<(n + 255) / 256, 256, 0, st>>>((const OffT *)c.rowptr, xrow, n, X.blk_base);
  std::vector<uint32_t> nb((size_t)n + 1);
  GDN_CUDA(cudaMemcpyAsync(nb.data(), X.blk_base, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  uint64_t tot = 0;
  for (int32_t i = 0; i < n; i++) { const uint32_t v = nb[i]; nb[i] = (uint32_t)tot; tot += v; }
  nb[n] = (uint32_t)tot;
  if (tot >= 0x7fffffffull) return give_up();
  GDN_CUDA(cudaMemcpyAsync(X.blk_base, nb.data(), sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  if (cudaMalloc((void **)&X.vals, sizeof(float) * 512 * (size_t)tot) != cudaSuccess || cudaMalloc((void **)&X.S, sizeof(double) * (size_t)tot) != cudaSuccess ||
      cudaMalloc((void **)&X.mx, sizeof(uint32_t) * (size_t)tot) != cudaSuccess || cudaMalloc((void **)&X.Q, sizeof(uint32_t) * (size_t)tot) != cudaSuccess ||
      cudaMalloc((void **)&X.plan, (size_t)tot + 16) != cudaSuccess)
    return give_up();
  X.n_rows = n; X.n_blocks = (uint32_t)tot;
  g->device_bytes += (sizeof(float) * 512 + 17) * (size_t)tot;
  return GDN_OK;
}

template <typename OffT>
static int spmv_t(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y, gdn_stats *st) {
  const DevCsr &c = g->in;
  GDN_CHECK(spmv_hot_columns(g));
  GDN_CHECK(spmv_exact_setup<OffT>(g));
  GatherArgs a = {};
  fill_sched(c, a);
  a.n_heavy_rows = c.n_heavy_rows_lt; a.n_heavy_segs = c.n_heavy_segs_lt;      // the longer rows: exact passes below
  a.Ax = d_Ax; a.vec = d_x; a.y = d_y;
  cudaStream_t s = lib().stream;
  const OffT *rp = (const OffT *)c.rowptr;
  const int32_t *col = c.col;
  kev_reset();
  GDN_CUDA(cudaEventRecord(lib().ev0, s));
  int launches = 0;
  if (g->spmv_col) {
    spmv_scatter_x<<<lib().sm_count * 8, 256, 0, s>>>(d_x, g->pull.newid, g->spmv_x, g->m);
    a.vec = g->spmv_x;
    col = g->spmv_col;
    launches++;
  }
  // 24 warps per SM at 76-80 registers measured best on urand-24 (profiles/r1_spmv_pipe_sweep.txt): 32 warps force 64
  // registers and spill (4.2 ms), 16 warps are short of gathers in flight (2.8 ms).  Negative results kept in profiles/:
  // a TMA-staged index/value stream (cp.async.bulk per 1 K-entry item: 2.6x slower), column passes over an L2-sized
  // window (r1_spmv_column_passes.txt), the one-trip-at-a-time kernel (+18 %).
  constexpr int kPipeThreads = 384, kPipeCtas = 2;
  const size_t smem = sizeof(float) * (size_t)kCap * (kPipeThreads / 32);
  GDN_CUDA(cudaFuncSetAttribute(spmv_pipe<OffT, kPipeThreads, kPipeCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (an L2 bulk prefetch of the next items' col / Ax, the trick that pays in pr_band_kernel, was measured 18 % SLOWER here:
  // 2.82 -> 3.33 ms on urand-24 -- this kernel is bound by the sector rate of its gathers, not by stream latency)
  kev_begin();
  spmv_pipe<OffT, kPipeThreads, kPipeCtas><<<lib().sm_count * kPipeCtas, kPipeThreads, smem, s>>>(rp, col, a);
  kev_end();
  launches++;
  if (a.n_heavy_rows > 0) {
    finalize_heavy<OffT, kModeSpmv><<<heavy_grid(c), kThreads, 0, s>>>(rp, a);
    launches++;
  }
  if (g->spmv_exact.n_rows > 0) {
    const auto &X = g->spmv_exact;
    ExactArgs xa;
    xa.n_exact = X.n_rows; xa.n_blocks_total = (int32_t)X.n_blocks; xa.blk_base = X.blk_base; xa.vals = X.vals;
    xa.S = X.S; xa.mx = X.mx; xa.plan = X.plan; xa.Q = X.Q;
    const int32_t *xrow = c.heavy_row + c.n_heavy_rows_lt;
    const int sm = lib().sm_count, rgrid = (X.n_rows + 7) / 8;
    spmv_exact_gather<OffT><<<sm * 8, 256, 0, s>>>(rp, col, d_Ax, a.vec, xrow, xa);
    spmv_exact_plan<<<rgrid, 256, 0, s>>>(xrow, d_y, xa);
    exact_qsum<<<sm * 8, 256, 0, s>>>(nullptr, xa);
    spmv_exact_combine<<<rgrid, 256, 0, s>>>(xrow, d_y, xa);
    launches += 4;
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
