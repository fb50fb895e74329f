// ordered_sum.cuh -- the column-order fp32 sum of a long row WITHOUT its latency chain (pull.cu, exact slices).
//
// A sequential fp32 sum (src/pr/omp_base.cc:28-30) of 10^6 addends is a chain of 10^6 dependent adds: 6.7 cycles each
// beside its loads (tools/fadd_chain_microbench.cu), 3.4 ms for the longest row of Kronecker scale 26 -- the whole
// iteration.  But the ROUNDING of a sequential sum of non-negative numbers is almost order-free: while the accumulator
// stays inside one binade [2^k, 2^(k+1)) every add rounds its addend to a multiple of ulp = 2^(k-23) and adds it exactly,
//        acc_after = (M + sum_i rne(x_i / ulp)) * ulp            (M = acc / ulp, a 24-bit integer)
// whatever the order.  So a row is cut into blocks of 512 columns and summed in four small passes:
//   1  pr_exact_gather    one warp per block: the values, row-major, and the block's real-valued sum S_b
//   2  pr_exact_plan      one warp per row: prefix of S -> the binade the accumulator will be in at each block (a GUESS:
//                         the rounded accumulator drifts from the real prefix by ~sqrt(n) ulps) and whether the block
//                         lies safely inside it
//   3  pr_exact_qsum      one warp per safe block: Q_b = sum_i rne(x_i / ulp) for the guessed binade -- all blocks in parallel
//   4  pr_exact_combine   one warp per row walks its blocks IN ORDER with the true (exponent, M): a block whose guess
//                         holds and whose Q_b does not carry is one integer add; every other block (the first one, the
//                         ~25 that carry into the next binade, a wrong guess, a negative or non-finite addend) is redone
//                         by ordered_block below, which emulates the adds of that block exactly -- so a wrong guess
//                         costs time, never a bit.
// One case is NOT order-free: an addend EXACTLY half-way between two multiples of ulp is rounded by the hardware to the
// even ACCUMULATOR, not to the even addend (probability 2^-d per add, d = exponent gap between accumulator and addend:
// a few adds per long row, nearly every add of its first block).  A block that holds such a tie is summed by true
// sequential adds (pass 3 un-plans it, ordered_block takes its fallback), so the result IS the sequential sum, bit for
// bit.  The critical path of a 10^6-entry row is ~2000 integer adds plus ~40 careful blocks, ~10 of them sequential.
#pragma once
#include "pull.cuh"
#include "ordered_core.cuh"

namespace gdn {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// block k -> (slice, row, block in row)
__device__ __forceinline__ void exact_locate(const ExactArgs &x, uint32_t k, int &e, int &r, uint32_t &b, uint32_t &nb) {
  int lo = 0, hi = x.n_exact;                                              // last e with blk_base[e] <= k
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x.blk_base[mid] <= k) lo = mid; else hi = mid; }
  e = lo;
  nb = (x.blk_base[e + 1] - x.blk_base[e]) >> 5;
  const uint32_t off = k - x.blk_base[e];
  r = (int)(off / nb);
  b = off - (uint32_t)r * nb;
}

// Pass 1.  One warp per block: lane l takes index groups 4 l .. 4 l + 3 of the block (16 columns of the row, strided in
// the lane = row layout of the slice), gathers, writes the 64 bytes of its values row-major, and the warp leaves S_b, max.
// (Its time is the gathers: the ids of a hub row are mostly cold, one HBM sector each.  Reading the tile the way it lies and
// transposing through shared memory was measured slower -- 0.97 vs 0.52 ms at Kronecker scale 26 -- for lack of loads in flight.)
__global__ void __launch_bounds__(256, 4)
pr_exact_gather(SellArgs a, ExactArgs x) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const uint64_t pol = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
  for (uint32_t k = warp; k < (uint32_t)x.n_blocks_total; k += nwarps) {
    int e, r; uint32_t b, nb;
    exact_locate(x, k, e, r, b, nb);
    // blocks past the end of THIS row (the slice is as wide as its longest row) hold padding only: their values were
    // zeroed once (exact_setup), their sum is zero
    const uint32_t row_groups = ((uint32_t)a.sdeg[e * 32 + r] + 3) >> 2;
    if (b * kOrdBlockGroups >= row_groups) {
      if (lane == 0) { x.S[k] = 0.0; x.mx[k] = 0u; }
      continue;
    }
    const uint32_t g0 = a.slice_ptr[e], ngl = (a.slice_ptr[e + 1] - g0) >> 5;
    double s = 0.0;
    uint32_t mbits = 0;
    float4 *dst = x.vals + (size_t)k * kOrdBlockGroups + lane * 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t q = b * kOrdBlockGroups + lane * 4 + i;
      int4 c = make_int4(-1, -1, -1, -1);
      if (q < ngl) c = ld_stream_v4(a.sell + g0 + (size_t)q * 32 + r, pol);
      const int id[4] = {c.x, c.y, c.z, c.w};
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        v[u] = 0.f;                                                        // padding adds +0.0f: leaves an fp32 sum unchanged
        if (id[u] >= 0) {
          const float *p = a.contrib_in + id[u];
          if (tier_id(a, id[u]) < a.warm) v[u] = ld_gather_f32(p, pol_last);
          else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[u]) : "l"(p), "l"(pol));
        }
        s += (double)v[u];
        mbits = max(mbits, __float_as_uint(v[u]));
      }
      dst[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mbits = max(mbits, __shfl_xor_sync(kFull, mbits, o));
    if (lane == 0) { x.S[k] = s; x.mx[k] = mbits; }
  }
}

// Pass 2 for the exact slices of PageRank (rows start from zero).
__global__ void __launch_bounds__(256, 4)
pr_exact_plan(SellArgs a, ExactArgs x) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;                    // exact row index: slice row / 32, row % 32
  if (row >= x.n_exact * 32) return;
  const int e = row >> 5, r = row & 31;
  const uint32_t nb = (x.blk_base[e + 1] - x.blk_base[e]) >> 5;
  exact_plan_row(x, x.blk_base[e] + (uint32_t)r * nb, nb, 0.0, lane);
}

// ... then the PageRank row epilogue.
__global__ void __launch_bounds__(256, 4)
pr_exact_combine(SellArgs a, ExactArgs x) {
  __shared__ float s_stage[8][512];
  if (*a.done) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  double err = 0.0;
  const int row = (int)warp;
  if (row < x.n_exact * 32 && (int64_t)row < a.n_nz_rows) {
    const int e = row >> 5, r = row & 31;
    const uint32_t nb = (x.blk_base[e + 1] - x.blk_base[e]) >> 5;
    const uint32_t ab = exact_combine_row(x, x.blk_base[e] + (uint32_t)r * nb, nb, 0u, lane, s_stage[wib]);
    if (lane == 0) pr_epilogue(a, (int64_t)row, __uint_as_float(ab), err);
  }
  err = warp_sum(err);
  if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  contrib_flush(a);
}

}  // namespace gdn
