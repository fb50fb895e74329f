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

namespace gdn {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kOrdBlockGroups = 128;      // index groups (4 columns each) per block: 512 columns, 16 per lane
constexpr uint32_t kOrdOne = 0x1000000u;  // 2^24: the mantissa integer M of the accumulator lives in [2^23, 2^24)

// Accumulator bits usable by the integer path: normal, > 2^-103 (so that 1 / ulp is a normal float) and far from overflow.
__device__ __forceinline__ bool ord_acc_ok(uint32_t ab) { return ab >= 0x0C000000u && ab < 0x7f000000u; }
// 1 / ulp(acc) = 2^(23 - (e - 127)) for biased exponent e
__device__ __forceinline__ float ord_scale(uint32_t e) { return __uint_as_float((277u - e) << 23); }
// t = x / ulp lies exactly half-way between two integers (t < 2^23: above that a float has no fraction bits)
__device__ __forceinline__ bool ord_tie(float t) { return t < 8388608.f && t - floorf(t) == 0.5f; }

// One block, exactly: the accumulator (bits ab) after adding the 512 columns of a block in order; lane l holds columns
// 16 l .. 16 l + 15 in x[].  stage = 512 floats of shared memory owned by the warp (used only by the sequential fallback).
__device__ __forceinline__ uint32_t ordered_block(uint32_t ab, const float (&x)[16], int lane, float *stage) {
  // non-negative and finite <=> bits < 0x7f800000 as unsigned (-0.0 takes the sequential path too: harmless)
  uint32_t mbits = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) mbits = max(mbits, __float_as_uint(x[i]));
  const bool good = __all_sync(kFull, mbits < 0x7f800000u);
  int pos = 0;                                                             // columns of the block already in the accumulator
  while (pos < 512) {
    bool tie = false;                                                      // a half-way addend under the current ulp?
    if (good && ord_acc_ok(ab)) {
      const float sc = ord_scale(ab >> 23);
#pragma unroll
      for (int i = 0; i < 16; i++) tie |= lane * 16 + i >= pos && ord_tie(__fmul_rn(x[i], sc));
      tie = __any_sync(kFull, tie);
    }
    if (!good || !ord_acc_ok(ab) || tie) {
      // one true add at a time (every lane computes the same accumulator)
#pragma unroll
      for (int i = 0; i < 16; i++) stage[lane * 16 + i] = x[i];
      __syncwarp();
      float acc = __uint_as_float(ab);
      for (int i = pos; i < 512; i++) acc = __fadd_rn(acc, stage[i]);
      __syncwarp();
      return __float_as_uint(acc);
    }
    const uint32_t e = ab >> 23, M = (ab & 0x7fffffu) | 0x800000u;
    const float scale = ord_scale(e);
    // q = rne(x / ulp), clamped at 2^24 (a carry for sure); columns before `pos` are already in the accumulator
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
      run += lane * 16 + i >= pos ? __float2uint_rn(fminf(__fmul_rn(x[i], scale), 16777216.f)) : 0u;
    uint32_t incl = min(run, kOrdOne);                                     // (saturating: 32 x 16 x 2^24 would overflow)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl = min(incl + t, kOrdOne);
    }
    const unsigned carry = __ballot_sync(kFull, M + incl >= kOrdOne);
    if (!carry) {                                                          // the rest of the block stays inside the binade
      const uint32_t Mn = M + __shfl_sync(kFull, incl, 31);
      return (e << 23) | (Mn & 0x7fffffu);
    }
    const int L = __ffs(carry) - 1;
    uint32_t before = __shfl_up_sync(kFull, incl, 1);                      // sum of the lanes before this one ...
    if (lane == 0) before = 0;
    int sub = 16;
    float xc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {                                         // ... then column by column (meaningful in lane L)
      const uint32_t q = lane * 16 + i >= pos ? __float2uint_rn(fminf(__fmul_rn(x[i], scale), 16777216.f)) : 0u;
      if (sub == 16) {
        if (M + before + q >= kOrdOne) { sub = i; xc = x[i]; } else before += q;
      }
    }
    const int subL = __shfl_sync(kFull, sub, L);
    const uint32_t Mb = M + __shfl_sync(kFull, before, L);
    const float accb = __uint_as_float((e << 23) | (Mb & 0x7fffffu));
    ab = __float_as_uint(__fadd_rn(accb, __shfl_sync(kFull, xc, L)));       // the add that carries: rounded with the new ulp
    pos = L * 16 + subL + 1;
  }
  return ab;
}

// ------------------------------------------------------------------ layout of the exact rows (built by pull.cu exact_setup)
// Exact row j = (slice e, row r); its values are stored row-major, padded to whole blocks; per-block tables are indexed by
// blk = blk_base[e] + r * n_blocks(e) + b.
struct ExactArgs {
  int32_t n_exact;                 // exact slices
  int32_t n_blocks_total;
  const uint32_t *blk_base;        // [n_exact + 1] first block of slice e (32 rows x n_blocks(e) blocks each)
  float4 *vals;                    // [n_blocks_total * 128] row-major values, block after block
  double *S;                       // [n_blocks_total] real-valued sum of the block
  uint32_t *mx;                    // [n_blocks_total] largest addend (bits); >= 0x7f800000: a negative / non-finite addend
  uint8_t *plan;                   // [n_blocks_total] guessed biased exponent of the accumulator at the block; 0 = no fast path, 1 = all-zero block
  uint32_t *Q;                     // [n_blocks_total] sum of rne(x / ulp) under the guess
};

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

// Pass 2.  One warp per row: where will the accumulator be when block b starts?  Real-valued prefix P_b of the block sums;
// the block is planned for the integer path when [P_b, P_b + S_b] lies inside one binade with 2^-9 of margin on both sides
// (the rounded accumulator drifts from the real prefix), every addend is below 2^14 ulps (no clamp, no overflow of the
// 32-bit block sum) and non-negative.  Block 0 starts from zero: always careful.
__global__ void __launch_bounds__(256, 4)
pr_exact_plan(SellArgs a, ExactArgs x) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * 256 + threadIdx.x) >> 5;                    // exact row index: slice row / 32, row % 32
  if (row >= x.n_exact * 32) return;
  const int e = row >> 5, r = row & 31;
  const uint32_t nb = (x.blk_base[e + 1] - x.blk_base[e]) >> 5;
  const uint32_t k0 = x.blk_base[e] + (uint32_t)r * nb;
  double base = 0.0;
  for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
    const uint32_t b = b0 + lane;
    const double s = b < nb ? x.S[k0 + b] : 0.0;
    double incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    const double P = base + incl - s;                                      // real prefix at the start of block b
    if (b < nb) {
      uint8_t plan = 0;
      if (x.mx[k0 + b] == 0u) plan = 1;                                   // nothing but +0.0f (padding past the end of the row): adds nothing
      const float pf = (float)P, qf = (float)(P + s);
      const uint32_t pb = __float_as_uint(pf), qb = __float_as_uint(qf);
      const uint32_t ex = pb >> 23;
      if (b > 0 && x.mx[k0 + b] < 0x7f800000u && ord_acc_ok(pb) && (qb >> 23) == ex &&
          (pb & 0x7fffffu) > 0x4000u && (qb & 0x7fffffu) < 0x7fc000u &&       // 2^-9 away from both ends of the binade
          __fmul_rn(__uint_as_float(x.mx[k0 + b]), ord_scale(ex)) < 16384.f && plan == 0)
        plan = (uint8_t)ex;
      x.plan[k0 + b] = plan;
    }
    base += __shfl_sync(kFull, incl, 31);
  }
}

// Pass 3.  One warp per planned block: Q_b = sum of rne(x / ulp) for the planned binade; a block with a half-way addend
// loses its plan (pass 4 then adds it sequentially).
__global__ void __launch_bounds__(256, 4)
pr_exact_qsum(SellArgs a, ExactArgs x) {
  if (*a.done) return;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
  for (uint32_t k = warp; k < (uint32_t)x.n_blocks_total; k += nwarps) {
    const uint32_t ex = x.plan[k];
    if (ex <= 1) continue;                                                 // warp-uniform: careful block / all-zero block
    const float scale = ord_scale(ex);
    const float4 *src = x.vals + (size_t)k * kOrdBlockGroups + lane * 4;
    uint32_t run = 0;
    bool tie = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 v = __ldcs(src + i);
      const float t[4] = {__fmul_rn(v.x, scale), __fmul_rn(v.y, scale), __fmul_rn(v.z, scale), __fmul_rn(v.w, scale)};
#pragma unroll
      for (int u = 0; u < 4; u++) { run += __float2uint_rn(t[u]); tie |= ord_tie(t[u]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) run += __shfl_xor_sync(kFull, run, o);
    tie = __any_sync(kFull, tie);
    if (lane == 0) {
      x.Q[k] = run;
      if (tie) x.plan[k] = 0;
    }
  }
}

// Pass 4.  One warp per row, blocks in order with the TRUE accumulator; then the row epilogue.
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
    const uint32_t k0 = x.blk_base[e] + (uint32_t)r * nb;
    uint32_t ab = 0;                                                       // bits of the accumulator
    // the plans and integer sums of 32 blocks at a time, one per lane, requested one run ahead
    uint32_t plan_n = lane < nb ? x.plan[k0 + lane] : 0u, q_n = lane < nb ? x.Q[k0 + lane] : 0u;
    for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
      const uint32_t plan_l = plan_n, q_l = plan_n > 1u ? q_n : 0u;
      const bool bl_past = b0 + lane >= nb;
      const uint32_t bn = b0 + 32 + lane;
      plan_n = bn < nb ? x.plan[k0 + bn] : 0u;
      q_n = bn < nb ? x.Q[k0 + bn] : 0u;
      const int n = (int)min(32u, nb - b0);
      // all 32 planned for the binade we are in and no carry over the whole run: one add
      uint32_t qs = q_l;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) qs += __shfl_xor_sync(kFull, qs, o);
      const uint32_t M0 = (ab & 0x7fffffu) | 0x800000u;
      if (__all_sync(kFull, plan_l == 1u || bl_past)) continue;             // a run of all-zero blocks
      if (n == 32 && __all_sync(kFull, plan_l == (ab >> 23) || plan_l == 1u) && ord_acc_ok(ab) && qs < kOrdOne && M0 + qs < kOrdOne) {
        ab = (ab & 0xff800000u) | ((M0 + qs) & 0x7fffffu);
        continue;
      }
      for (int i = 0; i < n; i++) {
        const uint32_t p = __shfl_sync(kFull, plan_l, i), q = __shfl_sync(kFull, q_l, i);
        const uint32_t M = (ab & 0x7fffffu) | 0x800000u;
        if (p == 1u) continue;
        if (p != 0 && p == (ab >> 23) && q < kOrdOne && M + q < kOrdOne) {
          ab = (ab & 0xff800000u) | ((M + q) & 0x7fffffu);
        } else {
          const float4 *src = x.vals + (size_t)(k0 + b0 + i) * kOrdBlockGroups + lane * 4;
          float v[16];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const float4 t = __ldcs(src + u);
            v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
          }
          ab = ordered_block(ab, v, lane, s_stage[wib]);
        }
      }
    }
    if (lane == 0) pr_epilogue(a, (int64_t)row, __uint_as_float(ab), err);
  }
  err = warp_sum(err);
  if (lane == 0) a.err_partial[a.err_slot0 + warp] = err;
  contrib_flush(a);
}

}  // namespace gdn
