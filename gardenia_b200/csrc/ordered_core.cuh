// ordered_core.cuh -- the building blocks of the ordered sum (see ordered_sum.cuh for the method): the exact per-block
// emulation, the per-block tables and the row-level plan / integer-sum / in-order-combine steps.  Shared by the exact
// slices of PageRank (pull.cu via ordered_sum.cuh) and the long rows of SpMV (gather.cu).
#pragma once
#include "common.cuh"

namespace gdn {

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

// Pass 2.  One warp per row: where will the accumulator be when block b starts?  Real-valued prefix P_b of the block sums
// (from `start`, the value the accumulator holds before the row: 0 for PageRank, y[row] for SpMV);
// the block is planned for the integer path when [P_b, P_b + S_b] lies inside one binade with 2^-9 of margin on both sides
// (the rounded accumulator drifts from the real prefix), every addend is below 2^14 ulps (no clamp, no overflow of the
// 32-bit block sum) and non-negative.  Block 0 is always careful.
__device__ __forceinline__ void exact_plan_row(const ExactArgs &x, uint32_t k0, uint32_t nb, double start, int lane) {
  double base = start;
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
static __global__ void __launch_bounds__(256, 4)
exact_qsum(const int32_t *__restrict__ done, ExactArgs x) {
  if (done && *done) return;
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

// Pass 4.  One warp per row, blocks in order with the TRUE accumulator (bits ab on entry: what the row's sum starts from).
__device__ __forceinline__ uint32_t exact_combine_row(const ExactArgs &x, uint32_t k0, uint32_t nb, uint32_t ab, int lane, float *stage) {
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
        ab = ordered_block(ab, v, lane, stage);
      }
    }
  }
  return ab;
}

}  // namespace gdn
