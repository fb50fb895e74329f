// ordered_sum.cuh -- the column-order fp32 sum of a long row without its latency chain (pull.cu, exact slices).
#pragma once
#include "pull.cuh"

namespace gdn {

constexpr int kChainHot = 16384;          // hot-table entries of a CTA that first adds rows of exact slices (64 KB)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ exact slices, default mode: ordered sum by emulation
// The chain above is bound by the latency of a dependent fp32 add (6.7 cycles beside its shared-memory loads, measured:
// tools/fadd_chain_microbench.cu): 4.2 ms for the 1.0 M-entry row of Kronecker scale 26 -- the whole iteration.  But the
// ROUNDING of a sequential sum of non-negative numbers is almost order-free: while the accumulator stays inside one
// binade [2^k, 2^(k+1)) every add rounds its addend to a multiple of ulp = 2^(k-23) and adds it exactly, so
//        acc_after = (M + sum_i rne(x_i / ulp)) * ulp            (M = acc / ulp, a 24-bit integer)
// whatever the order -- an integer prefix sum, parallel over the lanes of a warp.  Only the ~25 adds of a row that carry the
// accumulator into the next binade are done as true fp32 adds (they round with the new ulp), after which the rest of the
// block is re-scaled.  The result equals src/pr/omp_base.cc:28-30 bit for bit except where an addend falls EXACTLY
// half-way between two multiples of ulp with M odd (round-half-even looks at M: probability 2^-19 per add, one ulp of
// the accumulator each) -- against sqrt(n) ulps for a re-ordered sum.  Blocks with a negative / non-finite addend or a
// tiny accumulator (the first 256 columns, arbitrary start vectors) are added sequentially.
constexpr int kOrdDepth = 8;              // blocks a warp keeps in flight (cp.async): 8 x 512 columns (its copies queue behind the gathers of the other warps)
constexpr int kOrdRows = 8;               // rows of an exact slice per CTA (its first 8 warps; the others start on the queue at once)
constexpr int kOrdSplit = 32 / kOrdRows;  // CTAs that share one exact slice
constexpr size_t kOrdSmem = (size_t)kChainHot * sizeof(float) + (size_t)kOrdRows * kOrdDepth * 32 * 4 * sizeof(float4);
static_assert(kOrdSmem <= (size_t)kHotMax * sizeof(float), "ordered-sum CTA: table + one ring per warp within the full table's 192 KB");

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One warp: the column-order fp32 sum of ONE row of an exact slice.  vals points at the row's first value group; group q
// of the row sits 32 float4 further for every q (lane = row layout of the slice).  A block is 128 groups = 512 columns:
// lane l holds groups 4 l .. 4 l + 3 of it, i.e. 16 consecutive columns.
__device__ __forceinline__ float ordered_row_sum(const float4 *vals, uint32_t ngl, float4 *ring, int lane) {
  const uint32_t n_blocks = (ngl + 127) / 128;
  auto request = [&](uint32_t b) {
    if (b < n_blocks) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t q = b * 128 + lane * 4 + i;
        if (q < ngl) cp_async16(ring + ((b % kOrdDepth) * 32 + lane) * 4 + i, vals + (size_t)q * 32);
      }
    }
    cp_async_commit();
  };
  for (uint32_t b = 0; b < (uint32_t)kOrdDepth - 1; b++) request(b);
  uint32_t ab = 0;                                                         // bits of the accumulator
  for (uint32_t b = 0; b < n_blocks; b++) {
    request(b + kOrdDepth - 1);
    cp_async_wait<kOrdDepth - 1>();
    const float4 *slot = ring + ((b % kOrdDepth) * 32 + lane) * 4;         // (a lane reads back what it copied itself)
    float x[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b * 128 + lane * 4 + i < ngl) v = slot[i];
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
    // non-negative and finite <=> bits < 0x7f800000 as unsigned (-0.0 takes the slow path too: harmless)
    uint32_t mbits = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) mbits = max(mbits, __float_as_uint(x[i]));
    int pos = 0;                                                           // columns of the block already in the accumulator
    const bool slow = b < 1 || !__all_sync(kFull, mbits < 0x7f800000u);
    if (!slow && ab >= 0x0C000000u && ab < 0x7f000000u) {
      // Common case, kept lean (the long rows are bound by the instructions issued per block): every addend is below
      // 2^14 ulps of the accumulator, so neither a clamp nor a saturating scan is needed, and the block carries into the
      // next binade iff M + (sum of all q) >= 2^24.
      const uint32_t e = ab >> 23, M = (ab & 0x7fffffu) | 0x800000u;
      const float scale = __uint_as_float((277u - e) << 23);
      if (__all_sync(kFull, __fmul_rn(__uint_as_float(mbits), scale) < 16384.f)) {
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) run += __float2uint_rn(__fmul_rn(x[i], scale));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) run += __shfl_xor_sync(kFull, run, o);
        if (M + run < 0x1000000u) { ab = (e << 23) | ((M + run) & 0x7fffffu); continue; }
      }
    }
    while (pos < 512) {
      if (slow || ab < 0x0C000000u || ab >= 0x7f000000u) {
        // one true add at a time (every lane computes the same accumulator; the values come straight from the ring)
        __syncwarp();
        float acc = __uint_as_float(ab);
        const float *flat = reinterpret_cast<const float *>(ring + (size_t)(b % kOrdDepth) * 32 * 4);
        for (int i = pos; i < 512; i++) acc = __fadd_rn(acc, (b * 128 + (uint32_t)(i >> 2) < ngl) ? flat[i] : 0.f);
        ab = __float_as_uint(acc);
        break;
      }
      const uint32_t e = ab >> 23, M = (ab & 0x7fffffu) | 0x800000u;
      const float scale = __uint_as_float((277u - e) << 23);               // 1 / ulp(acc) = 2^(23 - (e - 127))
      // q = rne(x / ulp), clamped at 2^24 (a carry for sure); columns before `pos` are already in the accumulator
      uint32_t run = 0;
#pragma unroll
      for (int i = 0; i < 16; i++)
        run += lane * 16 + i >= pos ? __float2uint_rn(fminf(__fmul_rn(x[i], scale), 16777216.f)) : 0u;
      uint32_t incl = min(run, 0x1000000u);                                // (saturating: 32 x 16 x 2^24 would overflow)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl = min(incl + t, 0x1000000u);
      }
      const unsigned carry = __ballot_sync(kFull, M + incl >= 0x1000000u);
      if (!carry) {                                                        // the whole block stays inside the binade
        const uint32_t Mn = M + __shfl_sync(kFull, incl, 31);
        ab = (e << 23) | (Mn & 0x7fffffu);
        break;
      }
      const int L = __ffs(carry) - 1;
      uint32_t before = __shfl_up_sync(kFull, incl, 1);                    // sum of the lanes before this one ...
      if (lane == 0) before = 0;
      int sub = 16;
      float xc = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i++) {                                       // ... then column by column (meaningful in lane L)
        const uint32_t q = lane * 16 + i >= pos ? __float2uint_rn(fminf(__fmul_rn(x[i], scale), 16777216.f)) : 0u;
        if (sub == 16) {
          if (M + before + q >= 0x1000000u) { sub = i; xc = x[i]; } else before += q;
        }
      }
      const int subL = __shfl_sync(kFull, sub, L);
      const uint32_t Mb = M + __shfl_sync(kFull, before, L);
      const float accb = __uint_as_float((e << 23) | (Mb & 0x7fffffu));
      ab = __float_as_uint(__fadd_rn(accb, __shfl_sync(kFull, xc, L)));     // the add that carries: rounded with the new ulp
      pos = L * 16 + subL + 1;
    }
  }
  cp_async_wait<0>();
  return __uint_as_float(ab);
}

}  // namespace gdn
