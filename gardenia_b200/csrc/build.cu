// build.cu -- edge list -> symmetric, sorted, duplicate-free, self-loop-free CSR ON THE GPU.
//
// Replaces the reference's builder for synthetic graphs -- Builder::MakeGraph / MakeCSR / SquishCSR, include/builder.h:
// 66-119,152-195,220-257: per-vertex degree counting with atomics, a scatter, then std::sort + unique + erase per row on
// the host, 40+ s at Kronecker scale 26 on 16 cores -- with one formulation that suits the machine: every undirected
// input edge {u, v} becomes the two 64-bit keys (u << 32 | v) and (v << 32 | u) (self loops become an all-ones sentinel),
// the keys are radix-sorted (least-significant digit first, 8 bits per pass, only the bits ids can have), and the CSR is
// read off the sorted array: adjacent duplicates dropped by a flag scan, column = low word, row offsets = lower bounds.
// The result is a pure function of the edge SET, so it is bit-identical to host/generator.cc build_symmetric_csr (and
// through it to the reference's builder, tests/test_host.py) whatever the order in which the atomics of either would land.
//
// All kernels are ours (no CUB / Thrust): a reduce-then-scan radix sort with warp-level multi-split ranking
// (__match_any_sync) and shared-memory staging so that every pass writes runs of consecutive keys; a three-kernel
// exclusive scan; binary-search row offsets.
#include "common.cuh"
#include <omp.h>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

namespace gdn {

namespace {

constexpr int kTileThreads = 256;
constexpr int kKeysPerThread = 16;
constexpr int kTile = kTileThreads * kKeysPerThread;          // keys per CTA and pass
constexpr unsigned long long kSentinel = ~0ull;

struct EdgePair { int32_t u, v; };

__global__ void __launch_bounds__(256)
max_id_kernel(const EdgePair *__restrict__ el, int64_t n, int *out) {
  int mx = 0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) mx = max(mx, max(el[e].u, el[e].v));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// both directions of every edge (the symmetrization of include/builder.h:76-87,220-236); self loops -> sentinel
__global__ void __launch_bounds__(256)
expand_keys(const EdgePair *__restrict__ el, int64_t n, unsigned long long *__restrict__ keys) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
    const EdgePair p = el[e];
    const unsigned long long a = ((unsigned long long)(uint32_t)p.u << 32) | (uint32_t)p.v;
    const unsigned long long b = ((unsigned long long)(uint32_t)p.v << 32) | (uint32_t)p.u;
    ulonglong2 kk;
    kk.x = p.u == p.v ? kSentinel : a;
    kk.y = p.u == p.v ? kSentinel : b;
    reinterpret_cast<ulonglong2 *>(keys)[e] = kk;
  }
}

// ---- radix pass, part 1: per-tile digit histograms, digit-major (hist[d * n_tiles + tile])
__global__ void __launch_bounds__(kTileThreads)
radix_hist(const unsigned long long *__restrict__ keys, uint64_t n, int shift, uint32_t n_tiles, uint32_t *__restrict__ hist) {
  __shared__ uint32_t s_h[256];
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)tile * kTile;
#pragma unroll 4
    for (int i = 0; i < kKeysPerThread; i++) {
      const uint64_t idx = base + (uint64_t)i * kTileThreads + threadIdx.x;
      if (idx < n) atomicAdd(&s_h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_tiles + tile] = s_h[threadIdx.x];
    __syncthreads();
  }
}

// ---- exclusive scan of uint32 counts into uint64 offsets: block sums, scan of the sums (one CTA), offsets
constexpr int kScanTile = 4096;
__global__ void __launch_bounds__(256)
scan_block_sums(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ sums) {
  __shared__ uint64_t s[8];
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t acc = 0;
  for (int i = threadIdx.x; i < kScanTile; i += 256) if (base + i < n) acc += in[base + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 8; w++) t += s[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024)
scan_sums_inplace(uint64_t *sums, uint64_t n) {               // one CTA, exclusive, in place
  __shared__ uint64_t s_w[32];
  __shared__ uint64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint64_t b = 0; b < n; b += 1024) {
    const uint64_t i = b + threadIdx.x;
    const uint64_t v = i < n ? sums[i] : 0;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    if (w == 0) {
      uint64_t x = s_w[lane], y = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(kFull, y, o); if (lane >= o) y += t; }
      s_w[lane] = y - x;
    }
    __syncthreads();
    const uint64_t carry = s_carry;
    if (i < n) sums[i] = carry + s_w[w] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[w] + incl;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256)
scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ sums, uint64_t *__restrict__ out) {
  __shared__ uint64_t s_w[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t carry = sums[blockIdx.x];
  for (int r = 0; r < kScanTile / 256; r++) {
    const uint64_t i = base + (uint64_t)r * 256 + threadIdx.x;
    const uint64_t v = i < n ? in[i] : 0;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    uint64_t wbase = 0, tot = 0;
    for (int q = 0; q < 8; q++) { if (q < w) wbase += s_w[q]; tot += s_w[q]; }
    if (i < n) out[i] = carry + wbase + incl - v;
    carry += tot;
    __syncthreads();
  }
}

// ---- radix pass, part 2: stable scatter.  Warp w of a CTA owns keys [w * 512, (w + 1) * 512) of the tile and takes them
// 32 at a time in order; __match_any_sync ranks a key among the keys of the same digit in its round, a per-warp counter
// array carries the rank across rounds.  The tile is then laid out digit by digit in shared memory and written in runs.
__global__ void __launch_bounds__(kTileThreads)
radix_scatter(const unsigned long long *__restrict__ keys, unsigned long long *__restrict__ out, uint64_t n, int shift,
              uint32_t n_tiles, const uint64_t *__restrict__ offs) {
  __shared__ uint32_t s_cnt[8][256];                          // per warp: keys of digit d seen so far (then: its base in the tile)
  __shared__ uint32_t s_start[256];                           // first position of digit d in the staged tile
  __shared__ unsigned long long s_keys[kTile];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < 8 * 256; i += kTileThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)tile * kTile + (uint64_t)w * (kTile / 8);
    unsigned long long k[kKeysPerThread];
    uint32_t rank[kKeysPerThread];
#pragma unroll
    for (int r = 0; r < kKeysPerThread; r++) {
      const uint64_t idx = base + (uint64_t)r * 32 + lane;
      const bool on = idx < n;
      k[r] = on ? keys[idx] : kSentinel;
      const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
      const unsigned same = __match_any_sync(kFull, on ? d : 256u + lane);      // (lanes past the end match nobody)
      const uint32_t before = s_cnt[w][d];
      rank[r] = before + __popc(same & lt);
      __syncwarp();
      if (on && (same & lt) == 0) s_cnt[w][d] = before + __popc(same);          // the first lane of each group
      __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan over the 8 warps, and the digit's start in the tile
    {
      const int d = threadIdx.x;
      uint32_t run = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { const uint32_t c = s_cnt[q][d]; s_cnt[q][d] = run; run += c; }
      // exclusive scan of `run` over the 256 digits
      uint32_t incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
      __shared__ uint32_t s_ws[8];
      if (lane == 31) s_ws[w] = incl;
      __syncthreads();
      uint32_t wb = 0;
      for (int q = 0; q < w; q++) wb += s_ws[q];
      s_start[d] = wb + incl - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kKeysPerThread; r++) {
      const uint64_t idx = base + (uint64_t)r * 32 + lane;
      if (idx < n) {
        const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
        s_keys[s_start[d] + s_cnt[w][d] + rank[r]] = k[r];
      }
    }
    __syncthreads();
    const uint64_t tile_n = min((uint64_t)kTile, n - (uint64_t)tile * kTile);
    for (uint32_t i = threadIdx.x; i < tile_n; i += kTileThreads) {
      const unsigned long long key = s_keys[i];
      const uint32_t d = (uint32_t)(key >> shift) & 255u;
      out[offs[(uint64_t)d * n_tiles + tile] + (i - s_start[d])] = key;
    }
    __syncthreads();
  }
}

// ---- read the CSR off the sorted keys
__device__ __forceinline__ bool keep_key(const unsigned long long *keys, uint64_t i) {
  const unsigned long long k = keys[i];
  return k != kSentinel && (i == 0 || keys[i - 1] != k);
}
__global__ void __launch_bounds__(256)
flag_counts(const unsigned long long *__restrict__ keys, uint64_t n, uint32_t *__restrict__ cnt) {   // kept keys per 256
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const unsigned b = __ballot_sync(kFull, i < n && keep_key(keys, i));
  __shared__ uint32_t s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; q++) t += s[q]; cnt[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256)
compact_keys(const unsigned long long *__restrict__ keys, uint64_t n, const uint64_t *__restrict__ offs, int32_t *__restrict__ col,
             unsigned long long *__restrict__ uniq) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const bool keep = i < n && keep_key(keys, i);
  const unsigned b = __ballot_sync(kFull, keep);
  __shared__ uint32_t s[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = __popc(b);
  __syncthreads();
  uint32_t wb = 0;
  for (int q = 0; q < w; q++) wb += s[q];
  if (keep) {
    const uint64_t pos = offs[blockIdx.x] + wb + __popc(b & ((1u << lane) - 1u));
    const unsigned long long k = keys[i];
    col[pos] = (int32_t)(uint32_t)k;
    uniq[pos] = k;                                            // (in place over the other key buffer: pos <= i, see below)
  }
}
// rowptr[v] = number of unique keys below (v << 32)
__global__ void __launch_bounds__(256)
row_offsets(const unsigned long long *__restrict__ uniq, uint64_t n_uniq, int64_t m, uint64_t *__restrict__ rowptr) {
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v <= m; v += (int64_t)gridDim.x * 256) {
    const unsigned long long key = (unsigned long long)v << 32;
    uint64_t lo = 0, hi = n_uniq;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (uniq[mid] < key) lo = mid + 1; else hi = mid; }
    rowptr[v] = lo;
  }
}
__global__ void __launch_bounds__(256)
max_degree_kernel(const uint64_t *__restrict__ rowptr, int64_t m, unsigned long long *out) {
  unsigned long long mx = 0;
  for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < m; v += (int64_t)gridDim.x * 256) mx = max(mx, (unsigned long long)(rowptr[v + 1] - rowptr[v]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

struct Buf {
  void *p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <typename T> T *as() const { return (T *)p; }
  int alloc(size_t bytes) { GDN_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16))); return GDN_OK; }
};

// exclusive scan of n uint32 counts -> n uint64 offsets; *total = sum
int exclusive_scan(const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *total, cudaStream_t st) {
  const uint64_t nb = (n + kScanTile - 1) / kScanTile;
  Buf sums;
  GDN_CHECK(sums.alloc(sizeof(uint64_t) * (nb + 1)));
  scan_block_sums<<<(unsigned)nb, 256, 0, st>>>(in, n, sums.as<uint64_t>());
  // total = last block's exclusive offset + its sum: keep the sums' total by scanning nb + 1 entries (the extra one is zero)
  GDN_CUDA(cudaMemsetAsync(sums.as<uint64_t>() + nb, 0, sizeof(uint64_t), st));
  scan_sums_inplace<<<1, 1024, 0, st>>>(sums.as<uint64_t>(), nb + 1);
  scan_apply<<<(unsigned)nb, 256, 0, st>>>(in, n, sums.as<uint64_t>(), out);
  if (total) GDN_CUDA(cudaMemcpyAsync(total, sums.as<uint64_t>() + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  return GDN_OK;
}

}  // namespace

// el: n_edges (u, v) pairs on the HOST.  Outputs are new[]-allocated host arrays (the host Graph adopts them).
int gpu_build_symmetric_csr(const void *el_host, int64_t n_edges, int64_t *m_out, uint64_t **rowptr_out, int32_t **col_out,
                            uint64_t *nnz_out, int32_t *maxdeg_out, double ms[4]) {
  GDN_CHECK(ensure_init());
  if (n_edges < 0 || (uint64_t)n_edges > (1ull << 33)) { set_error("gdn_build_csr: bad edge count"); return GDN_ERR_ARG; }
  cudaStream_t st = lib().stream;
  const int sm = lib().sm_count;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  auto t0 = now();
  const uint64_t n = 2 * (uint64_t)n_edges;                   // keys
  Buf el, ka, kb, d_scalar;
  GDN_CHECK(el.alloc(sizeof(EdgePair) * (size_t)n_edges));
  GDN_CHECK(ka.alloc(sizeof(unsigned long long) * n));
  GDN_CHECK(kb.alloc(sizeof(unsigned long long) * n));
  GDN_CHECK(d_scalar.alloc(16));
  GDN_CUDA(cudaMemcpyAsync(el.p, el_host, sizeof(EdgePair) * (size_t)n_edges, cudaMemcpyHostToDevice, st));
  GDN_CUDA(cudaMemsetAsync(d_scalar.p, 0, 16, st));
  max_id_kernel<<<sm * 8, 256, 0, st>>>(el.as<EdgePair>(), n_edges, d_scalar.as<int>());
  expand_keys<<<sm * 8, 256, 0, st>>>(el.as<EdgePair>(), n_edges, ka.as<unsigned long long>());
  int max_id = 0;
  GDN_CUDA(cudaMemcpyAsync(&max_id, d_scalar.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  cudaFree(el.p); el.p = nullptr;
  const int64_t m = n_edges ? (int64_t)max_id + 1 : 0;        // FindMaxVertexID + 1, include/builder.h:244-245
  if (ms) ms[0] = since(t0);
  t0 = now();
  // ---- sort: the bits an id can have, low word then high word
  int bits = 1;
  while ((1ll << bits) < std::max<int64_t>(m, 2)) bits++;
  const uint64_t n_tiles64 = (n + kTile - 1) / kTile;
  if (n_tiles64 > 0x7fffffffull / 256) { set_error("gdn_build_csr: too many keys"); return GDN_ERR_ARG; }
  const uint32_t n_tiles = (uint32_t)n_tiles64;
  Buf hist, offs;
  GDN_CHECK(hist.alloc(sizeof(uint32_t) * 256 * (size_t)std::max<uint32_t>(n_tiles, 1)));
  GDN_CHECK(offs.alloc(sizeof(uint64_t) * 256 * (size_t)std::max<uint32_t>(n_tiles, 1)));
  unsigned long long *src = ka.as<unsigned long long>(), *dst = kb.as<unsigned long long>();
  const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles64, 1), (uint64_t)sm * 16);
  for (int word = 0; word < 2 && n > 0; word++) {
    for (int sh = 0; sh < bits; sh += 8) {
      const int shift = word * 32 + sh;
      radix_hist<<<grid, kTileThreads, 0, st>>>(src, n, shift, n_tiles, hist.as<uint32_t>());
      GDN_CHECK(exclusive_scan(hist.as<uint32_t>(), 256ull * n_tiles, offs.as<uint64_t>(), nullptr, st));
      radix_scatter<<<grid, kTileThreads, 0, st>>>(src, dst, n, shift, n_tiles, offs.as<uint64_t>());
      std::swap(src, dst);
    }
  }
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  if (ms) ms[1] = since(t0);
  t0 = now();
  // ---- unique + self-loop removal (SquishCSR, include/builder.h:152-183) and the CSR arrays
  const uint64_t nflag = (n + 255) / 256;
  Buf cnt, coff, col, rowptr;
  GDN_CHECK(cnt.alloc(sizeof(uint32_t) * std::max<uint64_t>(nflag, 1)));
  GDN_CHECK(coff.alloc(sizeof(uint64_t) * std::max<uint64_t>(nflag, 1)));
  uint64_t nnz = 0;
  if (n > 0) {
    flag_counts<<<(unsigned)nflag, 256, 0, st>>>(src, n, cnt.as<uint32_t>());
    GDN_CHECK(exclusive_scan(cnt.as<uint32_t>(), nflag, coff.as<uint64_t>(), &nnz, st));
  }
  GDN_CHECK(col.alloc(sizeof(int32_t) * std::max<uint64_t>(nnz, 1)));
  GDN_CHECK(rowptr.alloc(sizeof(uint64_t) * (size_t)(m + 1)));
  // the unique keys go into the OTHER buffer (the sorted one is still being read)
  if (n > 0) compact_keys<<<(unsigned)nflag, 256, 0, st>>>(src, n, coff.as<uint64_t>(), col.as<int32_t>(), dst);
  row_offsets<<<sm * 8, 256, 0, st>>>(dst, nnz, m, rowptr.as<uint64_t>());
  GDN_CUDA(cudaMemsetAsync(d_scalar.p, 0, 16, st));
  max_degree_kernel<<<sm * 8, 256, 0, st>>>(rowptr.as<uint64_t>(), m, d_scalar.as<unsigned long long>());
  unsigned long long maxdeg = 0;
  GDN_CUDA(cudaMemcpyAsync(&maxdeg, d_scalar.p, sizeof(maxdeg), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  GDN_CUDA(cudaGetLastError());
  if (ms) ms[2] = since(t0);
  t0 = now();
  uint64_t *h_rowptr = new uint64_t[m + 1];
  int32_t *h_col = new int32_t[std::max<uint64_t>(nnz, 1)];
  // first touch by all host threads: a pageable destination whose pages do not exist yet is copied at 2 GB/s
  {
    const int64_t pg = 4096 / sizeof(int32_t);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nnz; i += pg) h_col[i] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i <= m; i += 512) h_rowptr[i] = 0;
  }
  GDN_CUDA(cudaMemcpyAsync(h_rowptr, rowptr.p, sizeof(uint64_t) * (size_t)(m + 1), cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaMemcpyAsync(h_col, col.p, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  if (ms) ms[3] = since(t0);
  *m_out = m; *rowptr_out = h_rowptr; *col_out = h_col; *nnz_out = nnz; *maxdeg_out = (int32_t)maxdeg;
  return GDN_OK;
}

}  // namespace gdn
