// comm.cu -- 1-D row partition and the per-iteration exchange over NVLink.
//
// The reference has no multi-GPU code on this path; this is the
// added parallelism strategy of SURVEY §8(e): contiguous equal-width vertex
// ranges, one process per GPU, ONE NCCL allgather per iteration of the PR/SpMV
// vector slice (or the BFS frontier-bitmap slice) plus an allreduce of the
// scalar that drives the host-side controller.
//
// NCCL is resolved with dlopen at gdn_comm_init time so that libgdn_b200.so
// loads (and the host-only entry points work) on machines without NCCL/GPUs.
// When the process already has torch's bundled libnccl.so.2 mapped, dlopen
// returns that copy.
#include "pull.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <omp.h>

namespace gdn {

// Minimal NCCL ABI (stable since 2.x): opaque comm, 128-byte unique id.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
};
static Nccl &nccl() { static Nccl n; return n; }

static int load_nccl() {
  Nccl &n = nccl();
  if (n.h) return GDN_OK;
  const char *names[] = {getenv("GDN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm) continue;
    n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.h) break;
  }
  if (!n.h) { set_error("cannot dlopen libnccl.so.2 (set GDN_NCCL_LIB): %s", dlerror()); return GDN_ERR_NCCL; }
#define SYM(field, name)                                                     \
  *(void **)(&n.field) = dlsym(n.h, name);                                   \
  if (!n.field) { set_error("NCCL symbol %s missing", name); return GDN_ERR_NCCL; }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllGather, "ncclAllGather")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
  return GDN_OK;
}

#define GDN_NCCL(call)                                                                   \
  do {                                                                                   \
    int r_ = (call);                                                                     \
    if (r_ != ncclSuccess) {                                                             \
      set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r_)); \
      return GDN_ERR_NCCL;                                                               \
    }                                                                                    \
  } while (0)

// ------------------------------------------------------------------ the gang: one process, one worker thread per GPU
// gdn_init_gpus(n) ("the replacement may spawn one host thread per GPU internally").  A worker owns a device
// context (Lib), takes the rank of its GPU and runs the same partitioned code as a one-process-per-GPU rank; what the
// ranks of separate processes exchange through NCCL / CUDA IPC (pointers of peer-mapped buffers, the renumbering, barriers)
// the workers exchange through shared host memory.  The PageRank iteration itself has no collective in either mode.
struct Gang {
  int n = 0;
  std::vector<std::thread> th;
  std::vector<Lib *> libs;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  int (*fn)(int, void *) = nullptr;
  void *arg = nullptr;
  unsigned long long seq = 0;          // jobs posted
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
  std::string err;                     // message of the first worker that failed
  // barrier of the workers
  std::mutex bmu;
  std::condition_variable bcv;
  int bcount = 0;
  unsigned long long bgen = 0;
  // pointer exchange / shared host scratch
  void *xchg[8] = {};
  void *shared = nullptr;
  size_t shared_bytes = 0;
};
static Gang &gang() { static Gang g; return g; }
static thread_local int tl_gang_rank = -1;

int gang_size() { return gang().n; }
bool gang_worker() { return tl_gang_rank >= 0; }

void gang_barrier() {
  Gang &g = gang();
  std::unique_lock<std::mutex> lk(g.bmu);
  const unsigned long long gen = g.bgen;
  if (++g.bcount == g.n) { g.bcount = 0; g.bgen++; g.bcv.notify_all(); }
  else g.bcv.wait(lk, [&] { return g.bgen != gen; });
}

void *gang_shared(size_t bytes) {
  Gang &g = gang();
  gang_barrier();                                     // everybody is done with the previous contents
  if (tl_gang_rank == 0 && g.shared_bytes < bytes) {
    if (g.shared) cudaFreeHost(g.shared);
    g.shared = nullptr; g.shared_bytes = 0;
    if (cudaHostAlloc(&g.shared, bytes + bytes / 8, cudaHostAllocPortable) == cudaSuccess) g.shared_bytes = bytes + bytes / 8;
    else cudaGetLastError();
  }
  gang_barrier();
  return g.shared_bytes >= bytes ? g.shared : nullptr;
}

static void gang_worker_main(int rank) {
  Gang &g = gang();
  tl_gang_rank = rank;
  lib_bind(g.libs[rank]);
  omp_set_num_threads(std::max(1, omp_get_num_procs() / g.n));      // the host side of a rank (layout preprocessing) is OpenMP code
  int rc = gdn_init(rank);
  if (rc == GDN_OK) {
    for (int p = 0; p < g.n && rc == GDN_OK; p++) {
      if (p == rank) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("GPU %d cannot map GPU %d (%s)", rank, p, cudaGetErrorString(e)); rc = GDN_ERR_CUDA; }
      cudaGetLastError();
    }
  }
  unsigned long long seen = 0;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    g.rc[rank] = rc;
    if (rc != GDN_OK && g.err.empty()) g.err = gdn_last_error();
    if (--g.pending == 0) g.cv_done.notify_all();
  }
  for (;;) {
    int (*fn)(int, void *);
    void *arg;
    {
      std::unique_lock<std::mutex> lk(g.mu);
      g.cv_job.wait(lk, [&] { return g.quit || g.seq != seen; });
      if (g.quit) break;
      seen = g.seq; fn = g.fn; arg = g.arg;
    }
    const int r = fn(rank, arg);
    std::lock_guard<std::mutex> lk(g.mu);
    g.rc[rank] = r;
    if (r != GDN_OK && g.err.empty()) g.err = gdn_last_error();
    if (--g.pending == 0) g.cv_done.notify_all();
  }
  gdn_finalize();
  lib_bind(nullptr);
}

int gang_run(int (*fn)(int, void *), void *arg) {
  Gang &g = gang();
  if (g.n == 0) { set_error("gdn_init_gpus has not been called"); return GDN_ERR_ARG; }
  std::unique_lock<std::mutex> lk(g.mu);
  g.fn = fn; g.arg = arg; g.pending = g.n; g.err.clear();
  g.seq++;
  g.cv_job.notify_all();
  g.cv_done.wait(lk, [&] { return g.pending == 0; });
  for (int r = 0; r < g.n; r++)
    if (g.rc[r] != GDN_OK) { set_error("GPU %d: %s", r, g.err.c_str()); return g.rc[r]; }
  return GDN_OK;
}

static int gang_start(int n) {
  Gang &g = gang();
  if (g.n == n) return GDN_OK;
  if (g.n) { set_error("gdn_init_gpus: already running on %d GPUs (gdn_finalize first)", g.n); return GDN_ERR_ARG; }
  g.n = n; g.quit = false; g.seq = 0; g.pending = n; g.rc.assign(n, 0); g.err.clear();
  g.libs.clear();
  for (int r = 0; r < n; r++) g.libs.push_back(new Lib());
  for (int r = 0; r < n; r++) g.th.emplace_back(gang_worker_main, r);
  std::unique_lock<std::mutex> lk(g.mu);
  g.cv_done.wait(lk, [&] { return g.pending == 0; });
  for (int r = 0; r < n; r++)
    if (g.rc[r] != GDN_OK) { const int rc = g.rc[r]; const std::string msg = g.err; lk.unlock(); set_error("GPU %d: %s", r, msg.c_str()); return rc; }
  return GDN_OK;
}

void gang_stop() {
  Gang &g = gang();
  if (!g.n) return;
  { std::lock_guard<std::mutex> lk(g.mu); g.quit = true; }
  g.cv_job.notify_all();
  for (auto &t : g.th) t.join();
  g.th.clear();
  for (Lib *l : g.libs) delete l;
  g.libs.clear();
  if (g.shared) cudaFreeHost(g.shared);
  g.shared = nullptr; g.shared_bytes = 0;
  g.n = 0;
}

int comm_size() { return gang_worker() ? gang().n : nccl().size; }
int comm_rank() { return gang_worker() ? tl_gang_rank : nccl().rank; }

// Equal-width ranges: width = ceil(m / nparts) rounded up to 1024 vertices, so that
// every slice is a whole number of 32-word bitmap groups (and of 64-bit host words).
int64_t partition_width(int64_t m, int nparts) {
  int64_t w = (m + nparts - 1) / nparts;
  return (w + 1023) / 1024 * 1024;
}

// In-place allgather of this rank's slice of a full-length fp32 vector, plus an
// allreduce(sum) of one double (the PR L1 delta).  No-op at 1 GPU.
int pr_exchange(gdn_graph *g, float *vec, double *err_slot) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  const int64_t w = partition_width(g->m, n.size);
  GDN_NCCL(n.AllGather(vec + (int64_t)n.rank * w, vec, (size_t)w, ncclFloat32, n.comm, lib().stream));
  if (err_slot) GDN_NCCL(n.AllReduce(err_slot, err_slot, 1, ncclFloat64, ncclSum, n.comm, lib().stream));
  return GDN_OK;
}

// Degree-sorted id space (pull.cu): each rank owns a hot slice [R*Hp, +Hp) and a cold
// slice [H + R*Wc, +Wc) of contrib -> two in-place allgathers + allreduce of the delta.
int pull_exchange(gdn_graph *g, float *vec, double *err_slot) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  const PullLayout &L = g->pull;
  if (L.Hp > 0) GDN_NCCL(n.AllGather(vec + (int64_t)n.rank * L.Hp, vec, (size_t)L.Hp, ncclFloat32, n.comm, lib().stream));
  if (L.Wc > 0) GDN_NCCL(n.AllGather(vec + L.H + (int64_t)n.rank * L.Wc, vec + L.H, (size_t)L.Wc, ncclFloat32, n.comm, lib().stream));
  if (err_slot) GDN_NCCL(n.AllReduce(err_slot, err_slot, 1, ncclFloat64, ncclSum, n.comm, lib().stream));
  return GDN_OK;
}

// In-place allgather of this rank's slice of a packed bitmap (32-bit words) and
// allreduce(sum) of n 64-bit counters.
int bitmap_exchange(gdn_graph *g, uint32_t *bm, long long *counters, int n_counters) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  const int64_t ww = partition_width(g->m, n.size) / 32;
  if (bm) GDN_NCCL(n.AllGather(bm + (int64_t)n.rank * ww, bm, (size_t)ww, ncclUint32, n.comm, lib().stream));
  if (counters) GDN_NCCL(n.AllReduce(counters, counters, (size_t)n_counters, ncclInt64, ncclSum, n.comm, lib().stream));
  return GDN_OK;
}

// own[i] |= OR over the other ranks' copies of this rank's slice
__global__ void or_merge(uint32_t *__restrict__ own, const uint32_t *__restrict__ xbuf, int64_t ww, int P, int R) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ww; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = own[i];
    for (int p = 0; p < P; p++)
      if (p != R) v |= xbuf[(int64_t)p * ww + i];
    own[i] = v;
  }
}

// Top-down exchange: every rank holds a full-length "discovered" bitmap.  Slice p
// is sent to rank p (grouped send/recv over NVLink), the owner ORs the P copies,
// and one allgather republishes the merged slices.  m/8 bytes out, m/8 bytes in.
int bfs_merge_or(gdn_graph *g, uint32_t *bm, uint32_t *xbuf) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  const int64_t ww = partition_width(g->m, n.size) / 32;
  GDN_NCCL(n.GroupStart());
  for (int p = 0; p < n.size; p++) {
    if (p == n.rank) continue;
    GDN_NCCL(n.Send(bm + (int64_t)p * ww, (size_t)ww, ncclUint32, p, n.comm, lib().stream));
    GDN_NCCL(n.Recv(xbuf + (int64_t)p * ww, (size_t)ww, ncclUint32, p, n.comm, lib().stream));
  }
  GDN_NCCL(n.GroupEnd());
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ww + 255) / 256, (int64_t)lib().sm_count * 4));
  or_merge<<<grid, 256, 0, lib().stream>>>(bm + (int64_t)n.rank * ww, xbuf, ww, n.size, n.rank);
  GDN_NCCL(n.AllGather(bm + (int64_t)n.rank * ww, bm, (size_t)ww, ncclUint32, n.comm, lib().stream));
  return GDN_OK;
}

int bfs_allgather_words(gdn_graph *g, uint32_t *bm) { return bitmap_exchange(g, bm, nullptr, 0); }

int allgather_i32(int32_t *buf, int64_t per_rank) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  GDN_NCCL(n.AllGather(buf + (int64_t)n.rank * per_rank, buf, (size_t)per_rank, ncclInt32, n.comm, lib().stream));
  return GDN_OK;
}

int allreduce_i64(long long *d_p, int cnt) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  GDN_NCCL(n.AllReduce(d_p, d_p, (size_t)cnt, ncclInt64, ncclSum, n.comm, lib().stream));
  return GDN_OK;
}


// ------------------------------------------------------------------ peer-mapped exchange (PageRank)
// The contrib vector of a row partition is exchanged WITHOUT a collective: every GPU maps the vectors of the others
// (CUDA IPC between the one-process-per-GPU ranks) and the row epilogues store each new value into all of them over
// NVLink while the iteration is still running (pull.cuh contrib_store).  What is left between two iterations is a barrier
// over the box and the sum of P deltas: one single-block kernel per GPU on device-side flags in a peer-mapped mailbox.
//   mailbox (per GPU):  flags[8]   flags[q] = last epoch GPU q has arrived at (written by q)
//                       vals[2][8] vals[e & 1][q] = the value GPU q published in epoch e
// Epochs only grow (one per pull_peer_sync call, the same sequence on every rank), so nothing is ever reset.
struct PeerBox {
  bool ready = false, failed = false;
  unsigned long long *mail = nullptr;          // this GPU's mailbox (device memory, exported)
  unsigned long long *peer_mail[8] = {};       // every GPU's mailbox as mapped here ([rank] = mail)
  unsigned long long epoch = 0;
  int *timeout_flag = nullptr;                 // device: a barrier gave up waiting (a peer died)
};
static PeerBox &box() { static thread_local PeerBox b; return b; }      // (a gang worker has its own)
constexpr size_t kMailBytes = 2u << 20;        // a whole 2 MB block: small cudaMalloc blocks are sub-allocated and not exportable alone

struct PeerMailArgs {
  unsigned long long *mail[8];
};

// Exchange one device pointer per rank: mine[rank] = local; the others are the peers' allocations mapped into this
// process.  The 64-byte IPC handles travel through one ncclAllGather.
static int ipc_exchange(void *local, void **mapped) {
  if (gang_worker()) {
    // one process: peer access is enabled (gang_worker_main), the pointers themselves are valid on every GPU
    Gang &g = gang();
    g.xchg[tl_gang_rank] = local;
    gang_barrier();
    for (int p = 0; p < g.n; p++) mapped[p] = g.xchg[p];
    gang_barrier();
    return GDN_OK;
  }
  Nccl &n = nccl();
  cudaStream_t st = lib().stream;
  cudaIpcMemHandle_t mine;
  GDN_CUDA(cudaIpcGetMemHandle(&mine, local));
  unsigned char *d_h = nullptr;
  GDN_CUDA(cudaMalloc((void **)&d_h, sizeof(cudaIpcMemHandle_t) * n.size));
  GDN_CUDA(cudaMemcpyAsync(d_h + sizeof(cudaIpcMemHandle_t) * n.rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  GDN_NCCL(n.AllGather(d_h + sizeof(cudaIpcMemHandle_t) * n.rank, d_h, sizeof(cudaIpcMemHandle_t), ncclUint8, n.comm, st));
  std::vector<cudaIpcMemHandle_t> all((size_t)n.size);
  GDN_CUDA(cudaMemcpyAsync(all.data(), d_h, sizeof(cudaIpcMemHandle_t) * n.size, cudaMemcpyDeviceToHost, st));
  GDN_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_h);
  for (int p = 0; p < n.size; p++) {
    if (p == n.rank) { mapped[p] = local; continue; }
    GDN_CUDA(cudaIpcOpenMemHandle(&mapped[p], all[p], cudaIpcMemLazyEnablePeerAccess));
  }
  return GDN_OK;
}

static int comm_barrier_host() {          // every rank has reached this point (and its stream has drained)
  if (gang_worker()) {
    GDN_CUDA(cudaStreamSynchronize(lib().stream));
    gang_barrier();
    return GDN_OK;
  }
  Nccl &n = nccl();
  if (n.size == 1 || !n.comm) return GDN_OK;
  int *d = nullptr;
  GDN_CUDA(cudaMalloc((void **)&d, sizeof(int)));
  GDN_CUDA(cudaMemsetAsync(d, 0, sizeof(int), lib().stream));
  GDN_NCCL(n.AllReduce(d, d, 1, ncclInt32, ncclSum, n.comm, lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  cudaFree(d);
  return GDN_OK;
}

static int peer_box_setup() {
  PeerBox &b = box();
  const int P = comm_size();
  if (b.ready || b.failed) return GDN_OK;
  if (P > 8) { b.failed = true; return GDN_OK; }
  b.failed = true;                           // until everything below has worked
  GDN_CUDA(cudaMalloc((void **)&b.mail, kMailBytes));
  GDN_CUDA(cudaMemsetAsync(b.mail, 0, kMailBytes, lib().stream));
  GDN_CUDA(cudaMalloc((void **)&b.timeout_flag, sizeof(int)));
  GDN_CUDA(cudaMemsetAsync(b.timeout_flag, 0, sizeof(int), lib().stream));
  GDN_CUDA(cudaStreamSynchronize(lib().stream));
  void *mapped[8] = {};
  GDN_CHECK(ipc_exchange(b.mail, mapped));
  for (int p = 0; p < P; p++) b.peer_mail[p] = (unsigned long long *)mapped[p];
  GDN_CHECK(comm_barrier_host());            // nobody signals into a mailbox that is not zeroed yet
  b.epoch = 0;
  b.failed = false;
  b.ready = true;
  return GDN_OK;
}

static void peer_box_teardown() {
  PeerBox &b = box();
  const int P = comm_size(), R = comm_rank();
  if (b.mail) {
    for (int p = 0; p < P && !gang_worker(); p++)
      if (p != R && b.peer_mail[p]) cudaIpcCloseMemHandle(b.peer_mail[p]);
    comm_barrier_host();                     // the peers have unmapped this mailbox before it is freed
    cudaFree(b.mail);
    cudaFree(b.timeout_flag);
  }
  b = PeerBox();
}

// Map the two contrib vectors of every other GPU (once per graph; collective: all ranks solve the same graph together).
int pull_peer_setup(gdn_graph *g) {
  const int P = comm_size();
  if (P == 1 || g->peer_ready || g->peer_failed) return GDN_OK;
  GDN_CHECK(peer_box_setup());
  if (!box().ready) { g->peer_failed = true; return GDN_OK; }
  for (int k = 0; k < 2; k++) {
    void *mapped[8] = {};
    const int rc = ipc_exchange(g->contrib[k], mapped);
    if (rc != GDN_OK) { g->peer_failed = true; return rc; }
    for (int p = 0; p < P; p++) g->peer_contrib[k][p] = (float *)mapped[p];
  }
  g->peer_ready = true;
  return GDN_OK;
}
bool pull_peer_ready(const gdn_graph *g) { return g->peer_ready; }

// Unmap the peers' vectors; the owner frees its own only after every rank has done so.
int pull_peer_release(gdn_graph *g) {
  const int P = comm_size(), R = comm_rank();
  if (!g->peer_ready) return GDN_OK;
  for (int k = 0; k < 2; k++)
    for (int p = 0; p < P; p++)
      if (p != R && g->peer_contrib[k][p]) { if (!gang_worker()) cudaIpcCloseMemHandle(g->peer_contrib[k][p]); g->peer_contrib[k][p] = nullptr; }
  g->peer_ready = false;
  return comm_barrier_host();
}

void pull_peer_args(const gdn_graph *g, int buf_out, SellArgs &a) {
  const int P = comm_size(), R = comm_rank();
  a.n_peers = 0;
  for (int p = 0; p < P; p++) {
    if (p == R) continue;
    a.peer_out[a.n_peers] = g->peer_contrib[buf_out][p];
    a.peer_other[a.n_peers] = g->peer_contrib[buf_out ^ 1][p];
    a.n_peers++;
  }
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One block per GPU.  (1) fixed-order sum of this GPU's partials, (2) publish it into slot [epoch & 1][rank] of every
// mailbox, (3) release-store the epoch into flag [rank] of every mailbox -- the stores of the kernels that ran before
// this one on the stream (the row epilogues' remote contrib values) are ordered before it -- (4) wait until all P flags
// of the own mailbox have reached the epoch, (5) add the P published values in rank order: every GPU gets the same bits.
__global__ void __launch_bounds__(256)
peer_sync_kernel(const double *__restrict__ partial, int n_partial, PeerMailArgs pm, int rank, int P,
                 unsigned long long epoch, double *out, int iter, double eps, int32_t *done, int *timeout_flag) {
  __shared__ double s[256];
  if (done && *done) return;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  const int t = threadIdx.x;
  if (t < P) {
    double *vals = reinterpret_cast<double *>(pm.mail[t] + 8);
    vals[(epoch & 1) * 8 + rank] = s[0];
    __threadfence_system();
    st_release_sys(pm.mail[t] + rank, epoch);
  }
  if (t < P) {
    const unsigned long long *flag = pm.mail[rank] + t;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < epoch) {
      if (globaltimer_ns() - t0 > 20000000000ull) { atomicExch(timeout_flag, 1); break; }   // 20 s: a peer is gone
    }
  }
  __syncthreads();
  if (t == 0 && out) {
    const double *vals = reinterpret_cast<const double *>(pm.mail[rank] + 8) + (epoch & 1) * 8;
    double tot = 0.0;
    for (int q = 0; q < P; q++) tot += vals[q];
    out[iter] = tot;
    if (done && tot < eps) *done = iter + 1;
  }
}

int pull_peer_sync(gdn_graph *g, const double *partial, int n_partial, double *err_out, int iter, double eps, int32_t *done,
                   cudaStream_t s) {
  PeerBox &b = box();
  const int P = comm_size(), R = comm_rank();
  (void)g;
  PeerMailArgs pm = {};
  for (int p = 0; p < P; p++) pm.mail[p] = b.peer_mail[p];
  b.epoch++;
  peer_sync_kernel<<<1, 256, 0, s>>>(partial, n_partial, pm, R, P, b.epoch, err_out, iter, eps, done, b.timeout_flag);
  return GDN_OK;
}

// after a solve: did any barrier give up?
int pull_peer_check() {
  PeerBox &b = box();
  if (!b.ready) return GDN_OK;
  int h = 0;
  GDN_CUDA(cudaMemcpy(&h, b.timeout_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) { set_error("peer barrier timed out (a rank of the row partition stopped)"); return GDN_ERR_NCCL; }
  return GDN_OK;
}

}  // namespace gdn

using namespace gdn;

extern "C" {

int gdn_partition_rows(int64_t m, int nparts, int64_t *bounds) {
  if (m <= 0 || nparts <= 0 || !bounds) { set_error("gdn_partition_rows: bad argument"); return GDN_ERR_ARG; }
  const int64_t w = partition_width(m, nparts);
  for (int p = 0; p <= nparts; p++) bounds[p] = std::min<int64_t>((int64_t)p * w, m);
  return GDN_OK;
}

int gdn_comm_unique_id(uint8_t id[128]) {
  GDN_CHECK(load_nccl());
  ncclUniqueId u;
  GDN_NCCL(nccl().GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return GDN_OK;
}

int gdn_comm_init(int rank, int nranks, const uint8_t id[128]) {
  GDN_CHECK(ensure_init());
  GDN_CHECK(load_nccl());
  Nccl &n = nccl();
  if (n.comm) gdn_comm_destroy();
  if (nranks < 1 || rank < 0 || rank >= nranks) { set_error("gdn_comm_init: bad rank/size"); return GDN_ERR_ARG; }
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  GDN_NCCL(n.CommInitRank(&n.comm, nranks, u, rank));
  n.rank = rank;
  n.size = nranks;
  return GDN_OK;
}

int gdn_comm_destroy(void) {
  Nccl &n = nccl();
  if (n.comm) { peer_box_teardown(); n.CommDestroy(n.comm); n.comm = nullptr; }
  n.rank = 0;
  n.size = 1;
  return GDN_OK;
}

int gdn_init_gpus(int ngpus) {
  const int have = gdn_device_count();
  if (have == 0) { set_error("no CUDA device: libgdn_b200 has no CPU fallback"); return GDN_ERR_NO_DEVICE; }
  if (ngpus < 1 || ngpus > have || ngpus > 8) { set_error("gdn_init_gpus: %d GPUs asked for, %d present (at most 8)", ngpus, have); return GDN_ERR_ARG; }
  if (ngpus == 1) { gang_stop(); return gdn_init(0); }
  GDN_CHECK(gdn_init(0));            // the calling thread keeps a context of its own (resident API, device buffers)
  return gang_start(ngpus);
}
int gdn_gpus(void) { return gang_size() ? gang_size() : 1; }

int gdn_comm_rank(void) { return nccl().rank; }
int gdn_comm_size(void) { return nccl().size; }

}  // extern "C"
