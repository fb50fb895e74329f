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
#include "common.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

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

int comm_size() { return nccl().size; }
int comm_rank() { return nccl().rank; }

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

int allreduce_i64(long long *d_p, int cnt) {
  Nccl &n = nccl();
  if (n.size == 1) return GDN_OK;
  GDN_NCCL(n.AllReduce(d_p, d_p, (size_t)cnt, ncclInt64, ncclSum, n.comm, lib().stream));
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
  if (n.comm) { n.CommDestroy(n.comm); n.comm = nullptr; }
  n.rank = 0;
  n.size = 1;
  return GDN_OK;
}

int gdn_comm_rank(void) { return nccl().rank; }
int gdn_comm_size(void) { return nccl().size; }

}  // extern "C"
