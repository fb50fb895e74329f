/*
 * gdn_b200.h -- C-ABI of libgdn_b200.so: the B200-native (sm_100a) CSR
 * traversal engine that drops in behind Gardenia's solver entry points.
 *
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 * Every function returns GDN_OK (0) or a negative gdn_status and never throws
 * or exits; gdn_last_error() holds a message for the calling thread.  There is
 * NO CPU fallback: without a CUDA device every compute entry point returns
 * GDN_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   gdn_bfs / gdn_bfs_i32                 void BFSSolver(Graph&, int source, DistT*)      src/bfs/bfs.h:43
 *                                         gen-1 BFSSolver(int m,int nnz,int source,int* in_row_offsets,...)
 *                                                                                         src/bfs/hybrid_base.cu:63
 *   gdn_pagerank_pull / _i32              void PRSolver(Graph&, ScoreT*)                  src/pr/pr.h:31
 *                                         gen-1 PRSolver(int m,int nnz,IndexT* in_row_offsets,...) src/pr/vector.cu:83
 *   gdn_spmv_csr / _i32                   void SpmvSolver(Graph&, const ValueT* Ax, const ValueT* x, ValueT* y)
 *                                                                                         src/spmv/spmv.h:29
 *                                         gen-1 SpmvSolver(int m,int nnz,IndexT* ApT,...) src/spmv/cusparse.cu:12
 *   gdn_graph_create/_destroy             the cudaMalloc+cudaMemcpy prologue / cudaFree epilogue every gen-2
 *                                         CUDA solver repeats, e.g. src/pr/warp.cu:137-155,197-204
 *   gdn_*_resident                        the timed region of those solvers, e.g. src/pr/base.cu:109-128
 *   gdn_read_graph*, gdn_generate         Graph(prefix, filetype, symmetrize, need_reverse) include/csr_graph.h:211-250,
 *                                         read_graph() include/graph_io.h:357-377,
 *                                         Builder::MakeGraph include/builder.h:258-274
 * INTEGRATION.md shows the C++ shims a Gardenia maintainer links instead of a
 * variant object file.
 */
#ifndef GDN_B200_H_
#define GDN_B200_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDN_VERSION 100          /* 0.1.0 */
#define GDN_INFINITY 1000000000  /* MYINFINITY, include/common.h:66 */
#define GDN_MAX_PR_ITER 128      /* >= MAX_ITER + 1, src/pr/pr.h:12 */
#define GDN_MAX_BFS_STEPS 256    /* recorded steps; longer runs keep counting in n_steps */

typedef enum {
  GDN_OK = 0,
  GDN_ERR_NO_DEVICE = -1,   /* no CUDA device / wrong architecture */
  GDN_ERR_CUDA = -2,        /* CUDA runtime error (message in gdn_last_error) */
  GDN_ERR_ARG = -3,         /* invalid argument */
  GDN_ERR_GRAPH = -4,       /* malformed CSR / needs reverse graph */
  GDN_ERR_IO = -5,          /* file not found / unreadable */
  GDN_ERR_NCCL = -6,
  GDN_ERR_NOMEM = -7
} gdn_status;

/* One step of the direction-optimizing controller (src/bfs/omp_beamer.cc:135-160). */
typedef struct {
  int32_t dir;          /* 0 top-down, 1 bottom-up */
  int32_t ns;           /* duration of the step on the device, nanoseconds (single-GPU path; 0 in partitioned mode) */
  int64_t frontier;     /* |frontier| expanded by this step */
  int64_t discovered;   /* vertices that received a depth */
  int64_t scout;        /* TD: scout_count, BU: awake_count */
  int64_t edges;        /* edges examined: TD = sum of out-degree over the frontier, BU = in-edges probed until the first hit */
  int64_t scanned;      /* BU: unvisited vertices swept (0 for TD) */
} gdn_bfs_step;

typedef struct {
  int32_t iterations;       /* the reference's `iterations = %d.` value */
  int32_t n_steps;          /* BFS: controller steps taken */
  double solve_ms;          /* CUDA-event time of the solve region, graph resident (cf. src/pr/base.cu:109-128) */
  double h2d_ms;            /* one-shot entry points: upload (+ preprocessing) wall time */
  double d2h_ms;            /* one-shot entry points: result download */
  int64_t kernel_launches;  /* kernels of ours launched inside the solve region */
  double kernel_ms;         /* CUDA-event time summed over the launches of the DOMINANT kernel of the solve
                               (PR/SpMV: gather_kernel; BFS: bu_sweep), measured on the library stream */
  int64_t kernel_calls;     /* number of launches kernel_ms covers */
  int64_t h2d_bytes, d2h_bytes;
  int64_t edges_reached;    /* BFS: sum of out-degree over reached vertices (directed entries) */
  int64_t vertices_reached; /* BFS */
  int64_t pr_layout;        /* PR: layout this solve walked (0 plain SELL / CSR, 1 banded, 2 segmented) */
  double pr_err[GDN_MAX_PR_ITER];          /* PR: per-iteration L1 delta (the reference's printed trace) */
  gdn_bfs_step steps[GDN_MAX_BFS_STEPS];   /* BFS */
} gdn_stats;

typedef struct gdn_graph gdn_graph;            /* device-resident CSR (+ row-block schedule) */
typedef struct gdn_host_graph gdn_host_graph;  /* host CSR produced by the readers / generator */

/* ---- library / device ------------------------------------------------------ */
int gdn_version(void);
const char *gdn_last_error(void);
/* Select the CUDA device (default 0) and create the library stream.  Fails with
 * GDN_ERR_NO_DEVICE when no sm_100 device is present. */
int gdn_init(int device);
/* Use the first `ngpus` GPUs of the box for the one-shot solvers (gdn_init(ngpus)): one worker thread per GPU
 * inside the library, 1-D row partition, the PageRank vector exchanged by peer-mapped stores over NVLink.  After this call
 * gdn_pagerank_pull / _i32 split every solve over the GPUs; everything else keeps running on device 0.  ngpus = 1 goes back
 * to one GPU.  (The resident API partitions across PROCESSES instead: gdn_comm_init, one rank per GPU.) */
int gdn_init_gpus(int ngpus);
int gdn_gpus(void);           /* GPUs the one-shot solvers use (1 unless gdn_init_gpus) */
int gdn_finalize(void);
int gdn_device_count(void);   /* 0 when no driver / device */
/* One-shot calls park their device blocks in a bounded arena instead of freeing them, so that the next
 * call on the same graph allocates nothing (csrc/pool.cu).  A process that shares the device with another allocator
 * calls this to hand the parked memory back; GDN_DEVICE_ARENA=0 in the environment turns the arena off. */
int gdn_device_trim(void);
/* PageRank summation order.  0 (default): rows are summed in the layout's order (heavy rows band by band / segment by
 * segment: within the 1e-6 L1 bar, not bit-identical).  1: graphs created from now on sum EVERY row sequentially in
 * column order, bit-identical to src/pr/omp_base.cc:28-30 (slower on skewed graphs: a hub row is one lane's work).
 * GDN_PR_EXACT=1 in the environment sets the default. */
int gdn_set_pr_exact_order(int on);

/* ---- one-shot solvers: HOST pointers, drop-in for the reference solvers ------
 * Upload, solve, download and free inside the call (ownership as in
 * src/pr/base.cu:83-101,133-139).  in_* may alias out_* (symmetrized graph,
 * include/csr_graph.h:241-246). */
int gdn_bfs(int64_t m, int64_t nnz,
            const uint64_t *out_rowptr, const int32_t *out_colidx,
            const uint64_t *in_rowptr, const int32_t *in_colidx,
            int32_t source, int32_t *depth_out /* m; GDN_INFINITY = unreached */,
            int32_t *parent_out /* nullable; -1 = unreached */, gdn_stats *st /* nullable */);
int gdn_bfs_i32(int32_t m, int32_t nnz,
                const int32_t *out_row_offsets, const int32_t *out_column_indices,
                const int32_t *in_row_offsets, const int32_t *in_column_indices,
                int32_t source, int32_t *depth_out, int32_t *parent_out, gdn_stats *st);

int gdn_pagerank_pull(int64_t m, int64_t nnz,
                      const uint64_t *in_rowptr, const int32_t *in_colidx,
                      const int32_t *out_degree /* m */, float *scores_inout /* m, pre-filled */,
                      float damp, double eps, int max_iter, gdn_stats *st);
int gdn_pagerank_pull_i32(int32_t m, int32_t nnz,
                          const int32_t *in_row_offsets, const int32_t *in_column_indices,
                          const int32_t *out_degree, float *scores_inout,
                          float damp, double eps, int max_iter, gdn_stats *st);

int gdn_spmv_csr(int64_t m, int64_t nnz, const uint64_t *Ap, const int32_t *Aj, const float *Ax,
                 const float *x, float *y_inout, gdn_stats *st);
int gdn_spmv_csr_i32(int32_t m, int32_t nnz, const int32_t *Ap, const int32_t *Aj, const float *Ax,
                     const float *x, float *y_inout, gdn_stats *st);

/* ---- resident API: graph stays in HBM across solves --------------------------
 * gdn_graph_create copies the CSR to the device (32-bit local offsets when
 * nnz < 2^32) and builds the row-block schedule the gather kernels use.
 * out_* may be NULL when only the pull kernels (PR/SpMV) will run; in_* may be
 * NULL or alias out_* for a symmetric graph.  row_lo/row_hi select a 1-D row
 * partition [row_lo,row_hi) of the m rows (pass 0,m for the whole graph). */
int gdn_graph_create(int64_t m, int64_t nnz,
                     const uint64_t *out_rowptr, const int32_t *out_colidx,
                     const uint64_t *in_rowptr, const int32_t *in_colidx,
                     int64_t row_lo, int64_t row_hi, gdn_graph **g);
/* gen-1 callers: int offsets (include/graph_io.h), whole graph. */
int gdn_graph_create_i32(int32_t m, int32_t nnz,
                         const int32_t *out_row_offsets, const int32_t *out_column_indices,
                         const int32_t *in_row_offsets, const int32_t *in_column_indices, gdn_graph **g);
/* PageRank divides by the OUT degree; a graph created from the in-CSR alone
 * (out_rowptr == NULL) must be told (host array of row_hi-row_lo entries). */
int gdn_graph_set_out_degree(gdn_graph *g, const int32_t *h_out_degree);
int gdn_graph_destroy(gdn_graph *g);
/* info[0]=m info[1]=nnz_local(in) info[2]=row_lo info[3]=row_hi info[4]=n_row_blocks
 * info[5]=n_heavy_segments info[6]=device_bytes info[7]=offset_bits */
int gdn_graph_info(const gdn_graph *g, int64_t info[8]);
/* PageRank pull layout of a resident graph (valid after the first gdn_pagerank_resident call):
 * info[0]=layout in use (0 plain, 1 banded, 2 segmented) info[1]=bands info[2]=ids per band info[3]=rows taking part
 * info[4]=column ids served from shared-memory bands info[5]=(row, band) pairs info[6]=band work items
 * info[7]=int4 groups of the main SELL array in use */
int gdn_graph_pull_info(const gdn_graph *g, int64_t info[8]);
/* Wall time of the (untimed, once per graph) preprocessing: ms[0] = gdn_graph_create (upload + host layout tables),
 * ms[1] = SELL array build, ms[2] = banded / segmented layout build, ms[3] = BFS hubs-first copy. */
int gdn_graph_prep_ms(const gdn_graph *g, double ms[4]);
/* Host-only probe of the banded layout's id -> band map over the id space
 * [hot prefix of H ids | P cold slices of Wc ids] (csrc/band.cu band_of / band_range): *band_out = band of new id `id`
 * (-1: beyond the first B bands), *local_out = its 16-bit index inside the band, *start_out / *len_out = the band's id range. */
/* Host-only probe of the banded layout's host tables (csrc/band.cu band_host_tables + band_host_check): builds them from
 * a count matrix cnt[B][n_rows], the remaining widths of the n_rows/32 band slices and the slice pointers of the plain
 * SELL array, and checks the invariants the kernels rely on.  0 = all hold, 1 = nothing qualifies, < 0 = broken invariant. */
int gdn_band_host_probe(int32_t B, int64_t n_rows, int32_t ids_per_unit, int32_t segmented, int32_t n_cta,
                        const uint32_t *cnt, const uint32_t *rem_w, const uint32_t *slice_ptr, int32_t n_slices,
                        int64_t *stats);
int gdn_band_map_probe(int64_t H, int64_t Wc, int32_t P, int32_t band, int32_t B, int64_t id, int32_t *band_out,
                       int32_t *local_out, int64_t *start_out, int32_t *len_out);

/* d_* are DEVICE pointers (cudaMalloc / torch tensors) of m elements; results
 * stay on the device.  Timed with CUDA events on the library stream. */
int gdn_bfs_resident(gdn_graph *g, int32_t source, int32_t *d_depth, int32_t *d_parent /* nullable */,
                     gdn_stats *st);
int gdn_pagerank_resident(gdn_graph *g, float *d_scores_inout, float damp, double eps, int max_iter,
                          gdn_stats *st);
/* y += A x over the in-CSR; d_Ax has nnz_local entries in in-edge order. */
int gdn_spmv_resident(gdn_graph *g, const float *d_Ax, const float *d_x, float *d_y_inout, gdn_stats *st);

/* Device buffers for callers without their own allocator. */
int gdn_dev_alloc(size_t bytes, void **d_ptr);
int gdn_dev_free(void *d_ptr);
int gdn_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes);
int gdn_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes);
int gdn_device_sync(void);
/* Page-lock / unlock a caller-owned host range so that the one-shot entry points
 * copy at full PCIe rate (cudaHostRegister). */
int gdn_host_pin(void *h_ptr, size_t bytes);
int gdn_host_unpin(void *h_ptr);

/* ---- 1-D row partition across the GPUs of one box -------------
 * Host-only: splits [0,m) into nparts contiguous, 64-aligned, equal-width
 * vertex ranges (NCCL allgather needs equal counts); bounds has nparts+1
 * entries, the last range is clipped to m.  width = bounds[1]. */
int gdn_partition_rows(int64_t m, int nparts, int64_t *bounds);
/* NCCL bootstrap: rank 0 calls gdn_comm_unique_id and broadcasts the 128 bytes
 * by any means (torch.distributed, a file); every rank then calls
 * gdn_comm_init.  One process per GPU. */
int gdn_comm_unique_id(uint8_t id[128]);
int gdn_comm_init(int rank, int nranks, const uint8_t id[128]);
int gdn_comm_destroy(void);
int gdn_comm_rank(void);
int gdn_comm_size(void);   /* 1 when no communicator */

/* ---- host graphs: readers and generator (no GPU needed) ----------------------- */
/* filetype "mtx" | "bin": gen-2 loader (include/csr_graph.h:211-250);
 * filetype "bin:mmap": the same binary triple mapped copy-on-write instead of read (symmetrize must be set): the
 *   processes of one box share one copy of the graph in the page cache;
 * filetype "sg": serialized graph of the GAP-style reader (include/reader.h:259-316; prefix or full path; offsets of
 *   4 bytes as the reference writes them or 8 as upstream GAP does, recognised by the file size; symmetrize is ignored,
 *   the file is a finished CSR -- include/builder.h:264);
 * filetype "auto": gen-1 loader dispatching on the suffix .mtx/.graph/.gr/.el
 * (include/graph_io.h:357-377), `prefix` is then the full path; a .sg path goes to the "sg" reader. */
int gdn_read_graph(const char *prefix, const char *filetype, int symmetrize, int need_reverse,
                   gdn_host_graph **hg);
/* kind 'g' = Kronecker (R-MAT), 'u' = uniform random; always symmetrized
 * (include/command_line.h:75-76). */
int gdn_generate(char kind, int scale, int degree, gdn_host_graph **hg);
/* The same graph as gdn_generate with the CSR built ON THE GPU (csrc/build.cu: 64-bit radix sort of both directions of
 * every edge, adjacent-unique, offsets by binary search) in place of the reference's host builder (include/builder.h:
 * 152-257).  ms (nullable): [0] edge streams on the host, [1] upload + keys, [2] sort, [3] unique + offsets, [4] download. */
int gdn_generate_gpu(char kind, int scale, int degree, gdn_host_graph **hg, double *ms);
/* Any edge list: n_edges (u, v) int32 pairs on the host -> symmetrized, squished CSR (include/builder.h:66-119,152-195). */
int gdn_build_csr_gpu(int64_t n_edges, const int32_t *pairs, gdn_host_graph **hg, double *ms);
int gdn_host_graph_free(gdn_host_graph *hg);
int64_t gdn_host_graph_m(const gdn_host_graph *hg);
int64_t gdn_host_graph_nnz(const gdn_host_graph *hg);
int gdn_host_graph_symmetric(const gdn_host_graph *hg);   /* in-CSR aliases out-CSR */
int gdn_host_graph_has_reverse(const gdn_host_graph *hg);
const uint64_t *gdn_host_graph_out_rowptr(const gdn_host_graph *hg);
const int32_t *gdn_host_graph_out_colidx(const gdn_host_graph *hg);
const uint64_t *gdn_host_graph_in_rowptr(const gdn_host_graph *hg);   /* NULL without reverse graph */
const int32_t *gdn_host_graph_in_colidx(const gdn_host_graph *hg);
const int32_t *gdn_host_graph_weights(const gdn_host_graph *hg);      /* gen-1 loader only, else NULL */
int gdn_host_graph_write_bin(const gdn_host_graph *hg, const char *prefix);
/* The .sg layout read by include/reader.h:259-316; offset_bytes 4 (the reference's SGOffset) or 8 (upstream GAP). */
int gdn_host_graph_write_sg(const gdn_host_graph *hg, const char *path, int offset_bytes);
/* OpenMP threads used by the host side (generator, readers, layout preprocessing).  Launchers such as
 * torchrun export OMP_NUM_THREADS=1, which would make the Kronecker generator 15x slower. */
int gdn_set_host_threads(int n);
/* Deterministic fp32 U[0,1) fill: std::mt19937(seed), (draw >> 8) * 2^-24. */
int gdn_fill_uniform(uint32_t seed, int64_t n, float *out);
/* BFS sources as SURVEY §8(d): mt19937(27491095) + uniform_int over [0,m-1],
 * rejecting degree-0 vertices. */
int gdn_pick_sources(const gdn_host_graph *hg, int n, int32_t *sources);

#ifdef __cplusplus
}
#endif
#endif /* GDN_B200_H_ */
