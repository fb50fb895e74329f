"""ctypes binding of libgdn_b200.so (include/gdn_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C gardenia_b200/csrc``.  There is no Python or CPU fallback: if the
library is missing, importing this module raises; if no B200 is present, every
compute entry point returns GDN_ERR_NO_DEVICE and the wrappers raise GdnError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgdn_b200.so")

GDN_MAX_PR_ITER = 128
GDN_MAX_BFS_STEPS = 256
GDN_INFINITY = 1000000000

GDN_OK = 0
GDN_ERR_NO_DEVICE = -1
GDN_ERR_CUDA = -2
GDN_ERR_ARG = -3
GDN_ERR_GRAPH = -4
GDN_ERR_IO = -5
GDN_ERR_NCCL = -6
GDN_ERR_NOMEM = -7


class GdnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgdn_b200 error {code}: {msg}")
        self.code = code


class BfsStep(C.Structure):
    _fields_ = [("dir", C.c_int32), ("ns", C.c_int32), ("frontier", C.c_int64),
                ("discovered", C.c_int64), ("scout", C.c_int64), ("edges", C.c_int64), ("scanned", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("n_steps", C.c_int32),
                ("solve_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("kernel_ms", C.c_double), ("kernel_calls", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("edges_reached", C.c_int64), ("vertices_reached", C.c_int64), ("pr_layout", C.c_int64),
                ("pr_err", C.c_double * GDN_MAX_PR_ITER),
                ("steps", BfsStep * GDN_MAX_BFS_STEPS)]

    def pr_trace(self):
        n = min(self.iterations, GDN_MAX_PR_ITER)
        return [self.pr_err[i] for i in range(n)]

    def bfs_steps(self):
        n = min(self.n_steps, GDN_MAX_BFS_STEPS)
        return [dict(dir=s.dir, ns=s.ns, frontier=s.frontier, discovered=s.discovered, scout=s.scout, edges=s.edges, scanned=s.scanned)
                for s in self.steps[:n]]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback)")

lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_i64, _i32, _f32, _f64 = C.c_int64, C.c_int32, C.c_float, C.c_double
_SP = C.POINTER(Stats)

# name -> (restype, argtypes); every name here must be declared in include/gdn_b200.h
SIGNATURES = {
    "gdn_version": (C.c_int, []),
    "gdn_last_error": (C.c_char_p, []),
    "gdn_init": (C.c_int, [C.c_int]),
    "gdn_init_gpus": (C.c_int, [C.c_int]),
    "gdn_gpus": (C.c_int, []),
    "gdn_finalize": (C.c_int, []),
    "gdn_device_count": (C.c_int, []),
    "gdn_device_trim": (C.c_int, []),
    "gdn_set_pr_exact_order": (C.c_int, [C.c_int]),
    "gdn_bfs": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _SP]),
    "gdn_bfs_i32": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _SP]),
    "gdn_pagerank_pull": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _f32, _f64, C.c_int, _SP]),
    "gdn_pagerank_pull_i32": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _f32, _f64, C.c_int, _SP]),
    "gdn_spmv_csr": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _SP]),
    "gdn_spmv_csr_i32": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _SP]),
    "gdn_graph_create": (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(_vp)]),
    "gdn_graph_create_i32": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "gdn_graph_set_out_degree": (C.c_int, [_vp, _vp]),
    "gdn_graph_destroy": (C.c_int, [_vp]),
    "gdn_graph_info": (C.c_int, [_vp, C.POINTER(_i64 * 8)]),
    "gdn_graph_pull_info": (C.c_int, [_vp, C.POINTER(_i64 * 8)]),
    "gdn_graph_prep_ms": (C.c_int, [_vp, C.POINTER(C.c_double * 4)]),
    "gdn_band_host_probe": (C.c_int, [C.c_int32, _i64, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_int32, _vp]),
    "gdn_band_map_probe": (C.c_int, [_i64, _i64, C.c_int32, C.c_int32, C.c_int32, _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(_i64), C.POINTER(C.c_int32)]),
    "gdn_bfs_resident": (C.c_int, [_vp, _i32, _vp, _vp, _SP]),
    "gdn_pagerank_resident": (C.c_int, [_vp, _vp, _f32, _f64, C.c_int, _SP]),
    "gdn_spmv_resident": (C.c_int, [_vp, _vp, _vp, _vp, _SP]),
    "gdn_dev_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "gdn_dev_free": (C.c_int, [_vp]),
    "gdn_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "gdn_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_size_t]),
    "gdn_device_sync": (C.c_int, []),
    "gdn_host_pin": (C.c_int, [_vp, C.c_size_t]),
    "gdn_host_unpin": (C.c_int, [_vp]),
    "gdn_partition_rows": (C.c_int, [_i64, C.c_int, _vp]),
    "gdn_comm_unique_id": (C.c_int, [_vp]),
    "gdn_comm_init": (C.c_int, [C.c_int, C.c_int, _vp]),
    "gdn_comm_destroy": (C.c_int, []),
    "gdn_comm_rank": (C.c_int, []),
    "gdn_comm_size": (C.c_int, []),
    "gdn_read_graph": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gdn_generate": (C.c_int, [C.c_char, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gdn_generate_gpu": (C.c_int, [C.c_char, C.c_int, C.c_int, C.POINTER(_vp), _vp]),
    "gdn_build_csr_gpu": (C.c_int, [_i64, _vp, C.POINTER(_vp), _vp]),
    "gdn_host_graph_free": (C.c_int, [_vp]),
    "gdn_host_graph_m": (_i64, [_vp]),
    "gdn_host_graph_nnz": (_i64, [_vp]),
    "gdn_host_graph_symmetric": (C.c_int, [_vp]),
    "gdn_host_graph_has_reverse": (C.c_int, [_vp]),
    "gdn_host_graph_out_rowptr": (_vp, [_vp]),
    "gdn_host_graph_out_colidx": (_vp, [_vp]),
    "gdn_host_graph_in_rowptr": (_vp, [_vp]),
    "gdn_host_graph_in_colidx": (_vp, [_vp]),
    "gdn_host_graph_weights": (_vp, [_vp]),
    "gdn_host_graph_write_bin": (C.c_int, [_vp, C.c_char_p]),
    "gdn_host_graph_write_sg": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "gdn_set_host_threads": (C.c_int, [C.c_int]),
    "gdn_fill_uniform": (C.c_int, [C.c_uint32, _i64, _vp]),
    "gdn_pick_sources": (C.c_int, [_vp, C.c_int, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.gdn_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != GDN_OK:
        raise GdnError(rc, last_error())
    return rc


def ptr(a):
    """Pointer of a numpy array / torch tensor / raw int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data
