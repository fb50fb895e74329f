"""Host and device graphs.

``Graph`` mirrors the constructor contract of the reference's gen-2 ``Graph``
(include/csr_graph.h:211-250): ``Graph(prefix, filetype, symmetrize,
need_reverse)``, ``V() E() out_rowptr() out_colidx() in_rowptr() in_colidx()
has_reverse_graph() get_degree(v)``.  ``Graph.generate('g'|'u', scale, degree)``
is the Kronecker/uniform generator (include/builder.h:258-274).  Arrays are
zero-copy numpy views of the C++ host graph.

``DeviceGraph`` keeps the CSR resident in HBM (gdn_graph_create) so that
repeated solves time only the kernels, like the reference's timed region
(src/pr/base.cu:109-128).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check


def _view(addr, n, dtype):
    if not addr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Graph:
    def __init__(self, prefix=None, filetype="bin", symmetrize=False, need_reverse=False, _handle=None):
        self._h = None
        if _handle is None:
            h = C.c_void_p()
            check(lib.gdn_read_graph(str(prefix).encode(), filetype.encode(), int(symmetrize),
                                     int(need_reverse), C.byref(h)))
            _handle = h
        self._h = _handle
        self._bind()

    @classmethod
    def generate(cls, kind, scale, degree=16):
        """kind 'g' = Kronecker (R-MAT, -g), 'u' = uniform random (-u); always symmetrized."""
        h = C.c_void_p()
        check(lib.gdn_generate(kind.encode()[0:1], int(scale), int(degree), C.byref(h)))
        return cls(_handle=h)

    @classmethod
    def generate_gpu(cls, kind, scale, degree=16):
        """Same graph as generate(), its CSR built on the GPU (csrc/build.cu); .build_ms holds the stage times."""
        h = C.c_void_p()
        ms = (C.c_double * 5)()
        check(lib.gdn_generate_gpu(kind.encode()[0:1], int(scale), int(degree), C.byref(h), C.cast(ms, C.c_void_p)))
        g = cls(_handle=h)
        g.build_ms = dict(zip(["edge_streams_host", "upload_keys", "sort", "unique_offsets", "download"], [float(x) for x in ms]))
        return g

    @classmethod
    def from_edges(cls, pairs):
        """Symmetrized, squished CSR of an (n, 2) int32 edge array, built on the GPU."""
        import numpy as np
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        h = C.c_void_p()
        check(lib.gdn_build_csr_gpu(len(pairs), pairs.ctypes.data, C.byref(h), None))
        return cls(_handle=h)

    @classmethod
    def from_file(cls, path, symmetrize=False):
        """gen-1 reader dispatching on the suffix (.mtx/.graph/.gr/.el), include/graph_io.h:357-377."""
        h = C.c_void_p()
        check(lib.gdn_read_graph(str(path).encode(), b"auto", int(symmetrize), 0, C.byref(h)))
        return cls(_handle=h)

    def _bind(self):
        h = self._h
        self.m = int(lib.gdn_host_graph_m(h))
        self.nnz = int(lib.gdn_host_graph_nnz(h))
        self.symmetric = bool(lib.gdn_host_graph_symmetric(h))
        self._out_rowptr = _view(lib.gdn_host_graph_out_rowptr(h), self.m + 1, np.uint64)
        self._out_colidx = _view(lib.gdn_host_graph_out_colidx(h), self.nnz, np.int32)
        ir = lib.gdn_host_graph_in_rowptr(h)
        self._has_reverse = bool(ir)
        if self._has_reverse:
            self._in_rowptr = _view(ir, self.m + 1, np.uint64)
            self._in_colidx = _view(lib.gdn_host_graph_in_colidx(h), self.nnz, np.int32)
        else:
            self._in_rowptr = self._in_colidx = None
        w = lib.gdn_host_graph_weights(h)
        self.weights = _view(w, self.nnz, np.int32) if w else None

    def __del__(self):
        if getattr(self, "_h", None):
            lib.gdn_host_graph_free(self._h)
            self._h = None

    # reference accessor names
    def V(self): return self.m
    def E(self): return self.nnz
    def out_rowptr(self): return self._out_rowptr
    def out_colidx(self): return self._out_colidx
    def in_rowptr(self): return self._in_rowptr
    def in_colidx(self): return self._in_colidx
    def has_reverse_graph(self): return self._has_reverse
    def get_degree(self, v): return int(self._out_rowptr[v + 1] - self._out_rowptr[v])
    def out_degrees(self):
        """int32 out-degree of every vertex (the `degree` array of the gen-1 readers, include/graph_io.h:357-377);
        computed once per Graph, like the reference's reader does at load time."""
        if getattr(self, "_deg", None) is None:
            self._deg = np.diff(self._out_rowptr).astype(np.int32)
        return self._deg

    def write_bin(self, prefix):
        check(lib.gdn_host_graph_write_bin(self._h, str(prefix).encode()))

    def write_sg(self, path, offset_bytes=4):
        """The serialized graph the GAP-style drivers read (include/reader.h:259-316)."""
        check(lib.gdn_host_graph_write_sg(self._h, str(path).encode(), int(offset_bytes)))

    def pick_sources(self, n=16):
        """GAP-style BFS sources: mt19937(27491095), degree-0 rejected."""
        out = np.zeros(n, dtype=np.int32)
        check(lib.gdn_pick_sources(self._h, n, out.ctypes.data))
        return out


def init_gpus(n):
    """Split every one-shot solve (PRSolver) over the first n GPUs of the box: gdn_init_gpus, one worker thread per GPU
    inside the library; n = 1 goes back to one GPU."""
    check(lib.gdn_init_gpus(int(n)))


def fill_uniform(seed, n):
    """fp32 U[0,1) from std::mt19937(seed): (draw >> 8) * 2**-24 (same stream as oracle/ref_driver.cc)."""
    out = np.empty(n, dtype=np.float32)
    check(lib.gdn_fill_uniform(int(seed), int(n), out.ctypes.data))
    return out


def partition_rows(m, nparts):
    b = np.zeros(nparts + 1, dtype=np.int64)
    check(lib.gdn_partition_rows(int(m), int(nparts), b.ctypes.data))
    return b


class DeviceGraph:
    """CSR resident on the current device (optionally one row partition of it)."""

    def __init__(self, g: Graph, row_lo=0, row_hi=None, device=0):
        check(lib.gdn_init(device))
        self.host = g
        self.m = g.m
        row_hi = g.m if row_hi is None else row_hi
        h = C.c_void_p()
        in_rp = g.in_rowptr() if (g.has_reverse_graph() and not g.symmetric) else None
        in_ci = g.in_colidx() if (g.has_reverse_graph() and not g.symmetric) else None
        if g.has_reverse_graph() and g.symmetric:
            in_rp, in_ci = g.out_rowptr(), g.out_colidx()
        check(lib.gdn_graph_create(g.m, g.nnz, _lib.ptr(g.out_rowptr()), _lib.ptr(g.out_colidx()),
                                   _lib.ptr(in_rp), _lib.ptr(in_ci), int(row_lo), int(row_hi), C.byref(h)))
        self._h = h
        self.row_lo, self.row_hi = int(row_lo), int(row_hi)
        if not g.has_reverse_graph():
            # only the forward CSR exists: it was uploaded as "symmetric"; BFS DO
            # would be wrong on it, so remember to refuse (src/bfs/omp_beamer.cc:98-102).
            self._no_reverse = True
        else:
            self._no_reverse = False

    def info(self):
        a = (C.c_int64 * 8)()
        check(lib.gdn_graph_info(self._h, C.byref(a)))
        keys = ["m", "nnz_local", "row_lo", "row_hi", "n_row_blocks", "n_heavy_segments", "device_bytes", "offset_bits"]
        return dict(zip(keys, list(a)))

    def pull_info(self):
        """PageRank pull layout in use (banded shared-memory layout of the heavy rows, csrc/band.cu)."""
        a = (C.c_int64 * 8)()
        check(lib.gdn_graph_pull_info(self._h, C.byref(a)))
        keys = ["banded", "bands", "band_ids", "band_rows", "band_entries", "band_pairs", "band_items", "main_groups"]
        return dict(zip(keys, list(a)))

    def prep_ms(self):
        """Wall time of the once-per-graph preprocessing (untimed by the solves): create, SELL build, band build, BFS copy."""
        a = (C.c_double * 4)()
        check(lib.gdn_graph_prep_ms(self._h, C.byref(a)))
        return dict(zip(["create", "sell_build", "band_build", "bfs_hubs_first"], [float(x) for x in a]))

    def close(self):
        if getattr(self, "_h", None):
            lib.gdn_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    @staticmethod
    def _sync_inputs(*tensors):
        # the library runs on its own stream; make sure the producer of any torch
        # tensor handed in has finished before our kernels read it
        for t in tensors:
            if t is not None and getattr(t, "is_cuda", False):
                import torch
                torch.cuda.current_stream(t.device).synchronize()
                return

    # --- resident solvers: torch CUDA tensors in/out, results stay on the device
    def bfs(self, source, depth, parent=None):
        self._sync_inputs(depth, parent)
        if self._no_reverse:
            raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "This algorithm requires the reverse graph constructed for directed graph")
        st = _lib.Stats()
        check(lib.gdn_bfs_resident(self._h, int(source), _lib.ptr(depth), _lib.ptr(parent), C.byref(st)))
        return st

    def pagerank(self, scores, damp=0.85, eps=1e-4, max_iter=100):
        self._sync_inputs(scores)
        if self._no_reverse:
            raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "PageRank pull needs the reverse graph")
        st = _lib.Stats()
        check(lib.gdn_pagerank_resident(self._h, _lib.ptr(scores), damp, eps, int(max_iter), C.byref(st)))
        return st

    def spmv(self, Ax, x, y):
        self._sync_inputs(Ax, x, y)
        if self._no_reverse:
            raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "SpMV runs over the in-CSR: load with symmetrize=1 or reverse=1")
        st = _lib.Stats()
        check(lib.gdn_spmv_resident(self._h, _lib.ptr(Ax), _lib.ptr(x), _lib.ptr(y), C.byref(st)))
        return st
