"""Solver entry points with the reference's names and argument meaning.

    BFSSolver(g, source, dist)        src/bfs/bfs.h:43   (dist pre-filled with MYINFINITY by the caller)
    PRSolver(g, scores)               src/pr/pr.h:31     (scores pre-filled with 1/m, in/out)
    SpmvSolver(g, Ax, x, y)           src/spmv/spmv.h:29 (y += A*x over the in-CSR, in/out)

Each call goes through the one-shot C-ABI (host pointers; upload, solve,
download inside the call) and prints the lines the reference's CUDA solvers
print (src/pr/base.cu:107,125,130-131; src/spmv/base.cu:69).  Errors follow the
reference's convention of refusing to continue, but as exceptions (GdnError)
rather than exit().
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .graph import Graph

MYINFINITY = _lib.GDN_INFINITY
EPSILON = 0.0001      # src/pr/pr.h:5
kDamp = 0.85          # src/pr/pr.h:6
MAX_ITER = 100        # src/pr/pr.h:12


def _need(a, dtype, n, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.size < n or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name} must be a contiguous numpy {np.dtype(dtype).name} array of >= {n} elements")


def BFSSolver(g: Graph, source: int, dist: np.ndarray, parent: np.ndarray = None, verbose=True):
    if not g.has_reverse_graph():
        # src/bfs/omp_beamer.cc:98-102
        raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "This algorithm requires the reverse graph constructed for directed graph")
    _need(dist, np.int32, g.m, "dist")
    if parent is not None:
        _need(parent, np.int32, g.m, "parent")
    if verbose:
        print("Launching CUDA BFS solver (sm_100a, direction-optimizing) ...")
    st = _lib.Stats()
    check(lib.gdn_bfs(g.m, g.nnz, _lib.ptr(g.out_rowptr()), _lib.ptr(g.out_colidx()),
                      _lib.ptr(g.in_rowptr()), _lib.ptr(g.in_colidx()), int(source),
                      _lib.ptr(dist), _lib.ptr(parent), C.byref(st)))
    if verbose:
        print(f"\titerations = {st.iterations}.")
        print(f"\truntime [b200_hybrid] = {st.solve_ms:f} ms.")
    return st


def PRSolver(g: Graph, scores: np.ndarray, verbose=True):
    if not g.has_reverse_graph():
        raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "PageRank pull needs the reverse graph")
    _need(scores, np.float32, g.m, "scores")
    if verbose:
        print("Launching CUDA PR solver (sm_100a, pull) ...")
    st = _lib.Stats()
    out_degree = g.out_degrees()
    check(lib.gdn_pagerank_pull(g.m, g.nnz, _lib.ptr(g.in_rowptr()), _lib.ptr(g.in_colidx()),
                                _lib.ptr(out_degree), _lib.ptr(scores), kDamp, EPSILON, MAX_ITER, C.byref(st)))
    if verbose:
        for i, e in enumerate(st.pr_trace()):
            print(" %2d    %f" % (i + 1, e))
        print(f"\titerations = {st.iterations}.")
        print(f"\truntime [b200_pull] = {st.solve_ms:f} ms.")
    return st


def SpmvSolver(g: Graph, Ax: np.ndarray, x: np.ndarray, y: np.ndarray, verbose=True):
    if not g.has_reverse_graph():
        raise _lib.GdnError(_lib.GDN_ERR_GRAPH, "SpMV runs over the in-CSR: load with symmetrize=1 or reverse=1")
    _need(Ax, np.float32, g.nnz, "Ax")
    _need(x, np.float32, g.m, "x")
    _need(y, np.float32, g.m, "y")
    if verbose:
        print("Launching CUDA SpMV solver (sm_100a) ...")
    st = _lib.Stats()
    check(lib.gdn_spmv_csr(g.m, g.nnz, _lib.ptr(g.in_rowptr()), _lib.ptr(g.in_colidx()), _lib.ptr(Ax),
                           _lib.ptr(x), _lib.ptr(y), C.byref(st)))
    if verbose and st.solve_ms > 0:
        gflops = 2.0 * g.nnz / st.solve_ms / 1e6
        gbytes = (16.0 * g.m + 12.0 * g.nnz) / st.solve_ms / 1e6      # src/spmv/spmv_util.h:6-13 ("reference model")
        print("\truntime [b200_csr] = %.4f ms ( %5.2f GFLOP/s %5.1f GB/s)" % (st.solve_ms, gflops, gbytes))
    return st
