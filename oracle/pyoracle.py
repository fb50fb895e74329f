"""ctypes binding of oracle/liboracle.so (the plain-C restatement, gdn_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never by gardenia_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
INFINITY = 1000000000


class BfsStep(C.Structure):
    _fields_ = [("dir", C.c_int32), ("pad", C.c_int32), ("frontier", C.c_int64), ("edges", C.c_int64),
                ("scanned", C.c_int64), ("discovered", C.c_int64), ("scout", C.c_int64)]


def build():
    """(Re)build liboracle.so -- and oracle/_ref when /root/reference is present."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.oracle_bfs_do.restype = C.c_int
    lib.oracle_bfs_do.argtypes = [i64, vp, vp, vp, vp, i32, vp, vp, C.c_int, C.POINTER(C.c_int)]
    lib.oracle_bfs_td.restype = C.c_int
    lib.oracle_bfs_td.argtypes = [i64, vp, vp, i32, vp]
    lib.oracle_bfs_verify.restype = i64
    lib.oracle_bfs_verify.argtypes = [i64, vp, vp, i32, vp]
    lib.oracle_bfs_check_parents.restype = i64
    lib.oracle_bfs_check_parents.argtypes = [i64, vp, vp, i32, vp, vp]
    lib.oracle_pr_pull.restype = C.c_int
    lib.oracle_pr_pull.argtypes = [i64, vp, vp, vp, vp, C.c_float, C.c_double, C.c_int, vp]
    lib.oracle_pr_residual.restype = C.c_double
    lib.oracle_pr_residual.argtypes = [i64, vp, vp, vp, C.c_float]
    lib.oracle_spmv.restype = None
    lib.oracle_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
    lib.oracle_max_relative_error.restype = C.c_float
    lib.oracle_max_relative_error.argtypes = [vp, vp, i64]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def bfs_do(m, out_rowptr, out_colidx, in_rowptr, in_colidx, source, max_steps=4096):
    """-> (dist int32[m], iterations, steps list)"""
    dist = np.empty(m, dtype=np.int32)
    steps = (BfsStep * max_steps)()
    n = C.c_int(0)
    it = lib().oracle_bfs_do(m, _p(out_rowptr), _p(out_colidx), _p(in_rowptr), _p(in_colidx), int(source),
                             _p(dist), C.cast(steps, C.c_void_p), max_steps, C.byref(n))
    if it < 0:
        raise ValueError("oracle_bfs_do: bad arguments")
    out = [dict(dir=s.dir, frontier=s.frontier, edges=s.edges, scanned=s.scanned, discovered=s.discovered,
                scout=s.scout) for s in steps[:min(n.value, max_steps)]]
    return dist, it, out


def bfs_td(m, rowptr, colidx, source):
    dist = np.empty(m, dtype=np.int32)
    it = lib().oracle_bfs_td(m, _p(rowptr), _p(colidx), int(source), _p(dist))
    return dist, it


def bfs_verify(m, rowptr, colidx, source, dist):
    return int(lib().oracle_bfs_verify(m, _p(rowptr), _p(colidx), int(source), _p(dist)))


def bfs_check_parents(m, in_rowptr, in_colidx, source, dist, parent):
    return int(lib().oracle_bfs_check_parents(m, _p(in_rowptr), _p(in_colidx), int(source), _p(dist), _p(parent)))


def pr_pull(m, in_rowptr, in_colidx, out_degree, scores=None, damp=0.85, eps=1e-4, max_iter=100):
    """-> (scores float32[m], iterations, err_trace list)"""
    if scores is None:
        scores = np.full(m, np.float32(1.0) / np.float32(m), dtype=np.float32)   # src/pr/main.cc:17-18
    else:
        scores = scores.copy()
    trace = np.zeros(max_iter + 1, dtype=np.float64)
    it = lib().oracle_pr_pull(m, _p(in_rowptr), _p(in_colidx), _p(np.ascontiguousarray(out_degree, dtype=np.int32)),
                              _p(scores), damp, eps, max_iter, _p(trace))
    return scores, it, list(trace[:min(it, max_iter)])


def pr_residual(m, out_rowptr, out_colidx, scores, damp=0.85):
    return float(lib().oracle_pr_residual(m, _p(out_rowptr), _p(out_colidx), _p(scores), damp))


def spmv(m, Ap, Aj, Ax, x, y):
    y = y.copy()
    lib().oracle_spmv(m, _p(Ap), _p(Aj), _p(Ax), _p(x), _p(y))
    return y


def max_relative_error(a, b):
    return float(lib().oracle_max_relative_error(_p(a), _p(b), len(a)))
