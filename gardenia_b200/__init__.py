"""gardenia_b200 -- B200-native (sm_100a) CSR traversal engine behind Gardenia's
solver API: direction-optimizing BFS and the pull-direction CSR gather shared
by PageRank and fp32 SpMV.  See DESIGN.md.

The compute path is libgdn_b200.so (hand-written CUDA); this package is the
thin host-side mirror of the reference's interface.  No CPU fallback exists.
"""
from ._lib import GdnError, Stats, GDN_INFINITY  # noqa: F401
from .graph import Graph, DeviceGraph, fill_uniform, partition_rows, init_gpus  # noqa: F401
from .solvers import BFSSolver, PRSolver, SpmvSolver, MYINFINITY, EPSILON, kDamp, MAX_ITER  # noqa: F401
