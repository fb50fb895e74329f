#!/usr/bin/env python
"""e2e_trace.py -- stage timings (GDN_TRACE) of the one-shot PageRank entry point on a cached graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GDN_TRACE"] = "1"
import numpy as np
import bench
import gardenia_b200 as gb
from gardenia_b200 import _lib
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
_, g = bench.load_graph("g", scale)
for a in (g.out_rowptr(), g.out_colidx()):
    _lib.lib.gdn_host_pin(a.ctypes.data, a.nbytes)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    s = np.full(g.m, np.float32(1.0) / np.float32(g.m), dtype=np.float32)
    t = time.time()
    st = gb.PRSolver(g, s, verbose=False)
    print(f"rep {rep}: wall {1e3 * (time.time() - t):.1f} ms  h2d {st.h2d_ms:.1f} solve {st.solve_ms:.1f} d2h {st.d2h_ms:.1f} iters {st.iterations}", file=sys.stderr)
