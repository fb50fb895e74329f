"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_pr_bfs(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                        os.path.join(ROOT, "tests", "_multi_worker.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


@pytest.mark.parametrize("ngpus", [2, 4, 8])
def test_gang_pagerank_one_process(ngpus):
    """gdn_init_gpus(n): the one-shot PRSolver call, unchanged, runs on n GPUs of the box from ONE process (a worker
    thread per GPU inside the library).  Same bar as on one GPU: iteration count and 1e-6 L1 against the oracle."""
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    import numpy as np
    import gardenia_b200 as gb
    from oracle import pyoracle as po
    gb.init_gpus(ngpus)
    try:
        for kind, scale in (("g", 20), ("u", 20)):
            g = gb.Graph.generate(kind, scale, 16)
            m, rp, ci = g.m, g.out_rowptr(), g.out_colidx()
            oscores, oit, _ = po.pr_pull(m, rp, ci, g.out_degrees())
            for _ in range(2):                       # the second call reuses the workers' parked device blocks
                hs = np.full(m, np.float32(1.0) / np.float32(m), dtype=np.float32)
                st = gb.PRSolver(g, hs, verbose=False)
                assert st.iterations == oit
                assert float(np.abs(hs.astype(np.float64) - oscores.astype(np.float64)).sum()) <= 1e-6
        # a small graph stays on one GPU
        g = gb.Graph.generate("g", 12, 16)
        hs = np.full(g.m, np.float32(1.0) / np.float32(g.m), dtype=np.float32)
        st = gb.PRSolver(g, hs, verbose=False)
        o, oit, _ = po.pr_pull(g.m, g.out_rowptr(), g.out_colidx(), g.out_degrees())
        assert st.iterations == oit and float(np.abs(hs.astype(np.float64) - o.astype(np.float64)).sum()) <= 1e-6
    finally:
        gb.init_gpus(1)
