"""The oracle (oracle/gdn_oracle.c) pinned against the reference:
  * the reference's own golden PR trace (test/reference/graph-pr.mtx.out:13-28),
  * outputs of the reference's OpenMP sources run in the build container
    (tests/golden/*.ref.npz, produced by tools/make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_case
from oracle import pyoracle as po


def test_reference_golden_pr_trace():
    gold = json.load(open(os.path.join(GOLDEN, "pr_trace_test_pr.json")))
    csr, _ = load_case("test_pr_dir")
    deg = np.diff(csr["out_rowptr"]).astype(np.int32)
    scores, iters, trace = po.pr_pull(csr["m"], csr["in_rowptr"], csr["in_colidx"], deg)
    assert iters == gold["iterations"] == 15
    # the reference prints %lf (6 decimals): compare at printed precision
    assert [round(t, 6) for t in trace] == gold["trace"]
    # KAT recorded in SURVEY §8(c)
    np.testing.assert_allclose(scores, [0.364168763, 0.196846426, 0.192254141, 0.246730655], rtol=0, atol=1e-9)


def test_pr_matches_reference(case):
    name, csr, ref = case
    deg = np.diff(csr["out_rowptr"]).astype(np.int32)
    scores, iters, trace = po.pr_pull(csr["m"], csr["in_rowptr"], csr["in_colidx"], deg)
    assert iters == int(ref["pr_iters"])
    assert np.array_equal(scores, ref["pr_scores"]), "oracle PR scores are not bit-identical to pr_omp_base"
    assert [round(t, 6) for t in trace] == [round(float(t), 6) for t in ref["pr_trace"]]
    # the reference's own acceptance test (src/pr/verifier.cc:40-54)
    assert po.pr_residual(csr["m"], csr["out_rowptr"], csr["out_colidx"], scores) < 1e-4


def test_bfs_matches_reference(case):
    name, csr, ref = case
    for s in ref["sources"]:
        dist, iters, steps = po.bfs_do(csr["m"], csr["out_rowptr"], csr["out_colidx"], csr["in_rowptr"],
                                       csr["in_colidx"], int(s))
        assert np.array_equal(dist, ref[f"bfs_dist_{s}"]), f"{name} source {s}"
        assert iters == int(ref[f"bfs_iters_{s}"])
        # independent restatements agree: TD-only BFS and the reference verifier
        dist_td, _ = po.bfs_td(csr["m"], csr["out_rowptr"], csr["out_colidx"], int(s))
        assert np.array_equal(dist, dist_td)
        assert po.bfs_verify(csr["m"], csr["out_rowptr"], csr["out_colidx"], int(s), dist) == 0
        assert sum(st["discovered"] for st in steps) + 1 == int((dist != po.INFINITY).sum())


def test_spmv_matches_reference(case):
    import gardenia_b200 as gb
    name, csr, ref = case
    m, nnz = csr["m"], csr["nnz"]
    # same stream as oracle/ref_driver.cc: one mt19937(13), Ax first then x
    both = gb.fill_uniform(13, nnz + m)
    Ax, x = both[:nnz].copy(), both[nnz:].copy()
    y = po.spmv(m, csr["in_rowptr"], csr["in_colidx"], Ax, x, np.zeros(m, dtype=np.float32))
    assert np.array_equal(y, ref["spmv_y"]), "oracle SpMV is not bit-identical to spmv_omp_base"
    assert po.max_relative_error(y, ref["spmv_y"]) == 0.0


def test_bfs_kat_4mtx():
    # 4.mtx symmetrized, source 0
    csr, _ = load_case("4_sym")
    dist, _, _ = po.bfs_do(csr["m"], csr["out_rowptr"], csr["out_colidx"], csr["in_rowptr"], csr["in_colidx"], 0)
    INF = po.INFINITY
    assert list(dist) == [0, 1, 1, 1, 1, INF, 1, 1, 1, 1, 1, 1, 1, 1]


def test_spmv_kat_4mtx():
    # Ax=.2, x=.3, y0=0 (the constants of src/spmv/main.cc:27-37)
    csr, _ = load_case("4_sym")
    m, nnz = csr["m"], csr["nnz"]
    y = po.spmv(m, csr["in_rowptr"], csr["in_colidx"], np.full(nnz, 0.2, np.float32), np.full(m, 0.3, np.float32),
                np.zeros(m, np.float32))
    want = [.72, .42, .6, .48, .48, 0, .24, .66, .6, .66, .36, .3, .3, .54]
    np.testing.assert_allclose(y, want, rtol=1e-6, atol=1e-7)


def test_bfs_bad_source():
    csr, _ = load_case("4_sym")
    with pytest.raises(ValueError):
        po.bfs_do(csr["m"], csr["out_rowptr"], csr["out_colidx"], csr["in_rowptr"], csr["in_colidx"], 99)


# ------------------------------------------------------------------ KATs recorded from the reference itself
# (bfs_omp_beamer with its own commented printf lines re-enabled, source 0; pr_omp_base; the reference generator)
_SCHEDULE_KATS = {
    # kind: (m, nnz, max degree, [(direction, discovered, scout_count)], BFS iterations, PR iterations)
    "g": (1048576, 31399374, 64637,
          [(0, 1, 1315), (0, 1314, 2676163), (1, 339311, None), (1, 301458, None), (1, 3171, None), (0, 13, 13), (0, 0, 0)], 7, 8),
    "u": (1048576, 33553952, 62,
          [(0, 29, 910), (0, 880, 29325), (0, 28045, 923290), (0, 585818, 19099150), (1, 433803, None), (1, 0, None)], 6, 6),
}


@pytest.mark.parametrize("kind", ["g", "u"])
def test_scale20_schedule_kats_of_the_reference(kind):
    """Generator + oracle reproduce what the reference printed at scale 20: graph size, the direction-optimizing
    schedule level by level (direction, next frontier, scout_count), `iterations`, and PageRank's iteration count."""
    import gardenia_b200 as gb
    m, nnz, maxdeg, sched, bfs_it, pr_it = _SCHEDULE_KATS[kind]
    g = gb.Graph.generate(kind, 20, 16)
    rp, ci = g.out_rowptr(), g.out_colidx()
    assert (g.m, g.nnz, int(np.diff(rp).max())) == (m, nnz, maxdeg)
    dist, it, steps = po.bfs_do(g.m, rp, ci, rp, ci, 0)
    assert it == bfs_it and len(steps) == len(sched)
    for s, (d, disc, scout) in zip(steps, sched):
        assert (s["dir"], s["discovered"]) == (d, disc), (s, d, disc)
        if scout is not None:
            assert s["scout"] == scout
    assert po.bfs_verify(g.m, rp, ci, 0, dist) == 0
    _, it_pr, _ = po.pr_pull(g.m, rp, ci, g.out_degrees())
    assert it_pr == pr_it


def test_kron22_size_kat():
    """Config C2 of BASELINE.json: `-g 22` has 4,194,302 vertices (m = max id + 1, include/builder.h:244-245), 128,311,436
    directed entries and maximum degree 162,839; BFS from source 0 takes 7 levels, PageRank 7 iterations."""
    import gardenia_b200 as gb
    g = gb.Graph.generate("g", 22, 16)
    rp, ci = g.out_rowptr(), g.out_colidx()
    assert (g.m, g.nnz, int(np.diff(rp).max())) == (4194302, 128311436, 162839)
    _, it, _ = po.bfs_do(g.m, rp, ci, rp, ci, 0)
    assert it == 7
    _, it_pr, _ = po.pr_pull(g.m, rp, ci, g.out_degrees())
    assert it_pr == 7
