"""The METHOD of the banded / segmented PageRank layouts against the oracle, on the CPU (tests/band_emulation.py is a
plain-Python statement of what csrc/band.cu + pull.cu compute): summing a row band by band, restarting every work item
and meeting in an exact fixed-point accumulator, stays within the 1e-6 L1 bar and keeps the iteration count."""
import numpy as np
import pytest

import gardenia_b200 as gb
from oracle import pyoracle as po
from band_emulation import pagerank_banded


@pytest.mark.parametrize("kind,scale,kw", [
    ("g", 11, dict(n_bands=16, band=64, cmin=2, dmin=8, seg_ids=16, hot=512)),     # many small bands, short items
    ("g", 11, dict(n_bands=64, band=49152, cmin=4, dmin=64)),                       # the production parameters: one band
    ("u", 11, dict(n_bands=8, band=300, cmin=1, dmin=1, seg_ids=8, hot=512)),       # segmented mode: every id of every row
])
def test_banded_method_matches_oracle(kind, scale, kw):
    g = gb.Graph.generate(kind, scale, 16)
    m, rp, ci = g.m, g.out_rowptr(), g.out_colidx()
    scores, it, moved = pagerank_banded(m, rp, ci, **kw)
    oscores, oit, _ = po.pr_pull(m, rp, ci, g.out_degrees())
    assert moved > 0
    if kw.get("cmin") == 1 and kw.get("dmin") == 1:
        assert moved == g.nnz
    assert it == oit
    l1 = float(np.abs(scores.astype(np.float64) - oscores.astype(np.float64)).sum())
    assert l1 <= 1e-6, l1
