"""The ordered sum (csrc/ordered_core.cuh), stated in numpy (tests/ordered_sum_emulation.py), is the sequential fp32 sum
bit for bit -- on PageRank-like contributions, uniform products, rows that start from a non-zero y, values with many
exact half-way addends (multiples of 2^-k), zeros, negatives and a single huge addend."""
import numpy as np
import pytest

from ordered_sum_emulation import ordered_sum, sequential


def _cases():
    rng = np.random.default_rng(7)
    yield "heavy tail", (rng.pareto(1.2, 40000) * 1e-9).astype(np.float32), np.float32(0)
    yield "contrib", np.exp(rng.normal(-20, 1.0, 200000)).astype(np.float32), np.float32(0)
    yield "uniform products", (rng.random(300000) * rng.random(300000)).astype(np.float32), np.float32(0)
    yield "from y", rng.random(20000).astype(np.float32), np.float32(123.456)
    yield "half-way addends", (rng.integers(0, 1 << 12, 25000) / 4096.0).astype(np.float32), np.float32(0)
    yield "coarse grid", (rng.integers(0, 64, 25000) / 8.0).astype(np.float32), np.float32(0.5)
    z = rng.random(10000).astype(np.float32); z[2000:6000] = 0
    yield "zero run", z, np.float32(0)
    s = (rng.random(12000) - 0.3).astype(np.float32)
    yield "signed", s, np.float32(0)
    h = rng.random(9000).astype(np.float32); h[4000] = np.float32(3e7)
    yield "one huge addend", h, np.float32(0)
    yield "short", rng.random(100).astype(np.float32), np.float32(0)


@pytest.mark.parametrize("name,x,start", list(_cases()), ids=[c[0] for c in _cases()])
def test_ordered_sum_is_the_sequential_sum(name, x, start):
    got, fast = ordered_sum(x, start)
    want = sequential(x, start)
    assert got.view(np.uint32) == want.view(np.uint32), (name, float(got), float(want))
    n_blocks = -(-len(x) // 512)
    if name in ("contrib", "uniform products", "from y"):
        assert fast >= (3 * n_blocks) // 4    # the integer path carries most blocks: the method is not a disguised fallback


def test_ties_are_what_breaks_the_integer_path():
    """Without the tie rule the integer path differs from the hardware on half-way addends: the rule is needed."""
    import ordered_sum_emulation as ose
    rng = np.random.default_rng(3)
    x = (rng.integers(0, 1 << 12, 60000) / 4096.0).astype(np.float32)
    want = sequential(x)
    real_qsum = ose.qsum
    try:
        ose.qsum = lambda blk, ex: (real_qsum(blk, ex)[0], False)
        got, _ = ose.ordered_sum(x)
    finally:
        ose.qsum = real_qsum
    assert got.view(np.uint32) != want.view(np.uint32)
    assert ose.ordered_sum(x)[0].view(np.uint32) == want.view(np.uint32)
