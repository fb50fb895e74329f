"""Executable statement of the ORDERED SUM (gardenia_b200/csrc/ordered_core.cuh) in numpy, for the CPU test suite.

The kernels sum a long row in the reference's order (src/pr/omp_base.cc:28-30, src/spmv/omp_base.cc:27-31: one fp32
accumulator, addends in column order) without its chain of dependent adds.  The method, restated here step by step:

  * while the accumulator stays inside one binade [2^k, 2^(k+1)) every add rounds its addend to a multiple of
    ulp = 2^(k-23) and adds it exactly:  acc' = (M + rne(x / ulp)) * ulp,  M = acc / ulp a 24-bit integer -- so a block of
    512 addends contributes the INTEGER Q = sum rne(x_i / ulp), whatever the order;
  * pass 2 (plan) guesses the binade each block will see from the real-valued prefix of the block sums and keeps the guess
    only when the block lies safely inside it; pass 3 (qsum) computes Q for the guessed binade and un-plans a block that
    holds an addend EXACTLY half-way between two multiples of ulp (the hardware rounds such an add to the even
    ACCUMULATOR, which an order-free integer cannot know); pass 4 (combine) walks the blocks in order with the true
    accumulator: a planned block whose guess holds and whose Q does not carry out of the binade is one integer add, every
    other block is added one by one (ordered_block; here: plain sequential fp32 adds, which is what it emulates).

`ordered_sum(x, start)` returns the fp32 result and how many blocks took the integer path; tests/test_ordered_sum_emulation.py
holds it to the sequential sum bit for bit."""
import numpy as np

BLOCK = 512
ONE = 1 << 24


def _bits(f):
    return int(np.float32(f).view(np.uint32))


def _from_bits(b):
    return np.uint32(b).view(np.float32)


def acc_ok(ab):                       # ord_acc_ok: normal, 1/ulp a normal float, far from overflow, non-negative
    return 0x0C000000 <= ab < 0x7F000000


def scale_of(e):                      # ord_scale: 1 / ulp(acc) = 2^(23 - (e - 127)) for biased exponent e
    return _from_bits((277 - e) << 23)


def sequential(x, start=np.float32(0.0)):
    acc = np.float32(start)
    for v in x:
        acc = np.float32(acc + np.float32(v))
    return acc


def plan(blocks, start):
    """pass 2: per block 0 (careful), 1 (all zero) or the guessed biased exponent."""
    out, prefix = [], float(start)
    for b, blk in enumerate(blocks):
        s = float(np.sum(blk.astype(np.float64)))
        mx = int(blk.view(np.uint32).max()) if len(blk) else 0
        p = 0
        if mx == 0:
            p = 1
        pb, qb = _bits(np.float32(prefix)), _bits(np.float32(prefix + s))
        ex = pb >> 23
        if (b > 0 and p == 0 and mx < 0x7F800000 and acc_ok(pb) and (qb >> 23) == ex and (pb & 0x7FFFFF) > 0x4000 and
                (qb & 0x7FFFFF) < 0x7FC000 and float(np.float32(_from_bits(mx) * scale_of(ex))) < 16384.0):
            p = ex
        out.append(p)
        prefix += s
    return out


def qsum(blk, ex):
    """pass 3: (Q, tie) for the guessed binade."""
    t = (blk * scale_of(ex)).astype(np.float32)           # exact: a power of two
    q = np.rint(t.astype(np.float64)).astype(np.int64)     # rne, like cvt.rni
    tie = bool(np.any((t < 8388608.0) & ((t - np.floor(t)) == 0.5)))
    return int(q.sum()), tie


def ordered_sum(x, start=np.float32(0.0)):
    x = np.asarray(x, dtype=np.float32)
    pad = (-len(x)) % BLOCK
    if pad:
        x = np.concatenate([x, np.zeros(pad, dtype=np.float32)])     # padding adds +0.0f
    blocks = x.reshape(-1, BLOCK)
    plans = plan(blocks, start)
    ab = _bits(start)
    fast = 0
    for blk, p in zip(blocks, plans):
        if p == 1:
            continue
        if p > 1:
            q, tie = qsum(blk, p)
            m = (ab & 0x7FFFFF) | 0x800000
            if not tie and p == (ab >> 23) and q < ONE and m + q < ONE:
                ab = (ab & 0xFF800000) | ((m + q) & 0x7FFFFF)
                fast += 1
                continue
        ab = _bits(sequential(blk, _from_bits(ab)))
    return _from_bits(ab), fast
