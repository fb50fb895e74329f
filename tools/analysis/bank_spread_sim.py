#!/usr/bin/env python
"""bank_spread_sim.py -- the re-ordering that csrc/band.cu `band_spread` applies to the ids of an item (32 rows = 32 lanes,
each reading one of ITS ids per step from a 49 152-word shared-memory table: bank = id mod 32), simulated on random ids.
At step p lane l prefers bank (l + p) mod 32 (a Latin square: conflict-free when every lane finds its preferred bank),
takes the nearest bank it still has an id in, and lanes that collide are settled lowest-lane-first over four rounds, the
losers looking for a bank nobody holds.  Printed: mean over steps of the most-loaded bank (= shared-memory wavefronts per
table read) before / after, for rows of L ids.  Measured on the real band arrays of Kronecker scale 26: 3.23 -> 2.15
(GDN_TRACE, band_build), which took pr_band_kernel from 1.04 to 0.87 ms once its index stream was prefetched
(profiles/r2_band_kernel_ab.txt).  The optimum (an edge colouring of the lane x bank multigraph) is ~1.3-1.75 here; a
load-aware greedy was simulated too and is WORSE than the Latin square (every lane rushes to the same loaded bank)."""
import random, sys
def nearest(m, pref):
    r = ((m >> pref) | (m << (32 - pref))) & 0xffffffff
    f = (r & -r).bit_length() - 1
    return (f + pref) & 31
def degree(banks):
    c = {}
    for b in banks:
        if b is not None: c[b] = c.get(b, 0) + 1
    return max(c.values()) if c else 0
def sim(rows):  # rows: 32 lists of ids (<=64)
    np_ = 8 * ((max(len(r) for r in rows) + 7) // 8)
    before = sum(degree([r[p] & 31 if p < len(r) else None for r in rows]) for p in range(np_))
    lists = [[[] for _ in range(32)] for _ in range(32)]
    for l, r in enumerate(rows):
        for i in r: lists[l][i & 31].append(i)
    avail = [sum(1 << b for b in range(32) if lists[l][b]) for l in range(32)]
    left = [len(r) for r in rows]
    out = [[] for _ in range(32)]
    after = 0
    for p in range(np_):
        real = [left[l] > 0 for l in range(32)]
        want = [nearest(avail[l], (l + p) & 31) if real[l] else None for l in range(32)]
        fixed = [not real[l] for l in range(32)]
        for rnd in range(4):
            fixed_banks = {want[l] for l in range(32) if fixed[l] and real[l]}
            groups = {}
            for l in range(32):
                if real[l]: groups.setdefault(want[l], []).append(l)
            for b, ls in groups.items():
                if b not in fixed_banks and not fixed[ls[0]]:
                    # lowest lane of the group wins (ls sorted)
                    fixed[ls[0]] = True
            held = 0
            for l in range(32):
                if fixed[l] and real[l]: held |= 1 << want[l]
            if all(fixed): break
            for l in range(32):
                if not fixed[l]:
                    alt = avail[l] & ~held
                    if alt: want[l] = nearest(alt, (l + p) & 31)
        after += degree(want)
        for l in range(32):
            if real[l]:
                b = want[l]
                out[l].append(lists[l][b].pop(0))
                if not lists[l][b]: avail[l] &= ~(1 << b)
                left[l] -= 1
    for l in range(32):
        assert sorted(out[l]) == sorted(rows[l])
    return before, after, np_
random.seed(1)
for L in (8, 16, 32, 64):
    tb = ta = tp = 0
    for _ in range(40):
        rows = [sorted(random.sample(range(49152), max(1, L - random.randint(0, L // 8)))) for _ in range(32)]
        b, a, n = sim(rows); tb += b; ta += a; tp += n
    print(L, "before %.2f after %.2f" % (tb / tp, ta / tp))
