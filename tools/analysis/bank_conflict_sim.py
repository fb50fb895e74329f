#!/usr/bin/env python
"""bank_conflict_sim.py -- how many shared-memory wavefronts a warp-wide LDS of pr_band_kernel needs (32 lanes = 32 rows,
each reading one of ITS ids of the band: bank = id mod 32), and how far a smarter order of a row's ids inside a band
slice could cut that.  Pure simulation on random residues; n = ids per row in the slice.

  baseline  ids in encounter order                                    ~3.5 wavefronts (ncu on Kron-26: 3.3)
  rotated   each lane sorts its ids by (residue - lane) mod 32         3.2 at n=32  (noise of +-sqrt(n) positions kills it)
  greedy    per step, lanes with the fewest choices pick first an id   2.1 at n=32, 1.7 at n=64, 1.4 at n=128
            whose residue is least used in this step
"""
import numpy as np
rng = np.random.default_rng(1)


def baseline(n, trials):
    r = []
    for _ in range(trials):
        ids = rng.integers(0, 32, size=(32, n))
        r.append(np.mean([np.bincount(ids[:, k], minlength=32).max() for k in range(n)]))
    return float(np.mean(r))


def rotated(n, trials):
    r = []
    for _ in range(trials):
        sched = np.empty((32, n), int)
        for l in range(32):
            x = rng.integers(0, 32, size=n)
            sched[l] = x[np.argsort((x - l) % 32, kind="stable")]
        r.append(np.mean([np.bincount(sched[:, k], minlength=32).max() for k in range(n)]))
    return float(np.mean(r))


def greedy(n, trials):
    r = []
    for _ in range(trials):
        ids = rng.integers(0, 32, size=(32, n))
        cnt = np.stack([np.bincount(ids[l], minlength=32) for l in range(32)])
        tot = 0
        for _k in range(n):
            used = np.zeros(32, int)
            for l in np.argsort([(cnt[l] > 0).sum() for l in range(32)]):
                avail = np.nonzero(cnt[l])[0]
                best = min(avail, key=lambda q: (used[q], -cnt[l][q]))
                used[best] += 1
                cnt[l][best] -= 1
            tot += used.max()
        r.append(tot / n)
    return float(np.mean(r))


if __name__ == "__main__":
    print(f"{'n':>5s} {'baseline':>9s} {'rotated':>8s} {'greedy':>7s}")
    for n in (4, 8, 16, 32, 64, 128):
        t = 40 if n <= 32 else 12
        print(f"{n:5d} {baseline(n, t):9.2f} {rotated(n, t):8.2f} {greedy(n, t):7.2f}")
