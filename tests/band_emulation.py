"""Executable statement of the arithmetic of the banded / segmented PageRank layouts (gardenia_b200/csrc/band.cu,
pull.cu) in plain numpy + Python loops, for small graphs.  Test infrastructure: it checks the METHOD against the
oracle on the CPU (tests/test_band_emulation.py) and documents, in ~100 lines, what the kernels compute:

  * vertices renumbered by degree, hottest first; sorted row j sums contrib[new id] over its columns;
  * band of a new id = band_of (hot prefix pieces, then cold-slice pieces; one GPU here);
  * a (row, band) pair with >= cmin ids of a row of length >= dmin is summed separately: sequential fp32 adds in column
    order, restarted every `seg_ids` ids (one work item), each partial converted to a 2^-56 fixed-point integer and
    added EXACTLY into the row's accumulator;
  * the remaining ids of the row are summed sequentially in fp32 in column order (the main array);
  * row value = float32(float64(main sum) + accumulator * 2^-56); then the reference's epilogue
    (src/pr/omp_base.cc:24-36): score = base + damp * sum, fp64 L1 delta, contrib = score / out_degree.
"""
import numpy as np

FIX = 2.0 ** 56


def band_of(c, H, band, n_bands):
    """one GPU: ids [0, H) in pieces of `band`, then the cold slice in pieces of `band` (csrc/band.cu band_of)."""
    if c < H:
        b = c // band
    else:
        b = -(-H // band) + (c - H) // band
    return b if b < n_bands else -1


def pagerank_banded(m, rowptr, colidx, n_bands=64, band=49152, cmin=4, dmin=64, seg_ids=512, hot=49152,
                    damp=np.float32(0.85), eps=1e-4, max_iter=100):
    rowptr = np.asarray(rowptr, dtype=np.int64)
    deg = np.diff(rowptr)
    perm = np.argsort(-deg, kind="stable")                 # sorted position -> old row  (degree desc, id asc)
    newid = np.empty(m, dtype=np.int64)
    newid[perm] = np.arange(m)
    H = min(hot, m)
    # split every sorted row into its band sections and its main remainder (column order kept)
    rows = []
    for j in range(m):
        r = perm[j]
        cols = newid[colidx[rowptr[r]:rowptr[r + 1]]]
        sections, main = {}, []
        if len(cols) >= dmin:
            bands = [band_of(int(c), H, band, n_bands) for c in cols]
            cnt = {}
            for b in bands:
                if b >= 0:
                    cnt[b] = cnt.get(b, 0) + 1
            for c, b in zip(cols, bands):
                if b >= 0 and cnt[b] >= cmin:
                    sections.setdefault(b, []).append(int(c))
                else:
                    main.append(int(c))
        else:
            main = [int(c) for c in cols]
        rows.append((sections, main))
    f32 = np.float32
    scores = np.full(m, f32(1.0) / f32(m), dtype=np.float32)       # sorted order; src/pr/main.cc:17-18
    base = (f32(1.0) - damp) / f32(m)
    sdeg = deg[perm].astype(np.float32)
    moved = sum(len(v) for s, _ in rows for v in s.values())
    with np.errstate(divide="ignore", invalid="ignore"):
        contrib = (scores / sdeg).astype(np.float32)
    it = 0
    for it in range(1, max_iter + 1):
        err = 0.0
        new_scores = np.empty_like(scores)
        for j, (sections, main) in enumerate(rows):
            acc = f32(0.0)
            for c in main:
                acc = f32(acc + contrib[c])
            fix = 0
            for b in sorted(sections):
                ids = sections[b]
                for k0 in range(0, len(ids), seg_ids):
                    part = f32(0.0)
                    for c in ids[k0:k0 + seg_ids]:
                        part = f32(part + contrib[c])
                    fix += int(np.rint(np.float64(part) * FIX))
            tot = f32(np.float64(acc) + fix / FIX) if sections else acc
            nw = f32(base + f32(damp * tot))
            err += abs(float(f32(nw - scores[j])))
            new_scores[j] = nw
        scores = new_scores
        with np.errstate(divide="ignore", invalid="ignore"):
            contrib = (scores / sdeg).astype(np.float32)
        if err < eps:
            break
    out = np.empty(m, dtype=np.float32)
    out[perm] = scores
    return out, it, moved
