"""N>1 host-side logic on CPU: world_size-2 gloo run of the 1-D row partition
and the per-iteration exchange protocol of SURVEY §8(e) (equal-width 64-aligned
slices, in-place allgather of the contrib slice / frontier-bitmap slice,
allreduce of the controller scalars).  The per-rank compute is the oracle
restricted to the rank's rows; the result must equal the single-rank oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gardenia_b200 as gb
    g = gb.Graph.generate("g", 10, 16)
    m, rp, ci = g.m, g.out_rowptr().astype(np.int64), g.out_colidx()
    bounds = gb.partition_rows(m, world)
    w = int(bounds[1])
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    assert w % 64 == 0 and w * world >= m
    deg = np.diff(rp).astype(np.int32)

    # ---- PageRank: local rows + allgather(contrib slice) + allreduce(err), src/pr/omp_base.cc:21-37
    damp, eps = np.float32(0.85), 1e-4
    base = (np.float32(1.0) - damp) / np.float32(m)
    scores = np.full(hi - lo, np.float32(1.0) / np.float32(m), dtype=np.float32)
    contrib = torch.zeros(w * world, dtype=torch.float32)          # padded full-length vector
    iters = 0
    with np.errstate(divide="ignore"):
        for it in range(100):
            contrib[lo:hi] = torch.from_numpy(scores / deg[lo:hi].astype(np.float32))
            dist.all_gather_into_tensor(contrib, contrib[rank * w:(rank + 1) * w].clone())
            c = contrib.numpy()
            err = 0.0
            for r in range(lo, hi):
                tot = np.float32(0)
                for e in range(rp[r], rp[r + 1]):
                    tot = np.float32(tot + c[ci[e]])
                new = np.float32(base + np.float32(damp * tot))
                err += float(abs(np.float32(new - scores[r - lo])))
                scores[r - lo] = new
            t = torch.tensor([err], dtype=torch.float64)
            dist.all_reduce(t)
            iters = it + 1
            if t.item() < eps:
                break
    full = torch.zeros(w * world, dtype=torch.float32)
    full[lo:hi] = torch.from_numpy(scores)
    dist.all_gather_into_tensor(full, full[rank * w:(rank + 1) * w].clone())

    # ---- BFS: bottom-up sweep of own rows against the global frontier bitmap, allgather of the
    # `next` slice (bit-packed, 64-aligned so slices are whole words), allreduce of awake_count
    src = int(g.pick_sources(1)[0])
    depth = np.full(m, -1, dtype=np.int32)
    depth[src] = 0
    front = np.zeros(w * world, dtype=bool)
    front[src] = True
    level = 0
    while True:
        nxt = np.zeros(w * world, dtype=bool)
        for v in range(lo, hi):
            if depth[v] < 0:
                for e in range(rp[v], rp[v + 1]):
                    if front[ci[e]]:
                        depth[v] = level + 1
                        nxt[v] = True
                        break
        packed = torch.from_numpy(np.packbits(nxt, bitorder="little"))
        wb = w // 8
        dist.all_gather_into_tensor(packed, packed[rank * wb:(rank + 1) * wb].clone())
        front = np.unpackbits(packed.numpy(), bitorder="little").astype(bool)
        awake = torch.tensor([int(nxt[lo:hi].sum())], dtype=torch.int64)
        dist.all_reduce(awake)
        # every rank learns the depths of the new frontier from the bitmap
        depth[front[:m] & (depth < 0)] = level + 1
        level += 1
        if awake.item() == 0:
            break
    if rank == 0:
        np.save(os.path.join(out_dir, "scores.npy"), full[:m].numpy())
        np.save(os.path.join(out_dir, "depth.npy"), depth)
        np.save(os.path.join(out_dir, "meta.npy"), np.array([iters, src]))
    dist.destroy_process_group()


def test_partition_bounds():
    import gardenia_b200 as gb
    for m, p in [(1024, 2), (4194302, 8), (67108864, 4), (100, 8), (65, 2)]:
        b = gb.partition_rows(m, p)
        assert b[0] == 0 and b[-1] == m and np.all(np.diff(b) >= 0)
        w = b[1] if p > 1 and b[1] < m else None
        if w:
            assert w % 64 == 0
        assert all((b[i + 1] - b[i]) in (b[1] - b[0], m - b[i], 0) for i in range(p))


def test_two_rank_exchange_matches_oracle(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    import gardenia_b200 as gb
    from oracle import pyoracle as po
    g = gb.Graph.generate("g", 10, 16)
    rp, ci = g.out_rowptr(), g.out_colidx()
    iters, src = np.load(tmp_path / "meta.npy")
    oscores, oit, _ = po.pr_pull(g.m, rp, ci, g.out_degrees())
    assert iters == oit
    assert np.array_equal(np.load(tmp_path / "scores.npy"), oscores)     # same fp32 order per row -> bit-equal
    odist, _, _ = po.bfs_do(g.m, rp, ci, rp, ci, int(src))
    d = np.load(tmp_path / "depth.npy")
    d = np.where(d < 0, po.INFINITY, d)
    assert np.array_equal(d, odist)
