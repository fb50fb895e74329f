"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun), 1-D row
partition, NCCL exchange inside libgdn_b200; rank 0 checks against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gardenia_b200 as gb  # noqa: E402
from gardenia_b200 import _lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib.gdn_init(local))
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _lib.check(_lib.lib.gdn_comm_unique_id(uid.ctypes.data))
    t = torch.from_numpy(uid).to(dev)
    dist.broadcast(t, 0)
    uid = t.cpu().numpy()
    _lib.check(_lib.lib.gdn_comm_init(rank, world, uid.ctypes.data))
    ok = True
    # third case: many small bands, so that the banded layout (csrc/band.cu) spreads over every rank's cold slice
    small = dict(GDN_PR_BANDS="96", GDN_PR_BAND_SIZE="512", GDN_PR_BAND_CMIN="2", GDN_PR_BAND_DMIN="8")
    seg = dict(GDN_PR_SEGMENT="1", GDN_PR_SEG_IDS="3000")           # segmented mode forced on a small uniform graph
    nccl = dict(GDN_PR_NCCL="1")                                    # the NCCL collectives instead of the peer-mapped exchange
    for kind, scale, env in (("g", 16, {}), ("u", 15, {}), ("g", 17, small), ("u", 16, seg), ("g", 16, nccl)):
        for k in list(small) + list(seg) + list(nccl):
            os.environ.pop(k, None)
        os.environ.update(env)
        g = gb.Graph.generate(kind, scale, 16)
        m = g.m
        b = gb.partition_rows(m, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        w = int(b[1])
        dg = gb.DeviceGraph(g, lo, hi, device=local)
        # ---- PageRank
        scores = torch.full((hi - lo,), float(np.float32(1.0) / np.float32(m)), dtype=torch.float32, device=dev)
        st = dg.pagerank(scores)
        pinfo = dg.pull_info()
        if env is seg:
            assert pinfo["banded"] == 2, pinfo
        elif kind == "g":
            assert pinfo["banded"] == 1 and pinfo["band_entries"] > 0, pinfo       # no silent fallback to the plain layout
        full = torch.zeros(w * world, dtype=torch.float32, device=dev)
        full[lo:hi] = scores
        dist.all_gather_into_tensor(full, full[rank * w:(rank + 1) * w].clone())
        # ---- BFS
        srcs = [int(s) for s in g.pick_sources(3)] + [0]
        depths = []
        for s in srcs:
            depth = torch.empty(m, dtype=torch.int32, device=dev)
            parent = torch.empty(m, dtype=torch.int32, device=dev)
            stb = dg.bfs(s, depth, parent)
            depths.append((depth.cpu().numpy(), stb.iterations, parent.cpu().numpy()))
        if rank == 0:
            from oracle import pyoracle as po
            rp, ci = g.out_rowptr(), g.out_colidx()
            oscores, oit, _ = po.pr_pull(m, rp, ci, g.out_degrees())
            l1 = float(np.abs(full[:m].cpu().numpy().astype(np.float64) - oscores.astype(np.float64)).sum())
            good = (st.iterations == oit) and l1 <= 1e-6
            print(f"[multi] {kind}{scale} world={world} PR iters {st.iterations} vs {oit} l1={l1:.3e} {'OK' if good else 'FAIL'} "
                  f"bands={pinfo['bands']}x{pinfo['band_ids']} band_entries={pinfo['band_entries']}", flush=True)
            ok &= good
            for s, (d, it, par) in zip(srcs, depths):
                od, oit, _ = po.bfs_do(m, rp, ci, rp, ci, s)
                good = np.array_equal(d, od) and it == oit and po.bfs_check_parents(m, rp, ci, s, od, par) == 0
                print(f"[multi] {kind}{scale} BFS src {s}: iters {it} vs {oit}, parents {'OK' if good else 'FAIL'}", flush=True)
                ok &= good
        dg.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    _lib.lib.gdn_comm_destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
