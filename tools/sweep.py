#!/usr/bin/env python
"""sweep.py -- the throughput table BASELINE.json asks for: BFS GTEPS, PageRank iterations/s and SpMV
GFLOP/s on synthetic Kronecker (-g) and uniform-random (-u) graphs at scales 22-27, one B200, each with
its fraction of the measured HBM roofline (algorithmic bytes of SURVEY 8(d) / kernel time / hbm_gbs).

    python tools/sweep.py --scales 22,23,24,25,26,27 --kinds g,u > profiles/rN_sweep.jsonl

One JSON line per (kind, scale); stderr carries progress.  Not the benchmark of record (bench.py is).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", default="22,23,24")
    ap.add_argument("--kinds", default="g,u")
    ap.add_argument("--sources", type=int, default=8)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import gardenia_b200 as gb

    peak, _ = bench.hbm_peak()
    dev = torch.device("cuda", 0)
    for kind in args.kinds.split(","):
        for scale in [int(x) for x in args.scales.split(",")]:
            t0 = time.time()
            g = gb.Graph.generate(kind, scale, 16)
            m, nnz = g.m, g.nnz
            bench.log(f"[sweep] {kind}{scale}: m={m} nnz={nnz} generated in {time.time() - t0:.1f}s")
            dg = gb.DeviceGraph(g, device=0)
            row = {"kind": "kron" if kind == "g" else "urand", "scale": scale, "m": m, "nnz": nnz}
            # ---- PageRank
            scores = torch.empty(m, dtype=torch.float32, device=dev)
            init = float(np.float32(1.0) / np.float32(m))
            best = None
            for _ in range(3):
                scores.fill_(init)
                st = dg.pagerank(scores)
                if best is None or st.solve_ms < best[0]:
                    best = (st.solve_ms, st.iterations, st.kernel_ms / max(st.kernel_calls, 1))
            alg = 4 * nnz + 20 * m + 4
            row["pr"] = {"iterations": best[1], "solve_ms": best[0], "iters_per_s": best[1] / (best[0] / 1e3),
                         "gather_ms": best[2], "roofline_frac_kernel": alg / (best[2] / 1e3) / 1e9 / peak,
                         "roofline_frac_solve": alg * best[1] / (best[0] / 1e3) / 1e9 / peak}
            del scores
            # ---- BFS
            depth = torch.empty(m, dtype=torch.int32, device=dev)
            srcs = [int(s) for s in g.pick_sources(args.sources)]
            dg.bfs(srcs[0], depth)
            ms = edges = algb = 0.0
            per = []
            for s in srcs:
                st = dg.bfs(s, depth)
                ms += st.solve_ms; edges += st.edges_reached / 2
                algb += bench.bfs_algorithmic_bytes(m, st.bfs_steps())
                per.append(st.edges_reached / 2 / (st.solve_ms / 1e3) / 1e9)
            row["bfs"] = {"gteps": edges / (ms / 1e3) / 1e9, "median_gteps": float(np.median(per)), "ms_per_bfs": ms / len(srcs),
                          "roofline_frac": algb / (ms / 1e3) / 1e9 / peak}
            del depth
            # ---- SpMV
            Ax = torch.from_numpy(gb.fill_uniform(13, nnz)).to(dev)
            x = torch.from_numpy(gb.fill_uniform(14, m)).to(dev)
            y = torch.zeros(m, dtype=torch.float32, device=dev)
            for _ in range(2):
                dg.spmv(Ax, x, y)
            k = min(dg.spmv(Ax, x, y).kernel_ms for _ in range(5))
            algs = 8 * nnz + 16 * m + 4
            row["spmv"] = {"ms": k, "gflops": 2.0 * nnz / (k / 1e3) / 1e9, "roofline_frac": algs / (k / 1e3) / 1e9 / peak}
            del Ax, x, y
            dg.close()
            del g
            torch.cuda.empty_cache()
            print(json.dumps(row), flush=True)
            bench.log(f"[sweep] {kind}{scale} done in {time.time() - t0:.1f}s: PR {row['pr']['iters_per_s']:.1f} it/s, "
                      f"BFS {row['bfs']['gteps']:.0f} GTEPS, SpMV {row['spmv']['gflops']:.0f} GFLOP/s")


if __name__ == "__main__":
    main()
