#!/usr/bin/env python
"""prof_run.py -- run ONE solver of the hot path a few times on a cached synthetic
graph, for `ncu` captures and quick timing (not a benchmark: bench.py is).

    python tools/prof_run.py pr   --kind g --scale 26 --reps 2
    python tools/prof_run.py bfs  --kind g --scale 26 --reps 4
    python tools/prof_run.py spmv --kind u --scale 24 --reps 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["pr", "bfs", "spmv"])
    ap.add_argument("--kind", default="g")
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--sweep", default="", help="';'-separated env settings, e.g. 'GDN_PR_WARM_MB=32;GDN_PR_WARM_MB=64,GDN_PR_BANDS=96'")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import gardenia_b200 as gb

    dev = torch.device("cuda", 0)
    _, g = bench.load_graph(args.kind, args.scale)
    dg = gb.DeviceGraph(g, device=0)
    out = {"what": args.what, "kind": args.kind, "scale": args.scale, "m": g.m, "nnz": g.nnz, "runs": []}
    sweeps = [x for x in args.sweep.split(";") if x] or [""]
    if args.what == "pr":
        scores = torch.empty(g.m, dtype=torch.float32, device=dev)
        for sw in sweeps:
            for kv in [x for x in sw.split(",") if x]:
                k, v = kv.split("=")
                os.environ[k] = v
            for _ in range(args.reps):
                scores.fill_(float(np.float32(1.0) / np.float32(g.m)))
                st = dg.pagerank(scores)
                out["runs"].append({"env": sw, "solve_ms": st.solve_ms, "iterations": st.iterations, "kernel_ms": st.kernel_ms,
                                    "kernel_calls": st.kernel_calls, "launches": st.kernel_launches,
                                    "checksum": float(scores.double().sum().item())})
    elif args.what == "bfs":
        depth = torch.empty(g.m, dtype=torch.int32, device=dev)
        for s in [int(x) for x in g.pick_sources(args.reps)]:
            st = dg.bfs(s, depth)
            steps = [(st.steps[i].dir, st.steps[i].frontier, st.steps[i].discovered, st.steps[i].scout, st.steps[i].edges, st.steps[i].scanned) for i in range(st.n_steps)]
            out["runs"].append({"source": s, "solve_ms": st.solve_ms, "iterations": st.iterations, "kernel_ms": st.kernel_ms,
                                "launches": st.kernel_launches, "edges_reached": st.edges_reached,
                                "gteps": st.edges_reached / 2 / st.solve_ms / 1e6, "steps": steps})
    else:
        Ax = torch.from_numpy(gb.fill_uniform(13, g.nnz)).to(dev)
        x = torch.from_numpy(gb.fill_uniform(14, g.m)).to(dev)
        for sw in sweeps:
            for kv in [x for x in sw.split(",") if x]:
                k, v = kv.split("=")
                os.environ[k] = v
            for _ in range(args.reps):
                y = torch.zeros(g.m, dtype=torch.float32, device=dev)
                st = dg.spmv(Ax, x, y)
                out["runs"].append({"env": sw, "solve_ms": st.solve_ms, "kernel_ms": st.kernel_ms,
                                    "checksum": float(y.double().sum().item())})
    dg.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
