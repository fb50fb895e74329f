#!/usr/bin/env python
"""sweep.py -- the throughput table BASELINE.json asks for: BFS GTEPS, PageRank iterations/s and SpMV
GFLOP/s on synthetic Kronecker (-g) and uniform-random (-u) graphs at scales 22-27, one B200, each with
its fraction of the measured HBM roofline (algorithmic bytes of SURVEY 8(d) / kernel time / hbm_gbs), the
reference's OpenMP kernels (oracle/_ref/ref_driver, this box's host cores) beside them, and the parity of
the two on that very graph (PageRank L1 + iteration count, BFS depths, SpMV per-row relative error).

    python tools/sweep.py --scales 22,23,24,25,26,27 --kinds g,u [--no-cpu] > profiles/rN_sweep.jsonl

One JSON line per (kind, scale); stderr carries progress.  Not the benchmark of record (bench.py is).
"""
import argparse
import glob
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
    ap.add_argument("--cpu-sources", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="pr,bfs,spmv", help="which of pr,bfs,spmv to run")
    ap.add_argument("--keep", action="store_true", help="leave the cached graph in tmpfs")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import gardenia_b200 as gb

    peak, _ = bench.hbm_peak()
    dev = torch.device("cuda", 0)
    ncpu = os.cpu_count() or 1
    gb._lib.lib.gdn_set_host_threads(ncpu)
    cpu = (not args.no_cpu) and os.path.exists(bench.REF_DRIVER)
    only = set(args.only.split(","))
    for kind in args.kinds.split(","):
        for scale in [int(x) for x in args.scales.split(",")]:
            t0 = time.time()
            pre, g = bench.load_graph(kind, scale)             # generated (CSR built on the GPU) + cached for the reference
            m, nnz = g.m, g.nnz
            dg = gb.DeviceGraph(g, device=0)
            row = {"kind": "kron" if kind == "g" else "urand", "scale": scale, "m": m, "nnz": nnz, "prep_ms": dg.prep_ms()}
            row["pr"], row["bfs"], row["spmv"] = {}, {}, {}
            pr_mine, bfs_mine, y_mine, best, srcs, edges = None, [], None, None, [], 0.0
            # ---- PageRank
            if "pr" in only:
                scores = torch.empty(m, dtype=torch.float32, device=dev)
                init = float(np.float32(1.0) / np.float32(m))
                for _ in range(3):
                    scores.fill_(init)
                    st = dg.pagerank(scores)
                    if best is None or st.solve_ms < best[0]:
                        best = (st.solve_ms, st.iterations, st.kernel_ms / max(st.kernel_calls, 1))
                alg = 4 * nnz + 20 * m + 4
                row["pr"] = {"iterations": best[1], "solve_ms": best[0], "iters_per_s": best[1] / (best[0] / 1e3),
                             "ms_per_iter": best[0] / best[1], "roofline_frac": alg * best[1] / (best[0] / 1e3) / 1e9 / peak}
                pr_mine = scores.cpu().numpy()
                del scores
            # ---- BFS
            if "bfs" in only:
                depth = torch.empty(m, dtype=torch.int32, device=dev)
                srcs = [int(s) for s in g.pick_sources(args.sources)]
                dg.bfs(srcs[0], depth)
                ms = algb = 0.0
                per = []
                for i, s in enumerate(srcs):
                    st = dg.bfs(s, depth)
                    ms += st.solve_ms; edges += st.edges_reached / 2
                    algb += bench.bfs_algorithmic_bytes(m, st.bfs_steps())
                    per.append(st.edges_reached / 2 / (st.solve_ms / 1e3) / 1e9)
                    if i < args.cpu_sources:
                        d = depth.cpu().numpy()
                        bfs_mine.append((np.where(d == gb.GDN_INFINITY, -1, d).astype(np.int8), st.iterations))   # ref_driver: -1 = unreached
                row["bfs"] = {"gteps": edges / (ms / 1e3) / 1e9, "median_gteps": float(np.median(per)), "ms_per_bfs": ms / len(srcs),
                              "roofline_frac": algb / (ms / 1e3) / 1e9 / peak}
                del depth
            # ---- SpMV (every launch of the call: the scatter of x on the hot-first path, the long rows' ordered sum)
            if "spmv" in only:
                both = gb.fill_uniform(13, nnz + m)                # the stream oracle/ref_driver.cc draws Ax, then x from
                Ax = torch.from_numpy(both[:nnz]).to(dev)
                x = torch.from_numpy(both[nnz:]).to(dev)
                del both
                y = torch.zeros(m, dtype=torch.float32, device=dev)
                for _ in range(2):
                    dg.spmv(Ax, x, y)
                runs = [dg.spmv(Ax, x, y) for _ in range(5)]
                k = min(r.solve_ms for r in runs)
                algs = 8 * nnz + 16 * m + 4
                row["spmv"] = {"ms": k, "main_kernel_ms": min(r.kernel_ms for r in runs), "launches": int(runs[0].kernel_launches),
                               "gflops": 2.0 * nnz / (k / 1e3) / 1e9, "roofline_frac": algs / (k / 1e3) / 1e9 / peak}
                y.zero_()
                dg.spmv(Ax, x, y)
                y_mine = y.cpu().numpy()
                del Ax, x, y
            dg.close()
            del g
            torch.cuda.empty_cache()
            # ---- the reference's OpenMP kernels on the same graph, and parity against their outputs
            if cpu:
                tag = os.path.join(bench.CACHE_DIR, f"sweep_{kind}{scale}")
                try:
                    if "pr" in only:
                        msr, its, _ = bench.run_reference_binary([bench.REF_DRIVER, "pr", "bin", pre, "1", tag + ".pr", "1"], ncpu)
                        ref = np.fromfile(tag + ".pr", dtype=np.float32)
                        l1 = float(np.abs(pr_mine.astype(np.float64) - ref.astype(np.float64)).sum())
                        row["pr"].update(cpu_iters_per_s=its[-1] / (msr[-1] / 1e3), cpu_iterations=its[-1], l1_vs_reference=l1,
                                         parity_ok=bool(l1 <= 1e-6 and its[-1] == best[1]))
                    if "bfs" in only:
                        cs = srcs[:args.cpu_sources]
                        msr, its, _ = bench.run_reference_binary([bench.REF_DRIVER, "bfs", "bin", pre, "1", "0",
                                                                  ",".join(str(s) for s in cs) + ",", tag + ".bfs"], ncpu)
                        ref = np.fromfile(tag + ".bfs", dtype=np.int8).reshape(len(cs), m)
                        same = all(np.array_equal(ref[i], bfs_mine[i][0]) for i in range(len(cs)))
                        row["bfs"].update(cpu_ms_per_bfs=sum(msr) / len(msr), cpu_gteps=(edges / len(srcs)) / (sum(msr) / len(msr) / 1e3) / 1e9,
                                          depths_identical=bool(same), iterations_identical=bool([b[1] for b in bfs_mine] == its[:len(cs)]),
                                          parity_ok=bool(same))
                    if "spmv" in only:
                        msr, _, _ = bench.run_reference_binary([bench.REF_DRIVER, "spmv", "bin", pre, "1", "0", "13", tag + ".spmv", "3"], ncpu)
                        ref = np.fromfile(tag + ".spmv", dtype=np.float32)
                        rel = float((np.abs(y_mine - ref) / np.maximum(np.abs(ref), 1e-30)).max())
                        row["spmv"].update(cpu_gflops=2.0 * nnz / (min(msr) / 1e3) / 1e9, maxrel_vs_reference=rel, parity_ok=bool(rel <= 1e-5),
                                           rows_bit_identical=float((y_mine == ref).mean()))
                    row["cpu_cores"] = ncpu
                except Exception as e:  # noqa: BLE001
                    row["cpu_error"] = str(e)[-300:]
                for f in glob.glob(tag + ".*"):
                    os.remove(f)
            if not args.keep:
                for f in glob.glob(pre + ".*"):                # the cached triple of this graph (tmpfs) is not needed again
                    os.remove(f)
            print(json.dumps(row), flush=True)
            bench.log(f"[sweep] {kind}{scale} done in {time.time() - t0:.1f}s: PR {row['pr'].get('iters_per_s', 0):.1f} it/s, "
                      f"BFS {row['bfs'].get('gteps', 0):.0f} GTEPS, SpMV {row['spmv'].get('gflops', 0):.0f} GFLOP/s")


if __name__ == "__main__":
    main()
