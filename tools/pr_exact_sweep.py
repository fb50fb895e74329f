#!/usr/bin/env python
"""pr_exact_sweep.py -- L1 distance to the reference's own scores and iteration time of the resident PageRank as a
function of the exact-slice threshold (GDN_PR_EXACT_COLS, csrc/pull.cu pull_prepare) on the cached Kronecker graph."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    settings = sys.argv[2].split(";") if len(sys.argv) > 2 else ["GDN_PR_EXACT_COLS=1000000000", "GDN_PR_EXACT_COLS=65536", "GDN_PR_EXACT_COLS=16384"]
    import numpy as np
    import torch
    import bench
    import gardenia_b200 as gb
    pre, g = bench.load_graph("g", scale)
    ref_path = bench.ref_out(f"ref_scores_g{scale}.f32")
    if not os.path.exists(ref_path):
        bench.cpu_reference_pr(pre, scale, 1, os.cpu_count() or 1)
    ref = np.fromfile(ref_path, dtype=np.float32).astype(np.float64)
    ref_it = json.load(open(ref_path + ".json"))["iterations"]
    m = g.m
    init = float(np.float32(1.0) / np.float32(m))
    for sw in settings:
        env = dict(kv.split("=") for kv in sw.split(",") if kv)
        os.environ.update(env)
        gb._lib.lib.gdn_set_pr_exact_order(int(env.get("GDN_PR_EXACT", "0")))
        dg = gb.DeviceGraph(g)
        s = torch.empty(m, dtype=torch.float32, device="cuda")
        best = None
        for _ in range(3):
            s.fill_(init)
            st = dg.pagerank(s)
            k = st.kernel_ms / max(st.kernel_calls, 1)
            best = k if best is None else min(best, k)
        l1 = float(np.abs(s.cpu().numpy().astype(np.float64) - ref).sum())
        info = dg.pull_info()
        print(json.dumps({"env": sw, "iter_ms": round(best, 3), "solve_ms": round(st.solve_ms, 2), "iterations": [st.iterations, ref_it], "l1_vs_reference": l1,
                          "layout": st.pr_layout, "band_entries": info["band_entries"], "main_groups": info["main_groups"], "prep_ms": dg.prep_ms()}), flush=True)
        dg.close()
        for k in env:
            os.environ.pop(k, None)
        del s


if __name__ == "__main__":
    main()
