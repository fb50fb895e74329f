import os, sys, numpy as np
sys.path.insert(0, '.')
import gardenia_b200 as gb
from oracle import pyoracle as po
scale = int(sys.argv[1])
g = gb.Graph.generate("g", scale, 16)
m = g.m
o, oit, otr = po.pr_pull(m, g.out_rowptr(), g.out_colidx(), g.out_degrees())
print("oracle iters", oit, [float(f"{x:.5g}") for x in otr[:8]], flush=True)
cases = ((1, {}), (2, {}), (2, {"GDN_PR_EXACT_COLS": "2000", "GDN_PR_EXACT_BUDGET": str(1 << 40)}), (2, {"GDN_PR_EXACT_BUDGET": "0"}))
if scale >= 26:
    cases = ((2, {}), (2, {"GDN_PR_EXACT_BUDGET": "0"}), (2, {"GDN_DEVICE_ARENA": "0"}), (2, {"GDN_PR_WARM_MB": "0"}))
for n, env in cases:
    for k in ("GDN_PR_EXACT_COLS", "GDN_PR_EXACT_BUDGET", "GDN_PR_NCCL", "GDN_PR_WARM_MB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    gb.init_gpus(n)
    hs = np.full(m, np.float32(1.0) / np.float32(m), dtype=np.float32)
    try:
        st = gb.PRSolver(g, hs, verbose=False)
        l1 = float(np.abs(hs.astype(np.float64) - o.astype(np.float64)).sum())
        print(f"gpus {n} env {env}: iters {st.iterations} l1 {l1:.3e} trace {[float(f'{x:.5g}') for x in st.pr_trace()[:8]]}", flush=True)
    except Exception as e:
        print(f"gpus {n} env {env}: FAILED {e}", flush=True)
gb.init_gpus(1)
