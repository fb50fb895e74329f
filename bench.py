#!/usr/bin/env python
"""bench.py -- the reference's headline benchmark on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
PageRank pull to convergence (damp .85, eps 1e-4, max 100 iterations) on the
synthetic Kronecker graph `-g 26 -k 16` (m = 67.1 M, nnz ~ 2.1 G directed
entries) produced by the reference generator's streams; 1-D row partition at
N > 1 with one NCCL allgather of the contrib slice per iteration.  One "step" =
one full PRSolver solve.  metric = PageRank iterations per second (whole job).

At N = 1 the same line also carries (key "also") BFS GTEPS on the same graph
(16 GAP-style sources) and SpMV GFLOP/s on urand-24, each with its own roofline.

Timing: CUDA events on the library's stream around the solve region of every
step (graph resident in HBM; the reference's timed region, src/pr/base.cu:109-128),
summed over K steps, max over ranks; barrier + synchronize on both sides.
Inputs (8.4 GB of column indices) are far larger than the 126 MB L2, so no
explicit flush is needed between iterations.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CACHE_DIR = os.environ.get("GDN_BENCH_CACHE", "/dev/shm/gdn_bench" if os.path.isdir("/dev/shm") else "/tmp/gdn_bench")
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def graph_prefix(kind, scale, degree=16):
    return os.path.join(CACHE_DIR, f"{'kron' if kind == 'g' else 'urand'}{scale}_k{degree}")


def ensure_graph(kind, scale, degree=16):
    """Generate (reference streams) once per box and cache the reference's binary triple
    (include/csr_graph.h:218-233) in tmpfs; returns (prefix, Graph-or-None)."""
    import gardenia_b200 as gb
    pre = graph_prefix(kind, scale, degree)
    if os.path.exists(pre + ".done"):
        return pre, None
    os.makedirs(CACHE_DIR, exist_ok=True)
    t = time.time()
    g = gb.Graph.generate(kind, scale, degree)
    log(f"[bench] generated {kind}{scale}: m={g.m} nnz={g.nnz} in {time.time() - t:.1f}s")
    t = time.time()
    g.write_bin(pre)
    open(pre + ".done", "w").write("ok\n")
    log(f"[bench] cached to {pre} in {time.time() - t:.1f}s")
    return pre, g


def load_graph(kind, scale, degree=16):
    import gardenia_b200 as gb
    pre, g = ensure_graph(kind, scale, degree)
    if g is None:
        t = time.time()
        g = gb.Graph(pre, "bin", True, False)
        log(f"[bench] loaded {pre}: m={g.m} nnz={g.nnz} in {time.time() - t:.1f}s")
    return pre, g


class ClockSampler:
    """nvidia-smi clocks + throttle reasons.  The sampler is started BEFORE the warm-up steps (nvidia-smi needs a few hundred
    ms to come up, and a multi-GPU timed region lasts 0.1 s) and samples every 20 ms with a timestamp; the reported values
    come from the samples inside the timed region [t0, t1] (`window: "timed"`), or -- if the region was too short to
    catch one -- from the samples since the start of the warm-up (`window: "warmup+timed"`), i.e. the same kernels under load."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.p = device, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    @staticmethod
    def _ts(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, t0=None, t1=None):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        out, _ = self.p.communicate(timeout=5)
        rows, mx = [], None
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                clk = float(f[2]); mx = float(f[3])
            except ValueError:
                continue
            rs = {name for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[6:10])
                  if v.lower().startswith("active")}
            rows.append((self._ts(f[0]), clk, rs))
        timed = [r for r in rows if t0 is not None and r[0] is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        use, window = (timed, "timed") if timed else (rows, "warmup+timed")
        sm = sorted(r[1] for r in use)
        reasons = set().union(*[r[2] for r in use]) if use else set()
        # under load = upper half of the samples (a sample taken between two solves may catch an idle clock)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def run_reference_binary(args_list, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="cores")
    t = time.time()
    r = subprocess.run(args_list, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-500:])
    ms = [float(x) for x in re.findall(r"runtime \[omp_\w+\] = ([0-9.]+) ms", r.stdout)]
    iters = [int(x) for x in re.findall(r"iterations = (\d+)", r.stdout)]
    return ms, iters, time.time() - t


def cpu_reference_pr(prefix, repeat, threads):
    """The reference's own pr_omp_base (oracle/_ref, unmodified sources) on the host cores."""
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    out = os.path.join(CACHE_DIR, "ref_scores.f32")
    if os.path.exists(drv):
        ms, iters, wall = run_reference_binary([drv, "pr", "bin", prefix, "1", out, str(repeat)], threads)
        tot_ms, tot_it = sum(ms), sum(iters)
        return dict(kind="reference", value=tot_it / (tot_ms / 1e3), solves=len(ms), iterations=tot_it, ms=tot_ms, wall_s=wall)
    # the reference did not travel: time the oracle port (OpenMP row loops) instead
    import numpy as np
    import gardenia_b200 as gb
    from oracle import pyoracle as po
    g = gb.Graph(prefix, "bin", True, False)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    t = time.time()
    _, it, _ = po.pr_pull(g.m, g.out_rowptr(), g.out_colidx(), g.out_degrees(), max_iter=3)
    dt = time.time() - t
    return dict(kind="port", value=min(it, 3) / dt, solves=1, iterations=min(it, 3), ms=dt * 1e3, wall_s=dt)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    pre, _ = ensure_graph("g", args.scale)
    repeat = max(1, min(args.steps, 2))
    r = cpu_reference_pr(pre, repeat, threads)
    meta = open(pre + ".meta.txt").read().split()
    m, nnz = int(meta[0]), int(meta[1])
    line = {
        "impl": "reference", "metric": "pagerank_pull_iterations_per_s", "value": r["value"], "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": r["solves"], "warmup": 0, "ms_per_step": r["ms"] / r["solves"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PageRank pull to convergence, Kronecker scale-{args.scale} ef16 (m={m}, nnz={nnz}), "
                               "reference pr_omp_base on the host cores", "l2": "inputs >> L2"},
        "cpu_baseline": {"value": r["value"], "unit": "iterations/s", "cores": threads, "kind": r["kind"],
                         "sample": f"{r['solves']} full solve(s), {r['iterations']} iterations, the reference's own 'runtime [omp_base]' line"},
        "e2e": {"value": r["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("GDN_BENCH_SCALE", "26")))
    ap.add_argument("--spmv-scale", type=int, default=int(os.environ.get("GDN_BENCH_SPMV_SCALE", "24")))
    ap.add_argument("--no-also", action="store_true", help="skip the BFS / SpMV side metrics")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import gardenia_b200 as gb
    from gardenia_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib.gdn_init(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            _lib.check(_lib.lib.gdn_comm_unique_id(buf.ctypes.data))
            uid = torch.from_numpy(buf)
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        uid_host = np.ascontiguousarray(uid.cpu().numpy())       # keep alive across the C call
        _lib.check(_lib.lib.gdn_comm_init(rank, world, uid_host.ctypes.data))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- graph
    # torchrun exports OMP_NUM_THREADS=1; the host side (generator, layout preprocessing) is OpenMP code
    ncpu = os.cpu_count() or 1
    if rank == 0:
        _lib.lib.gdn_set_host_threads(ncpu)                  # the other ranks wait at the barrier
        pre, g = load_graph("g", args.scale)
    barrier()
    if rank != 0:
        pre, g = load_graph("g", args.scale)
    _lib.lib.gdn_set_host_threads(max(1, ncpu // world))
    m, nnz = g.m, g.nnz
    bounds = gb.partition_rows(m, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    t0 = time.time()
    dg = gb.DeviceGraph(g, lo, hi, device=local_rank)
    info = dg.info()
    log(f"[bench] rank {rank}: rows [{lo},{hi}) nnz_local={info['nnz_local']} upload+schedule {time.time() - t0:.2f}s "
        f"blocks={info['n_row_blocks']} heavy_segs={info['n_heavy_segments']} dev_bytes={info['device_bytes'] / 1e9:.2f} GB")
    rows = hi - lo
    scores = torch.empty(rows, dtype=torch.float32, device=dev)
    init = float(np.float32(1.0) / np.float32(m))

    def pr_step():
        scores.fill_(init)                       # src/pr/main.cc:17-18
        return dg.pagerank(scores)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        st = pr_step()
    pinfo = dg.pull_info()
    if pinfo["banded"] == 2:
        kernel_name = (f"pr_seg_kernel x{pinfo['bands']} + pr_sell_pipe + pr_band_finalize (the launches of ONE PageRank iteration over this "
                       "rank's rows; segmented mode: one pass per L2-sized slice of the gathered vector)")
    elif pinfo["banded"]:
        # one iteration of the banded layout (csrc/band.cu) is four launches timed as one unit
        kernel_name = ("pr_band_kernel + pr_sell_pipe + pr_sell_finalize + pr_band_finalize (the launches of ONE PageRank iteration "
                       f"over this rank's rows; {pinfo['band_entries'] / max(info['nnz_local'], 1):.1%} of their column ids are gathered from {pinfo['bands']} "
                       "shared-memory bands)")
    else:
        kernel_name = "pr_sell_pipe (one launch = one PageRank iteration over all rows)"
    barrier()
    w0 = time.time()
    solve_ms = kern_ms = 0.0
    kern_calls = launches = iters = 0
    for _ in range(args.steps):
        st = pr_step()
        solve_ms += st.solve_ms; kern_ms += st.kernel_ms; kern_calls += st.kernel_calls
        launches += st.kernel_launches; iters += st.iterations
    barrier()
    wall_ms = (time.time() - w0) * 1e3
    clocks = sampler.stop(w0, w0 + wall_ms / 1e3) if rank == 0 else None
    solve_ms = allmax(solve_ms)
    kern_ms_max = allmax(kern_ms)
    value = iters / (solve_ms / 1e3)
    peak, peak_src = hbm_peak()
    # algorithmic bytes of ONE gather_kernel launch = one PR iteration over this rank's rows
    # (SURVEY §8(d)): 4*nnz (in_colidx) + 20*rows (offsets, contrib gather once, contrib write, scores r/w) + 4
    alg_bytes = 4 * info["nnz_local"] + 20 * rows + 4
    avg_kernel_ms = kern_ms / max(kern_calls, 1)
    achieved = alg_bytes / (avg_kernel_ms / 1e3) / 1e9
    traffic = None          # ncu capture of the single-GPU iteration only; a rank of a partition moves a different amount
    try:
        if world > 1:
            raise LookupError("no ncu capture for a partitioned run")
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"pr_gather_kron{args.scale}")
    except Exception:
        pass

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    # One step = one PRSolver call on HOST arrays: CSR + degree array + scores cross PCIe, the layout is built, the solve
    # runs, the scores come back and every device buffer is freed, all inside the timed call (the ownership contract of the
    # reference's CUDA solvers, src/pr/base.cu:83-139).  The caller's arrays are page-locked once, like a loader would; the
    # caller's own `scores[i] = 1/m` initialisation (src/pr/main.cc:17-18) sits between the calls, outside the clock, as
    # it sits outside the reference's solver.  One untimed call first (allocator growth, first-touch of the pinned pages).
    e2e_steps = max(1, min(args.steps, 5))
    h_scores = torch.empty(rows, dtype=torch.float32).pin_memory()
    for a in (g.out_rowptr(), g.out_colidx(), g.out_degrees()):
        _lib.lib.gdn_host_pin(a.ctypes.data, a.nbytes)       # page-lock the caller's CSR + degree array once (untimed)
    dg.close()
    barrier()
    e_iters = 0
    h2d = d2h = 0
    e2e_s = 0.0
    e2e_calls = []
    e2e_parts = []           # per call: [upload + layout, solve, download] ms as the library timed them
    for k in range(e2e_steps + 1):
        if world == 1:
            hs = h_scores.numpy()
            hs.fill(init)
            t_call = time.perf_counter()
            st = gb.PRSolver(g, hs, verbose=False)           # upload + layout + solve + download inside the call
            dt = time.perf_counter() - t_call
            h2d, d2h = st.h2d_bytes, st.d2h_bytes
        else:
            h_scores.fill_(init)
            barrier()
            # one solve per upload: the banded layout (1.3 s of preprocessing) would not amortise -- plain SELL layout,
            # as the single-GPU one-shot entry point chooses by itself
            os.environ["GDN_PR_BANDS"] = "0"
            t_call = time.perf_counter()
            dgi = gb.DeviceGraph(g, lo, hi, device=local_rank)
            sc = h_scores.to(dev, non_blocking=True)
            st = dgi.pagerank(sc)
            h_scores.copy_(sc)
            torch.cuda.synchronize()
            dgi.close()
            barrier()
            dt = time.perf_counter() - t_call
            h2d, d2h = 8 * (rows + 1) + 4 * info["nnz_local"] + 4 * rows, 4 * rows
        if k == 0:
            continue                                          # warm-up call
        e_iters += st.iterations
        e2e_s += dt
        e2e_calls.append(round(dt * 1e3, 1))
        e2e_parts.append([round(float(getattr(st, "h2d_ms", 0.0)), 1), round(float(st.solve_ms), 1), round(float(getattr(st, "d2h_ms", 0.0)), 1)])
    e2e_s = allmax(e2e_s)
    e2e = {"value": e_iters / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "ms_per_step": e2e_s * 1e3 / e2e_steps, "ms_per_call": e2e_calls, "h2d_solve_d2h_ms_per_call": e2e_parts,
           "median_ms_per_call": sorted(e2e_calls)[len(e2e_calls) // 2],
           "timed": "wall clock around each PRSolver call on pinned host arrays (1 warm-up call)"}

    also = {}
    if world == 1 and not args.no_also:
        also = side_metrics(args, g, local_rank, peak, pre)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference_pr(pre, 1, os.cpu_count() or 1)
            cpu = {"value": r["value"], "unit": "iterations/s", "cores": os.cpu_count() or 1, "kind": r["kind"],
                   "sample": f"{r['solves']} full solve of the same Kron-{args.scale} graph ({r['iterations']} iterations), "
                             f"reference 'runtime [omp_base]' = {r['ms']:.0f} ms"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "iterations/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": "pagerank_pull_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": solve_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PageRank pull to convergence (damp .85, eps 1e-4), Kronecker scale-{args.scale} ef16 "
                                   f"(m={m}, nnz={nnz} directed entries), reference generator streams",
                       "partition": f"1-D rows x{world}, NCCL allgather of contrib per iteration" if world > 1 else "single GPU",
                       "iterations_per_solve": iters / args.steps, "l2": "inputs (4*nnz B) >> 126 MB L2, no flush needed"},
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_kernel_ms,
                         "launches_timed": int(kern_calls),
                         "kernel_share_of_step": kern_ms_max / solve_ms if solve_ms else None},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if also:
            line["also"] = also
        print(json.dumps(line), flush=True)
    if world > 1:
        _lib.lib.gdn_comm_destroy()
        dist.destroy_process_group()
    return 0


def bfs_algorithmic_bytes(m, steps):
    """SURVEY §8(d) byte model evaluated over the schedule this BFS actually ran (gdn_stats.steps):
    TD  4|F| (queue) + 8|F| (offsets) + 4 E_td (columns) + 4 new (depth) + 4 new (queue)
    BU  3 m/8 (visited, front read; next write) + 8 V_bu (offsets) + 4 E_bu (columns probed) + 4 new (depth)
    conversions Q->B: 4|F| + m/8, B->Q: m/8 + 4|F|.  Bitmap probes per edge count as 0 (L2-resident)."""
    total, prev = 0, None
    for s in steps:
        if s["dir"] == 0:
            if prev == 1:
                total += m // 8 + 4 * s["frontier"]
            total += 12 * s["frontier"] + 4 * s["edges"] + 8 * s["discovered"]
        else:
            if prev != 1:
                total += 4 * s["frontier"] + m // 8
            total += 3 * (m // 8) + 8 * s["scanned"] + 4 * s["edges"] + 4 * s["discovered"]
        prev = s["dir"]
    return total


def side_metrics(args, g, device, peak, pre):
    """BFS GTEPS on the same Kronecker graph and SpMV GFLOP/s on urand (N=1 only)."""
    import numpy as np
    import torch
    import gardenia_b200 as gb
    out = {}
    dev = torch.device("cuda", device)
    m, nnz = g.m, g.nnz
    dg = gb.DeviceGraph(g, device=device)
    depth = torch.empty(m, dtype=torch.int32, device=dev)
    sources = [int(s) for s in g.pick_sources(16)]
    for s in sources[:3]:
        dg.bfs(s, depth)                              # warm-up (first call also builds the hubs-first copy)
    tot_ms = tot_edges = kern_ms = 0.0
    launches = 0
    alg = 0
    per = []
    for s in sources:
        st = dg.bfs(s, depth)
        tot_ms += st.solve_ms; tot_edges += st.edges_reached / 2; kern_ms += st.kernel_ms; launches += st.kernel_launches
        alg += bfs_algorithmic_bytes(m, st.bfs_steps())
        per.append(st.edges_reached / 2 / (st.solve_ms / 1e3) / 1e9)
    achieved = alg / (tot_ms / 1e3) / 1e9
    out["bfs"] = {"metric": "bfs_gteps", "value": tot_edges / (tot_ms / 1e3) / 1e9, "unit": "GTEPS",
                  "workload": f"direction-optimizing BFS, Kronecker scale-{args.scale}, 16 GAP-style sources, undirected edges of the reached component / solve time",
                  "sources": sources, "median_gteps": float(np.median(per)), "ms_per_bfs": tot_ms / len(sources),
                  "bu_sweep_share": kern_ms / tot_ms, "gpu_launches": int(launches),
                  "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                               "algorithmic_bytes_per_bfs": alg / len(sources),
                               "note": "bytes of SURVEY 8(d) over the schedule actually run (hubs-first rows probe fewer in-edges than the oracle order); whole BFS incl. per-level host syncs"}}
    if not args.no_cpu:
        try:
            drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
            ms, iters, _ = run_reference_binary([drv, "bfs", "bin", pre, "1", "0", str(sources[0]), os.path.join(CACHE_DIR, "ref_depth.i32"), "3"], os.cpu_count() or 1)
            ref_ms = min(ms)
            st = dg.bfs(sources[0], depth)
            out["bfs"]["cpu_baseline"] = {"value": st.edges_reached / 2 / (ref_ms / 1e3) / 1e9, "unit": "GTEPS", "cores": os.cpu_count() or 1, "kind": "reference",
                                          "sample": f"bfs_omp_beamer, source {sources[0]}, best of 3, {ref_ms:.1f} ms"}
        except Exception as e:  # noqa: BLE001
            out["bfs"]["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
    dg.close()
    del depth
    # SpMV on urand
    upre, gu = load_graph("u", args.spmv_scale)
    dgu = gb.DeviceGraph(gu, device=device)
    Ax = torch.from_numpy(gb.fill_uniform(13, gu.nnz)).to(dev)
    x = torch.from_numpy(gb.fill_uniform(14, gu.m)).to(dev)
    y = torch.zeros(gu.m, dtype=torch.float32, device=dev)
    for _ in range(3):
        dgu.spmv(Ax, x, y)
    ms = 0.0
    reps = 10
    for _ in range(reps):
        st = dgu.spmv(Ax, x, y)
        ms += st.kernel_ms
    ms /= reps
    alg = 8 * gu.nnz + 16 * gu.m + 4
    out["spmv"] = {"metric": "spmv_gflops", "value": 2.0 * gu.nnz / (ms / 1e3) / 1e9, "unit": "GFLOP/s",
                   "workload": f"fp32 CSR SpMV y+=Ax, uniform-random scale-{args.spmv_scale} (m={gu.m}, nnz={gu.nnz})",
                   "ms": ms, "iterations_per_s": 1e3 / ms,
                   "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg}}
    dgu.close()
    del Ax, x, y
    if not args.no_cpu:
        try:
            drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
            rms, _, _ = run_reference_binary([drv, "spmv", "bin", upre, "1", "0", "13", os.path.join(CACHE_DIR, "ref_y.f32"), "3"], os.cpu_count() or 1)
            ref_ms = min(rms)
            out["spmv"]["cpu_baseline"] = {"value": 2.0 * gu.nnz / (ref_ms / 1e3) / 1e9, "unit": "GFLOP/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                           "sample": f"spmv_omp_base, best of 3, {ref_ms:.1f} ms"}
        except Exception as e:  # noqa: BLE001
            out["spmv"]["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
    return out


if __name__ == "__main__":
    sys.exit(main())
