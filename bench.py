#!/usr/bin/env python
"""bench.py -- the reference's headline benchmarks on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--metric pr|bfs|spmv] [--scale S]

--metric pr (default; BASELINE.json configs[3], the configuration the metric is quoted on):
    PageRank pull to convergence (damp .85, eps 1e-4, max 100 iterations) on the synthetic Kronecker graph
    `-g 26 -k 16` (m = 67.1 M, nnz ~ 2.1 G directed entries) produced by the reference generator's streams; 1-D row
    partition at N > 1, the contrib vector exchanged by peer-mapped stores from the row epilogues.  One "step" = one full
    PRSolver solve.  metric = PageRank iterations per second (whole job).  At N = 1 the line also carries (key "also")
    the two other metrics below on their N = 1 configurations.
--metric bfs (configs[1] / configs[4]): direction-optimizing BFS from 16 GAP-style sources, Kronecker scale 26 at N = 1
    (scale 22 = `--scale 22`), scale 27 row-partitioned at N > 1.  One step = the 16 BFS.  metric = GTEPS.
--metric spmv (configs[2]): fp32 CSR SpMV on uniform-random scale 24.  One step = one SpMV.  metric = GFLOP/s.

Every line carries `parity`: the GPU result compared with the output the REFERENCE's own OpenMP solver (oracle/_ref,
unmodified sources) wrote for the same input on this box -- PR: L1 distance of the scores + equal iteration count; BFS:
bit-exact depths for all 16 sources; SpMV: maximum relative error per row.

Timing: CUDA events on the library's stream around the solve region of every step (graph resident in HBM; the
reference's timed region, src/pr/base.cu:109-128), summed over K steps, max over ranks; barrier + synchronize on both
sides.  Inputs (>= 2 GB of column indices) are far larger than the 126 MB L2, so no explicit flush is needed.
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
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
METRICS = {"pr": ("pagerank_pull_iterations_per_s", "iterations/s"), "bfs": ("bfs_gteps", "GTEPS"), "spmv": ("spmv_gflops", "GFLOP/s")}


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
    try:                                  # edge streams on the host (libstdc++ RNG semantics), CSR built on the GPU (csrc/build.cu)
        g = gb.Graph.generate_gpu(kind, scale, degree)
        how = "CSR built on the GPU: " + ", ".join(f"{k} {v / 1e3:.2f}s" for k, v in g.build_ms.items())
    except Exception as e:  # noqa: BLE001  (no device in this process yet / out of memory: the host builder gives the same arrays)
        g = gb.Graph.generate(kind, scale, degree)
        how = f"host builder ({e})"
    log(f"[bench] generated {kind}{scale}: m={g.m} nnz={g.nnz} in {time.time() - t:.1f}s ({how})")
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
        # under torchrun every rank maps the ONE cached copy (page cache) instead of reading its own 9-17 GB
        g = gb.Graph(pre, "bin:mmap" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "bin", True, False)
        log(f"[bench] loaded {pre}: m={g.m} nnz={g.nnz} in {time.time() - t:.1f}s")
    return pre, g


def graph_meta(pre):
    meta = open(pre + ".meta.txt").read().split()
    return int(meta[0]), int(meta[1])


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


# ---------------------------------------------------------------------------------------------- the reference on the host cores
def run_reference_binary(args_list, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="spread", OMP_PLACES="cores")
    t = time.time()
    r = subprocess.run(args_list, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-500:])
    ms = [float(x) for x in re.findall(r"runtime \[omp_\w+\] = ([0-9.]+) ms", r.stdout)]
    iters = [int(x) for x in re.findall(r"iterations = (\d+)", r.stdout)]
    return ms, iters, time.time() - t


def ref_out(name):
    return os.path.join(CACHE_DIR, name)


def cpu_reference_pr(prefix, scale, repeat, threads):
    """The reference's own pr_omp_base (oracle/_ref, unmodified sources) on the host cores; leaves its scores and
    iteration count in the cache directory for the parity check."""
    out = ref_out(f"ref_scores_g{scale}.f32")
    if os.path.exists(REF_DRIVER):
        ms, iters, wall = run_reference_binary([REF_DRIVER, "pr", "bin", prefix, "1", out + ".tmp", str(repeat)], threads)
        os.replace(out + ".tmp", out)
        json.dump({"iterations": iters[-1], "ms": ms}, open(out + ".json", "w"))
        tot_ms, tot_it = sum(ms), sum(iters)
        return dict(kind="reference", value=tot_it / (tot_ms / 1e3), solves=len(ms), iterations=tot_it, ms=tot_ms, wall_s=wall)
    # the reference did not travel: time the oracle port (OpenMP row loops) instead
    import gardenia_b200 as gb
    from oracle import pyoracle as po
    g = gb.Graph(prefix, "bin", True, False)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    t = time.time()
    _, it, _ = po.pr_pull(g.m, g.out_rowptr(), g.out_colidx(), g.out_degrees(), max_iter=3)
    dt = time.time() - t
    return dict(kind="port", value=min(it, 3) / dt, solves=1, iterations=min(it, 3), ms=dt * 1e3, wall_s=dt)


def cpu_reference_bfs(prefix, scale, sources, threads):
    """bfs_omp_beamer from every source (one graph load); int8 depths of all sources left for the parity check."""
    out = ref_out(f"ref_depth_g{scale}.i8")
    ms, iters, wall = run_reference_binary([REF_DRIVER, "bfs", "bin", prefix, "1", "0", ",".join(str(s) for s in sources) + ",",
                                            out + ".tmp"], threads)
    os.replace(out + ".tmp", out)
    json.dump({"sources": [int(s) for s in sources], "iterations": iters, "ms": ms}, open(out + ".json", "w"))
    return dict(kind="reference", ms=ms, iterations=iters, wall_s=wall)


def cpu_reference_spmv(prefix, scale, threads, repeat=3):
    out = ref_out(f"ref_y_u{scale}.f32")
    ms, _, wall = run_reference_binary([REF_DRIVER, "spmv", "bin", prefix, "1", "0", "13", out + ".tmp", str(repeat)], threads)
    os.replace(out + ".tmp", out)
    return dict(kind="reference", ms=ms, wall_s=wall)


def pick_sources_from_file(pre, n=16):
    import gardenia_b200 as gb
    g = gb.Graph(pre, "bin", True, False)
    return [int(s) for s in g.pick_sources(n)], g


def reference_arm(args):
    """`--impl reference`: the reference's own OpenMP implementation of the path on the host cores, same config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    metric, unit = METRICS[args.metric]
    if args.metric == "pr":
        pre, _ = ensure_graph("g", args.scale)
        r = cpu_reference_pr(pre, args.scale, max(1, min(args.steps, 2)), threads)
        m, nnz = graph_meta(pre)
        value, steps, ms_per_step = r["value"], r["solves"], r["ms"] / r["solves"]
        workload = (f"PageRank pull to convergence, Kronecker scale-{args.scale} ef16 (m={m}, nnz={nnz}), "
                    "reference pr_omp_base on the host cores")
        sample = f"{r['solves']} full solve(s), {r['iterations']} iterations, the reference's own 'runtime [omp_base]' line"
        kind = r["kind"]
    elif args.metric == "bfs":
        pre, _ = ensure_graph("g", args.scale)
        sources, g = pick_sources_from_file(pre)
        deg = g.out_degrees()
        r = cpu_reference_bfs(pre, args.scale, sources, threads)
        import numpy as np
        d8 = np.fromfile(ref_out(f"ref_depth_g{args.scale}.i8"), dtype=np.int8).reshape(len(sources), g.m)
        edges = sum(float(deg[d8[i] >= 0].sum()) / 2 for i in range(len(sources)))
        value, steps, ms_per_step = edges / (sum(r["ms"]) / 1e3) / 1e9, 1, sum(r["ms"])
        workload = (f"direction-optimizing BFS, Kronecker scale-{args.scale} ef16 (m={g.m}, nnz={g.nnz}), 16 GAP-style sources, "
                    "reference bfs_omp_beamer on the host cores")
        sample = f"16 BFS, one run each, the reference's own 'runtime [omp_beamer]' lines ({sum(r['ms']):.0f} ms in total)"
        kind = "reference"
    else:
        pre, _ = ensure_graph("u", args.scale)
        m, nnz = graph_meta(pre)
        r = cpu_reference_spmv(pre, args.scale, threads)
        best = min(r["ms"])
        value, steps, ms_per_step = 2.0 * nnz / (best / 1e3) / 1e9, len(r["ms"]), best
        workload = f"fp32 CSR SpMV y+=Ax, uniform-random scale-{args.scale} (m={m}, nnz={nnz}), reference spmv_omp_base on the host cores"
        sample = f"best of {len(r['ms'])} SpMVs, the reference's own 'runtime [omp_base]' line"
        kind = "reference"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": 0, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.metric != "bfs" else "int32",
        "data": "synthetic", "config": {"workload": workload, "l2": "inputs >> L2"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------- ours
class Ctx:
    """Process / device context of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from gardenia_b200 import _lib
        import numpy as np
        self.args, self.torch, self.dist, self._lib, self.np = args, torch, dist, _lib, np
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        _lib.check(_lib.lib.gdn_init(self.local_rank))
        if self.world > 1:
            # communicator set-up prints "NCCL version ..." on STDOUT (seen in front of the one JSON line at N=8): while the
            # communicators are made, file descriptor 1 points at stderr
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=self.dev)
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = np.zeros(128, dtype=np.uint8)
                _lib.check(_lib.lib.gdn_comm_unique_id(buf.ctypes.data))
                uid = torch.from_numpy(buf)
            uid = uid.to(self.dev)
            dist.broadcast(uid, 0)
            self._uid = np.ascontiguousarray(uid.cpu().numpy())       # keep alive across the C call
            _lib.check(_lib.lib.gdn_comm_init(self.rank, self.world, self._uid.ctypes.data))
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            # a barrier that leaves the GPUs alone (an NCCL barrier is a kernel spinning on every waiting rank's GPU)
            self.cpu_group = dist.new_group(backend="gloo")
        self.ncpu = os.cpu_count() or 1
        self.peak, self.peak_src = hbm_peak()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def allmax(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def load(self, kind, scale):
        """Rank 0 generates / loads first (with all host threads), the others read the cached file afterwards."""
        lib = self._lib.lib
        if self.rank == 0:
            lib.gdn_set_host_threads(self.ncpu)          # torchrun exports OMP_NUM_THREADS=1; the host side is OpenMP code
            pre, g = load_graph(kind, scale)
        self.barrier()
        if self.rank != 0:
            pre, g = load_graph(kind, scale)
        lib.gdn_set_host_threads(max(1, self.ncpu // self.world))
        return pre, g

    def close(self):
        if self.world > 1:
            self._lib.lib.gdn_comm_destroy()
            self.dist.destroy_process_group()


def ensure_ref(ctx, path, make):
    """Rank 0 makes the reference's output file if the reference arm has not left it on this box; everybody waits."""
    err = None
    if ctx.rank == 0 and not os.path.exists(path):
        try:
            make()
        except Exception as e:  # noqa: BLE001
            err = str(e)
            log(f"[bench] reference output {path} unavailable: {e}")
    ctx.barrier()
    return os.path.exists(path), err


def bench_pr(ctx, args):
    import gardenia_b200 as gb
    np, torch, _lib = ctx.np, ctx.torch, ctx._lib
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    pre, g = ctx.load("g", args.scale)
    m, nnz = g.m, g.nnz
    bounds = gb.partition_rows(m, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    t0 = time.time()
    dg = gb.DeviceGraph(g, lo, hi, device=ctx.local_rank)
    info = dg.info()
    log(f"[bench] rank {rank}: rows [{lo},{hi}) nnz_local={info['nnz_local']} upload+schedule {time.time() - t0:.2f}s "
        f"dev_bytes={info['device_bytes'] / 1e9:.2f} GB")
    rows = hi - lo
    scores = torch.empty(rows, dtype=torch.float32, device=dev)
    init = float(np.float32(1.0) / np.float32(m))

    def pr_step():
        scores.fill_(init)                       # src/pr/main.cc:17-18
        return dg.pagerank(scores)

    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    t_first = time.time()
    st = pr_step()                               # first solve builds the SELL array and the banded layout (untimed, reported)
    first_ms = (time.time() - t_first) * 1e3
    for _ in range(args.warmup - 1):
        st = pr_step()
    pinfo, prep = dg.pull_info(), dg.prep_ms()
    if pinfo["banded"] == 2:
        kernel_name = (f"pr_seg_kernel x{pinfo['bands']} + pr_sell_pipe + pr_band_finalize (the launches of ONE PageRank iteration over this "
                       "rank's rows; segmented mode: one pass per L2-sized slice of the gathered vector)")
    elif pinfo["banded"]:
        kernel_name = ("pr_band_kernel + pr_sell_pipe + pr_sell_finalize + pr_band_finalize (the launches of ONE PageRank iteration "
                       f"over this rank's rows; {pinfo['band_entries'] / max(info['nnz_local'], 1):.1%} of their column ids are gathered from {pinfo['bands']} "
                       "shared-memory bands)")
    else:
        kernel_name = "pr_sell_pipe (one launch = one PageRank iteration over all rows)"
    ctx.barrier()
    w0 = time.time()
    solve_ms = kern_ms = 0.0
    kern_calls = launches = iters = 0
    for _ in range(args.steps):
        st = pr_step()
        solve_ms += st.solve_ms; kern_ms += st.kernel_ms; kern_calls += st.kernel_calls
        launches += st.kernel_launches; iters += st.iterations
    ctx.barrier()
    wall_ms = (time.time() - w0) * 1e3
    clocks = sampler.stop(w0, w0 + wall_ms / 1e3) if rank == 0 else None
    solve_ms = ctx.allmax(solve_ms)
    kern_ms_max = ctx.allmax(kern_ms)
    value = iters / (solve_ms / 1e3)
    # algorithmic bytes of ONE PR iteration over this rank's rows:
    # 4*nnz (in_colidx) + 20*rows (offsets, contrib gather once, contrib write, scores r/w) + 4
    alg_bytes = 4 * info["nnz_local"] + 20 * rows + 4
    avg_kernel_ms = kern_ms / max(kern_calls, 1)
    achieved = alg_bytes / (avg_kernel_ms / 1e3) / 1e9
    traffic = None          # ncu capture of the single-GPU iteration only; a rank of a partition moves a different amount
    try:
        if world == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"pr_gather_kron{args.scale}")
    except Exception:
        pass

    # ---------------------------------------------------------------- parity against the reference's own output on this box
    ref_path = ref_out(f"ref_scores_g{args.scale}.f32")
    have_ref, ref_err = ensure_ref(ctx, ref_path, lambda: cpu_reference_pr(pre, args.scale, 1, ctx.ncpu))
    parity = {"pr_l1": None, "pr_iters_equal": None, "reference": "pr_omp_base scores of the same graph (oracle/_ref, this box)"}
    if have_ref:
        ref = np.fromfile(ref_path, dtype=np.float32, count=rows, offset=4 * lo)
        mine = scores.cpu().numpy()
        l1 = ctx.allsum(float(np.abs(mine.astype(np.float64) - ref.astype(np.float64)).sum()))
        ref_it = json.load(open(ref_path + ".json"))["iterations"]
        parity.update(pr_l1=l1, pr_iters_equal=bool(st.iterations == ref_it), pr_iterations=[int(st.iterations), int(ref_it)],
                      ok=bool(l1 <= 1e-6 and st.iterations == ref_it), n_gpus=world)
    else:
        parity["error"] = ref_err

    # ---------------------------------------------------------------- e2e: host buffers through the public API
    # One step = one PRSolver call on HOST arrays: CSR + degree array + scores cross PCIe, the layout is built, the solve
    # runs, the scores come back, all inside the timed call (the ownership contract of the reference's CUDA solvers,
    # src/pr/base.cu:83-139).  The caller's arrays are page-locked once, like a loader would; the caller's own
    # `scores[i] = 1/m` initialisation (src/pr/main.cc:17-18) sits between the calls, outside the clock, as it sits outside
    # the reference's solver.  One untimed call first (allocator growth, first-touch of the pinned pages).
    e2e_steps = max(1, min(args.steps, 5))
    h_scores = torch.empty(rows, dtype=torch.float32).pin_memory()
    for a in (g.out_rowptr(), g.out_colidx(), g.out_degrees()):
        _lib.lib.gdn_host_pin(a.ctypes.data, a.nbytes)       # page-lock the caller's CSR + degree array once (untimed)
    dg.close()
    ctx.cpu_barrier()
    e_iters = 0
    h2d = d2h = 0
    e2e_s = 0.0
    e2e_calls, e2e_parts = [], []
    e2e_l1 = None
    e2e_last_trace = []
    # N > 1: the SAME call, from ONE process: gdn_init_gpus(N) makes the library split every one-shot solve over the N GPUs
    # (a worker thread per GPU, each uploading its rows over its own PCIe link).  Rank 0 makes the calls, the other ranks of
    # the launch have released their GPUs' memory and wait.
    if rank == 0:
        if world > 1:
            _lib.check(_lib.lib.gdn_init_gpus(world))
        full = torch.empty(m, dtype=torch.float32).pin_memory() if world > 1 else h_scores
        for k in range(e2e_steps + 1):
            hs = full.numpy()
            hs.fill(init)
            t_call = time.perf_counter()
            st = gb.PRSolver(g, hs, verbose=False)           # upload + layout + solve + download inside the call
            dt = time.perf_counter() - t_call
            h2d, d2h = st.h2d_bytes, st.d2h_bytes
            if k == 0:
                continue                                      # warm-up call
            e_iters += st.iterations
            e2e_s += dt
            e2e_last_trace = [float(f"{x:.6g}") for x in st.pr_trace()[:12]]
            e2e_calls.append(round(dt * 1e3, 1))
            e2e_parts.append([round(float(st.h2d_ms), 1), round(float(st.solve_ms), 1), round(float(st.d2h_ms), 1)])
        if have_ref:
            ref_full = np.fromfile(ref_path, dtype=np.float32)
            e2e_l1 = float(np.abs(full.numpy().astype(np.float64) - ref_full.astype(np.float64)).sum())
            parity["pr_l1_e2e"] = e2e_l1
            parity["ok"] = bool(parity.get("ok") and e2e_l1 <= 1e-6)
            del ref_full
        if world > 1:
            _lib.check(_lib.lib.gdn_init_gpus(1))
    ctx.cpu_barrier()            # (the other ranks wait on the CPU: their GPUs belong to rank 0's gang meanwhile)
    e2e = {"value": e_iters / e2e_s if rank == 0 else None, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "ms_per_step": e2e_s * 1e3 / e2e_steps, "ms_per_call": e2e_calls, "h2d_solve_d2h_ms_per_call": e2e_parts,
           "median_ms_per_call": sorted(e2e_calls)[len(e2e_calls) // 2] if e2e_calls else None,
           "iterations_per_call": e_iters / max(len(e2e_calls), 1), "l1_delta_trace_last_call": e2e_last_trace,
           "timed": "wall clock around each PRSolver call on pinned host arrays (1 warm-up call)" +
                    (f"; one process, gdn_init_gpus({world})" if world > 1 else "")} if rank == 0 else None
    for a in (g.out_rowptr(), g.out_colidx(), g.out_degrees()):
        _lib.lib.gdn_host_unpin(a.ctypes.data)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference_pr(pre, args.scale, 1, ctx.ncpu)
            cpu = {"value": r["value"], "unit": "iterations/s", "cores": ctx.ncpu, "kind": r["kind"],
                   "sample": f"{r['solves']} full solve of the same Kron-{args.scale} graph ({r['iterations']} iterations), "
                             f"reference 'runtime [omp_base]' = {r['ms']:.0f} ms"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "iterations/s", "cores": ctx.ncpu, "kind": "reference", "sample": f"failed: {e}"}

    line = {
        "metric": METRICS["pr"][0], "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": solve_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PageRank pull to convergence (damp .85, eps 1e-4), Kronecker scale-{args.scale} ef16 "
                               f"(m={m}, nnz={nnz} directed entries), reference generator streams",
                   "partition": (f"1-D rows x{world}; contrib exchanged by peer-mapped stores from the row epilogues + one device-side "
                                 "barrier per iteration" if world > 1 else "single GPU"),
                   "iterations_per_solve": iters / args.steps, "l2": "inputs (4*nnz B) >> 126 MB L2, no flush needed"},
        "preprocessing_ms": {"note": "once per resident graph, outside the timed solves (cf. the reference's untimed segmenting.h pass)",
                             "graph_create_upload": round(prep["create"], 1), "sell_build": round(prep["sell_build"], 1),
                             "band_build": round(prep["band_build"], 1), "first_solve_wall": round(first_ms, 1),
                             "equivalent_solves": round((prep["sell_build"] + prep["band_build"]) / max(solve_ms / args.steps, 1e-9), 1)},
        "wall_ms_per_step": wall_ms / args.steps,
        "clocks": clocks,
        "e2e": e2e,
        "parity": parity,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": ctx.peak, "unit": "GB/s",
                     "frac": achieved / ctx.peak, "frac_of_nominal_8000_gbs": achieved / 8000.0, "traffic": traffic, "peak_source": ctx.peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_kernel_ms,
                     "launches_timed": int(kern_calls),
                     "kernel_share_of_step": kern_ms_max / solve_ms if solve_ms else None},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    del g
    return line


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


def bench_bfs(ctx, args, scale, steps, warmup, side=False):
    """16 GAP-style sources per step; resident graph; depths checked bit-exactly against bfs_omp_beamer for every source."""
    import gardenia_b200 as gb
    np, torch = ctx.np, ctx.torch
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    pre, g = ctx.load("g", scale)
    m, nnz = g.m, g.nnz
    bounds = gb.partition_rows(m, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    dg = gb.DeviceGraph(g, lo, hi, device=ctx.local_rank)
    depth = torch.empty(m, dtype=torch.int32, device=dev)
    parent = torch.empty(m, dtype=torch.int32, device=dev)
    sources = [int(s) for s in g.pick_sources(16)]
    deg = g.out_degrees()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and not side:
        sampler.start()
    for k in range(max(warmup, 1)):
        for s in sources[:4]:
            dg.bfs(s, depth)                              # (the first call also builds the hubs-first copy)
    prep = dg.prep_ms()
    ctx.barrier()
    w0 = time.time()
    tot_ms = tot_edges = kern_ms = 0.0
    launches = alg = 0
    per = []
    for _ in range(steps):
        for s in sources:
            st = dg.bfs(s, depth)
            tot_ms += st.solve_ms; tot_edges += st.edges_reached / 2; kern_ms += st.kernel_ms; launches += st.kernel_launches
            if world == 1:
                alg += bfs_algorithmic_bytes(m, st.bfs_steps())
            per.append(st.edges_reached / 2 / (st.solve_ms / 1e3) / 1e9)
    ctx.barrier()
    wall_ms = (time.time() - w0) * 1e3
    clocks = sampler.stop(w0, w0 + wall_ms / 1e3) if (rank == 0 and not side) else None
    tot_ms = ctx.allmax(tot_ms)
    value = tot_edges / (tot_ms / 1e3) / 1e9

    # ---------------------------------------------------------------- parity: every source against bfs_omp_beamer on this box
    ref_path = ref_out(f"ref_depth_g{scale}.i8")
    have_ref, ref_err = (False, "skipped (--no-cpu)") if (args.no_cpu and not os.path.exists(ref_path)) else \
        ensure_ref(ctx, ref_path, lambda: cpu_reference_bfs(pre, scale, sources, ctx.ncpu))
    parity = {"bfs_depths_equal": None, "reference": "bfs_omp_beamer depths of the same graph and sources (oracle/_ref, this box)"}
    cpu = None
    if have_ref:
        meta = json.load(open(ref_path + ".json"))
        ok_depth = ok_iter = ok_parent = True
        bad = []
        if meta["sources"] == sources:
            ref8 = np.memmap(ref_path, dtype=np.int8, mode="r").reshape(len(sources), m)
            from oracle import pyoracle as po
            for i, s in enumerate(sources):
                st = dg.bfs(s, depth, parent)
                d = depth.cpu().numpy()
                d8 = np.where(d == gb.GDN_INFINITY, -1, d).astype(np.int8)
                same = bool(np.array_equal(d8, ref8[i]) and int(d[d != gb.GDN_INFINITY].max(initial=0)) < 127)
                ok_depth &= same
                ok_iter &= bool(st.iterations == meta["iterations"][i])
                if not same:
                    bad.append(s)
                if rank == 0 and i < 2:                   # parent tree of two sources, checked on the host (Graph500 rules)
                    ok_parent &= bool(po.bfs_check_parents(m, g.out_rowptr(), g.out_colidx(), s, d, parent.cpu().numpy()) == 0)
            ok_depth = bool(ctx.allsum(0.0 if ok_depth else 1.0) == 0.0)
            parity.update(bfs_depths_equal=ok_depth, bfs_iterations_equal=ok_iter, bfs_parent_tree_valid=ok_parent,
                          sources_checked=len(sources), mismatching_sources=bad, ok=bool(ok_depth and ok_iter and ok_parent), n_gpus=world)
            ref_edges = sum(float(deg[ref8[i] >= 0].sum()) / 2 for i in range(len(sources)))
            cpu = {"value": ref_edges / (sum(meta["ms"]) / 1e3) / 1e9, "unit": "GTEPS", "cores": ctx.ncpu, "kind": "reference",
                   "sample": f"bfs_omp_beamer, the same 16 sources, one run each, {sum(meta['ms']):.0f} ms in total"}
        else:
            parity["error"] = "reference depths on this box are for other sources"
    else:
        parity["error"] = ref_err

    # ---------------------------------------------------------------- e2e: BFSSolver on host arrays (N = 1)
    e2e = None
    if world == 1:
        dg.close()
        for a in (g.out_rowptr(), g.out_colidx()):           # page-lock the caller's CSR once (untimed), like a loader would
            ctx._lib.lib.gdn_host_pin(a.ctypes.data, a.nbytes)
        dist_h = np.empty(m, dtype=np.int32)
        calls = []
        edges = 0.0
        for k, s in enumerate(sources[:4]):
            dist_h.fill(gb.MYINFINITY)                    # src/bfs/main.cc:21, the caller's initialisation
            t_call = time.perf_counter()
            st = gb.BFSSolver(g, s, dist_h, verbose=False)
            dt = time.perf_counter() - t_call
            if k == 0:
                continue
            calls.append(dt)
            edges += st.edges_reached / 2
        e2e = {"value": edges / sum(calls) / 1e9, "unit": "GTEPS", "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes),
               "steps": len(calls), "ms_per_call": [round(c * 1e3, 1) for c in calls],
               "timed": "wall clock around each BFSSolver call on pinned host arrays (upload + solve + download; 1 warm-up call)"}
        for a in (g.out_rowptr(), g.out_colidx()):
            ctx._lib.lib.gdn_host_unpin(a.ctypes.data)
    else:
        dg.close()
        e2e = {"value": None, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "the one-shot BFSSolver entry point is single-GPU; the partitioned BFS runs on resident graphs only"}
    nb = len(sources) * steps
    line = {
        "metric": METRICS["bfs"][0], "value": value, "unit": "GTEPS", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": tot_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"direction-optimizing BFS (alpha 15, beta 18), Kronecker scale-{scale} ef16 (m={m}, nnz={nnz}), 16 GAP-style sources "
                               "per step, undirected edges of the reached component / solve time",
                   "partition": (f"1-D rows x{world}, frontier-bitmap allgather per level (NCCL)" if world > 1 else
                                 "single GPU, one cooperative kernel per BFS (device-side controller)"),
                   "l2": "column array >> 126 MB L2"},
        "sources": sources, "median_gteps": float(np.median(per)), "ms_per_bfs": tot_ms / nb,
        "preprocessing_ms": {"bfs_hubs_first_copy": round(prep["bfs_hubs_first"], 1), "graph_create_upload": round(prep["create"], 1)},
        "clocks": clocks, "e2e": e2e, "parity": parity, "gpu_launches": int(launches),
    }
    if world == 1:
        achieved = alg / (tot_ms / 1e3) / 1e9
        bfs_traffic = None       # ncu capture of ONE BFS of this graph (profiles/r2_ncu_bfs.txt), per launch like `achieved`
        try:
            bfs_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"bfs_kron{scale}")
        except Exception:
            pass
        line["roofline"] = {"bound": "hbm", "kernel": "bfs_persist (whole BFS: top-down expand + bottom-up sweeps + controller)",
                            "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak, "traffic": bfs_traffic,
                            "frac_of_nominal_8000_gbs": achieved / 8000.0, "peak_source": ctx.peak_src, "algorithmic_bytes_per_launch": alg / nb, "avg_launch_ms": tot_ms / nb,
                            "bu_sweep_share": kern_ms / tot_ms if tot_ms else None,
                            "note": "bytes of SURVEY 8(d) over the schedule actually run (hubs-first rows probe fewer in-edges than the oracle order)"}
    if cpu and rank == 0:
        line["cpu_baseline"] = cpu
    del g
    return line


def bench_spmv(ctx, args, scale, steps, warmup, side=False):
    import gardenia_b200 as gb
    np, torch = ctx.np, ctx.torch
    dev = ctx.dev
    if ctx.world > 1:
        raise SystemExit("--metric spmv is a single-GPU configuration (BASELINE.json configs[2]); a single SpMV needs no exchange")
    upre, gu = ctx.load("u", scale)
    m, nnz = gu.m, gu.nnz
    dgu = gb.DeviceGraph(gu, device=ctx.local_rank)
    both = gb.fill_uniform(13, nnz + m)                   # the stream oracle/ref_driver.cc draws Ax, then x from
    Ax_h, x_h = both[:nnz], both[nnz:]
    Ax = torch.from_numpy(Ax_h).to(dev)
    x = torch.from_numpy(x_h).to(dev)
    y = torch.zeros(m, dtype=torch.float32, device=dev)
    sampler = ClockSampler(ctx.local_rank)
    if not side:
        sampler.start()
    for _ in range(max(warmup, 3)):
        dgu.spmv(Ax, x, y)
    w0 = time.time()
    ms = tot = 0.0
    for _ in range(steps):
        st = dgu.spmv(Ax, x, y)
        ms += st.kernel_ms; tot += st.solve_ms
    wall_ms = (time.time() - w0) * 1e3
    clocks = sampler.stop(w0, w0 + wall_ms / 1e3) if not side else None
    ms /= steps
    alg = 8 * nnz + 16 * m + 4
    # parity: one SpMV into y = 0 against spmv_omp_base's y for the same seeded inputs
    y.zero_()
    dgu.spmv(Ax, x, y)
    ref_path = ref_out(f"ref_y_u{scale}.f32")
    cpu = None
    if not args.no_cpu:
        try:
            r = cpu_reference_spmv(upre, scale, ctx.ncpu)
            best = min(r["ms"])
            cpu = {"value": 2.0 * nnz / (best / 1e3) / 1e9, "unit": "GFLOP/s", "cores": ctx.ncpu, "kind": "reference",
                   "sample": f"spmv_omp_base, best of {len(r['ms'])}, {best:.1f} ms"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "sample": f"failed: {e}"}
    parity = {"spmv_maxrel": None, "reference": "spmv_omp_base y for the same seeded Ax, x (oracle/_ref, this box)"}
    if os.path.exists(ref_path):
        ref = np.fromfile(ref_path, dtype=np.float32)
        mine = y.cpu().numpy()
        rel = float((np.abs(mine - ref) / np.maximum(np.abs(ref), 1e-30)).max())
        parity.update(spmv_maxrel=rel, ok=bool(rel <= 1e-5), rows_bit_identical=float((mine == ref).mean()))
    dgu.close()
    del Ax, x, y
    # e2e: SpmvSolver on host arrays (page-locked once, untimed, like a loader would)
    for a in (gu.out_rowptr(), gu.out_colidx(), Ax_h):
        ctx._lib.lib.gdn_host_pin(a.ctypes.data, a.nbytes)
    y_h = np.zeros(m, dtype=np.float32)
    calls = []
    for k in range(3):
        y_h.fill(0)
        t_call = time.perf_counter()
        st = gb.SpmvSolver(gu, Ax_h, x_h, y_h, verbose=False)
        dt = time.perf_counter() - t_call
        if k:
            calls.append(dt)
    e2e = {"value": 2.0 * nnz / (sum(calls) / len(calls)) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(st.h2d_bytes),
           "d2h_bytes_per_step": int(st.d2h_bytes), "steps": len(calls), "ms_per_call": [round(c * 1e3, 1) for c in calls],
           "timed": "wall clock around each SpmvSolver call on pinned host arrays (upload + solve + download; 1 warm-up call)"}
    for a in (gu.out_rowptr(), gu.out_colidx(), Ax_h):
        ctx._lib.lib.gdn_host_unpin(a.ctypes.data)
    line = {
        "metric": METRICS["spmv"][0], "value": 2.0 * nnz / (tot / steps / 1e3) / 1e9, "unit": "GFLOP/s", "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": tot / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fp32 CSR SpMV y+=Ax, uniform-random scale-{scale} ef16 (m={m}, nnz={nnz}), seeded U[0,1) values",
                   "l2": "col + Ax (8*nnz B) >> 126 MB L2, no flush needed"},
        "ms": ms, "iterations_per_s": 1e3 / ms, "clocks": clocks, "e2e": e2e, "parity": parity, "gpu_launches": int(steps),
        "roofline": {"bound": "hbm", "kernel": "spmv_pipe", "achieved": alg / (ms / 1e3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                     "frac": alg / (ms / 1e3) / 1e9 / ctx.peak, "frac_of_nominal_8000_gbs": alg / (ms / 1e3) / 1e9 / 8000.0, "traffic": None, "peak_source": ctx.peak_src,
                     "algorithmic_bytes_per_launch": alg, "avg_launch_ms": ms},
    }
    try:
        line["roofline"]["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"spmv_urand{scale}")
    except Exception:
        pass
    if cpu:
        line["cpu_baseline"] = cpu
    del gu
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--metric", default=os.environ.get("GDN_BENCH_METRIC", "pr"), choices=["pr", "bfs", "spmv"])
    ap.add_argument("--scale", type=int, default=None)
    ap.add_argument("--no-also", action="store_true", help="--metric pr at N=1: skip the BFS / SpMV side lines")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.scale is None:
        env = os.environ.get("GDN_BENCH_SCALE")
        args.scale = int(env) if env else {"pr": 26, "bfs": 26 if max(world, args.gpus) == 1 else 27, "spmv": 24}[args.metric]
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)
    ctx = Ctx(args)
    if args.metric == "pr":
        line = bench_pr(ctx, args)
        if ctx.world == 1 and not args.no_also:
            # the two other headline metrics on their own single-GPU configurations, each a full line of its own
            # (`--metric bfs` / `--metric spmv` print them as THE line)
            also = {}
            try:
                also["bfs"] = bench_bfs(ctx, args, args.scale, 1, 1, side=True)
                also["spmv"] = bench_spmv(ctx, args, int(os.environ.get("GDN_BENCH_SPMV_SCALE", "24")), 10, 3, side=True)
            except Exception as e:  # noqa: BLE001
                also["error"] = repr(e)
            line["also"] = also
            ok = [line["parity"].get("ok")] + [also[k]["parity"].get("ok") for k in ("bfs", "spmv") if k in also]
            line["parity"]["all_metrics_ok"] = bool(all(x is True for x in ok))
            for k in ("bfs", "spmv"):
                if k in also:
                    line["parity"].update({kk: vv for kk, vv in also[k]["parity"].items() if kk.startswith(k)})
    elif args.metric == "bfs":
        line = bench_bfs(ctx, args, args.scale, max(1, min(args.steps, 5)), 1)
    else:
        line = bench_spmv(ctx, args, args.scale, max(args.steps, 10), args.warmup)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
