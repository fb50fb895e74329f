#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference, run in this container.

Requires /root/reference and oracle/_ref (built by `make -C oracle`).  The GPU
box has neither /root/reference nor a need for it: tests read only the files
this script commits.

What is recorded (all produced by the reference's own code):
  * pr_trace_test_pr.json    the golden L1-delta trace of test/reference/graph-pr.mtx.out:13-28
  * <name>.csr.npz           CSR built by the reference readers/generator
  * <name>.ref.npz           bfs_omp_beamer depths (per source), pr_omp_base scores + printed trace,
                             spmv_omp_base y for seeded Ax/x (seed 13)
  * big_hashes.json          sha256 of the same outputs on Kron-16 / urand-16 (too big to commit raw)
Small input graphs (datasets/*.mtx|graph|gr) are copied as data fixtures.
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden")
TMP = tempfile.mkdtemp(prefix="gdn_golden_")


def run(args):
    r = subprocess.run(args, check=True, capture_output=True, text=True)
    return r.stdout


def csr_from_ref(filetype, prefix, symmetrize, reverse):
    out = os.path.join(TMP, "csr")
    run([f"{BIN}/ref_driver", "csr", filetype, prefix, str(symmetrize), str(reverse), out])
    d = dict(out_rowptr=np.fromfile(out + ".out_rowptr.u64", dtype=np.uint64),
             out_colidx=np.fromfile(out + ".out_colidx.i32", dtype=np.int32))
    if os.path.exists(out + ".in_rowptr.u64"):
        d["in_rowptr"] = np.fromfile(out + ".in_rowptr.u64", dtype=np.uint64)
        d["in_colidx"] = np.fromfile(out + ".in_colidx.i32", dtype=np.int32)
        os.remove(out + ".in_rowptr.u64")
    return d


def pr_trace(stdout):
    # lines " %2d    %lf" printed by src/pr/omp_base.cc:35
    return [float(m.group(2)) for m in re.finditer(r"^\s*(\d+)\s+(\d+\.\d+)\s*$", stdout, re.M)]


def ref_results(filetype, prefix, symmetrize, reverse, sources):
    res = {}
    for s in sources:
        f = os.path.join(TMP, "dist.i32")
        o = run([f"{BIN}/ref_driver", "bfs", filetype, prefix, str(symmetrize), str(reverse), str(s), f])
        res[f"bfs_dist_{s}"] = np.fromfile(f, dtype=np.int32)
        res[f"bfs_iters_{s}"] = np.int32(re.search(r"iterations = (\d+)", o).group(1))
    f = os.path.join(TMP, "scores.f32")
    o = run([f"{BIN}/ref_driver", "pr", filetype, prefix, str(symmetrize), f])
    res["pr_scores"] = np.fromfile(f, dtype=np.float32)
    res["pr_trace"] = np.array(pr_trace(o), dtype=np.float64)
    res["pr_iters"] = np.int32(re.search(r"iterations = (\d+)", o).group(1))
    f = os.path.join(TMP, "y.f32")
    run([f"{BIN}/ref_driver", "spmv", filetype, prefix, str(symmetrize), str(reverse), "13", f])
    res["spmv_y"] = np.fromfile(f, dtype=np.float32)
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. the reference's own golden file
    txt = open(f"{REF}/test/reference/graph-pr.mtx.out").read()
    solver_part = txt.split("Verifying")[0]
    json.dump(dict(source="test/reference/graph-pr.mtx.out:13-28", trace=pr_trace(solver_part),
                   iterations=int(re.search(r"iterations = (\d+)", solver_part).group(1))),
              open(f"{OUT}/pr_trace_test_pr.json", "w"), indent=1)
    # 2. bundled fixtures (data files, copied verbatim)
    for f in ["test_pr.mtx", "4.mtx", "4.graph", "4.gr", "4w.mtx", "chesapeake.mtx"]:
        shutil.copy(f"{REF}/datasets/{f}", f"{OUT}/{f}")
    cases = [  # name, filetype, prefix, symmetrize, reverse, bfs sources
        ("test_pr_dir", "mtx", f"{REF}/datasets/test_pr", 0, 1, [0, 1, 3]),
        ("4_sym", "mtx", f"{REF}/datasets/4", 1, 0, [0, 5, 13]),
        ("4_dir", "mtx", f"{REF}/datasets/4", 0, 1, [0, 8]),
        ("chesapeake_sym", "mtx", f"{REF}/datasets/chesapeake", 1, 0, [0, 38]),
    ]
    # 3. synthetic graphs from the reference generator (bin triple)
    for kind, scale, k in [("g", 10, 16), ("u", 10, 16), ("g", 12, 8)]:
        pre = os.path.join(TMP, f"{kind}{scale}k{k}")
        run([f"{BIN}/ref_gen", f"-{kind}", str(scale), "-k", str(k), "-o", pre])
        cases.append((f"{'kron' if kind == 'g' else 'urand'}{scale}k{k}", "bin", pre, 1, 0, [0, 1, 2, 17]))
    for name, ft, pre, sym, rev, sources in cases:
        csr = csr_from_ref(ft, pre, sym, rev)
        np.savez_compressed(f"{OUT}/{name}.csr.npz", **csr)
        deg = np.diff(csr["out_rowptr"])
        sources = [s for s in sources if s < len(deg)]
        res = ref_results(ft, pre, sym, rev, sources)
        res["sources"] = np.array(sources, dtype=np.int32)
        np.savez_compressed(f"{OUT}/{name}.ref.npz", **res)
        print(name, "m", len(deg), "nnz", int(csr["out_rowptr"][-1]), "pr_iters", int(res["pr_iters"]))
    # 4. larger graphs: hashes only
    hashes = {}
    for kind, scale in [("g", 16), ("u", 16)]:
        pre = os.path.join(TMP, f"{kind}{scale}")
        run([f"{BIN}/ref_gen", f"-{kind}", str(scale), "-o", pre])
        csr = csr_from_ref("bin", pre, 1, 0)
        deg = np.diff(csr["out_rowptr"])
        src = int(np.argmax(deg > 0))
        res = ref_results("bin", pre, 1, 0, [src])
        h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        hashes[f"{kind}{scale}"] = dict(
            m=int(len(deg)), nnz=int(csr["out_rowptr"][-1]), colidx_sum=int(csr["out_colidx"].astype(np.int64).sum()),
            rowptr_sha256=h(csr["out_rowptr"]), colidx_sha256=h(csr["out_colidx"]),
            bfs_source=src, bfs_dist_sha256=h(res[f"bfs_dist_{src}"]), bfs_iters=int(res[f"bfs_iters_{src}"]),
            pr_iters=int(res["pr_iters"]), pr_trace=[float(x) for x in res["pr_trace"]],
            pr_scores_sha256=h(res["pr_scores"]), pr_scores_sum=float(res["pr_scores"].astype(np.float64).sum()),
            spmv_y_sha256=h(res["spmv_y"]))
        print(f"{kind}{scale}", hashes[f"{kind}{scale}"]["m"], hashes[f"{kind}{scale}"]["nnz"])
    json.dump(hashes, open(f"{OUT}/big_hashes.json", "w"), indent=1)
    shutil.rmtree(TMP, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
