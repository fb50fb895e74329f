#!/usr/bin/env python
"""sass_summary.py [libgdn_b200.so] -- per kernel: registers / shared memory (cuobjdump -res-usage) and the memory-path
instruction mix of its SASS (what kind of loads the hot loops are made of: LDG width and cache operators, LDS, bulk
copies = UBLKCP, mbarrier = SYNCS, atomics, warp match/vote/shuffle).  Written to profiles/ as the evidence that the
kernels are what DESIGN.md says they are.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["pr_sell_pipe", "pr_band_kernel", "pr_band_finalize_fix", "pr_exact_gather(", "pr_exact_qsum", "pr_exact_combine", "pr_seg_kernel",
           "spmv_pipe<unsigned int", "bfs_persist<unsigned int", "peer_sync_kernel", "radix_scatter", "radix_hist", "band_spread"]
KEEP = re.compile(r"^(LDG|LD\.|LDS|LDSM|STG|STS|ST\.|ATOM|ATOMS|ATOMG|RED|UBLKCP|UBLKPF|SYNCS|MATCH|VOTE|SHFL|REDUX|BAR|MEMBAR|FENCE|CCTL|ERRBAR|FADD|FFMA|FMUL|IMAD|F2I|I2F|DADD|DFMA|ELECT|UTMA)")


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gardenia_b200", "lib", "libgdn_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    usage = {}
    lines = res.splitlines()
    for i, ln in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", ln)
        if m and i + 1 < len(lines):
            usage[m.group(1)] = lines[i + 1].strip()
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    cur, hist = None, {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            hist[cur]["(all)"] += 1
            if KEEP.match(op):
                hist[cur][op] += 1
    for mangled, h in hist.items():
        name = demangle(mangled)
        if not any(k in name for k in KERNELS):
            continue
        print(f"== {name}")
        print(f"   {usage.get(mangled, '')}")
        print(f"   instructions: {h['(all)']}")
        groups = collections.defaultdict(list)
        for op, n in sorted(h.items()):
            if op == "(all)":
                continue
            groups[op.split(".")[0]].append(f"{op} x{n}")
        for gname in sorted(groups):
            print(f"   {gname:8s} " + ", ".join(groups[gname]))


if __name__ == "__main__":
    main()
