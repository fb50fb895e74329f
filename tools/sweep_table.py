#!/usr/bin/env python
"""sweep_table.py <sweep.jsonl> -- the rows of tools/sweep.py as a markdown table (profiles/rN_sweep.md)."""
import json
import sys


def f(x, fmt):
    return "—" if x is None else format(x, fmt)


def main():
    rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
    print("| graph | m | nnz | PR it/s (ms/iter, roofline) | PR ref it/s | PR L1 vs ref, iters | BFS GTEPS (ms, roofline) | BFS ref GTEPS | BFS depths | "
          "SpMV GFLOP/s (ms, roofline) | SpMV ref | SpMV max rel |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        p, b, s = r.get("pr", {}), r.get("bfs", {}), r.get("spmv", {})
        print(f"| {r['kind']}-{r['scale']} | {r['m'] / 1e6:.1f} M | {r['nnz'] / 1e9:.2f} G | "
              f"{f(p.get('iters_per_s'), '.1f')} ({f(p.get('ms_per_iter'), '.2f')}, {f(p.get('roofline_frac'), '.3f')}) | {f(p.get('cpu_iters_per_s'), '.2f')} | "
              f"{f(p.get('l1_vs_reference'), '.2e')}, {p.get('iterations')}={p.get('cpu_iterations')} | "
              f"{f(b.get('gteps'), '.0f')} ({f(b.get('ms_per_bfs'), '.2f')}, {f(b.get('roofline_frac'), '.3f')}) | {f(b.get('cpu_gteps'), '.1f')} | "
              f"{'identical' if b.get('depths_identical') else ('DIFFER' if 'depths_identical' in b else '—')} | "
              f"{f(s.get('gflops'), '.0f')} ({f(s.get('ms'), '.2f')}, {f(s.get('roofline_frac'), '.3f')}) | {f(s.get('cpu_gflops'), '.1f')} | {f(s.get('maxrel_vs_reference'), '.1e')} |")


if __name__ == "__main__":
    main()
