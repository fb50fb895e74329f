#!/usr/bin/env python
"""launch_list.py <launches.csv> -- per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(cold-cache, serialised launches: read the SHARES, not the absolute times)."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", r[ik])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {us:12.1f} {us / n:10.1f} {us / tot:7.1%}")


if __name__ == "__main__":
    main()
