#!/bin/bash
# full GPU suite at HEAD (banded layout default on, isolated rows settled by pr_sell_load, gather-form store),
# Kron-26 timing, and the standalone durations of the co-resident kernel pair (ncu serialises launches)
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c24_pytest.log 2>&1; tail -8 $O/c24_pytest.log
timeout 900 python tools/prof_run.py pr --kind g --scale 26 --reps 3 --sweep "GDN_PR_BANDS=64;GDN_PR_BANDS=96;GDN_PR_BANDS=0" > $O/c24_pr26.json 2> $O/c24_pr26.err
tail -2 $O/c24_pr26.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/c24_pr26.json'))
for r in d['runs']: print(f"{r['env']:50s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} launches {r['launches']} chk {r['checksum']!r}")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pr_ -c 40 --csv --log-file $O/c24_launches_ovl.csv \
  env GDN_PR_OVERLAP=1 python tools/prof_run.py pr --kind g --scale 26 --reps 1 > $O/c24_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/c24_launches_ovl.csv') if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:60], []).append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1e6:9.3f} ms")
PY
