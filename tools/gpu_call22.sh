#!/bin/bash
# banded shared-memory PageRank layout (csrc/band.cu): parity on small graphs, memcheck, then Kron-26 timing A/B
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or resident or plain_layout or synthetic" > $O/c22_pytest.log 2>&1; tail -15 $O/c22_pytest.log
cat > /tmp/small_band.py <<'PY'
import os, numpy as np, torch
os.environ.update(GDN_PR_BANDS="64", GDN_PR_BAND_SIZE="256", GDN_PR_BAND_CMIN="2", GDN_PR_BAND_DMIN="8")
import gardenia_b200 as gb
g = gb.Graph.generate("g", 13, 16)
dg = gb.DeviceGraph(g)
s = torch.full((g.m,), 1.0 / g.m, dtype=torch.float32, device="cuda")
st = dg.pagerank(s)
print("memcheck run:", st.iterations, dg.pull_info(), float(s.double().sum()))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/small_band.py > $O/c22_memcheck.log 2>&1; grep -E "ERROR SUMMARY|Invalid|memcheck run|at gdn" $O/c22_memcheck.log | head -20
GDN_TRACE=1 timeout 1500 python tools/prof_run.py pr --kind g --scale 26 --reps 2 \
  --sweep "GDN_PR_BANDS=0;GDN_PR_BANDS=64;GDN_PR_BANDS=64,GDN_PR_BAND_PD=8" > $O/c22_pr26.json 2> $O/c22_pr26.err
grep -E "band|error|Error" $O/c22_pr26.err | head -30; cat $O/c22_pr26.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pr_ -c 60 --csv --log-file $O/c22_launches.csv \
  python tools/prof_run.py pr --kind g --scale 26 --reps 1 --sweep "GDN_PR_BANDS=64" > $O/c22_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/c22_launches.csv') if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1e6:9.3f} ms")
PY
