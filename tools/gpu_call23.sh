#!/bin/bash
# banded PageRank: band sums overlapped with the main sums (two streams, co-resident CTAs); memcheck of the small case
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or plain_layout" > $O/c23_pytest.log 2>&1; tail -15 $O/c23_pytest.log
cat > /tmp/small_band.py <<'PY'
import os, numpy as np, torch
os.environ.update(GDN_PR_BANDS="64", GDN_PR_BAND_SIZE="256", GDN_PR_BAND_CMIN="2", GDN_PR_BAND_DMIN="8")
import gardenia_b200 as gb
g = gb.Graph.generate("g", 13, 16)
dg = gb.DeviceGraph(g)
for ovl in ("0", "1"):
    os.environ["GDN_PR_OVERLAP"] = ovl
    s = torch.full((g.m,), 1.0 / g.m, dtype=torch.float32, device="cuda")
    st = dg.pagerank(s)
    print("memcheck run:", ovl, st.iterations, dg.pull_info(), float(s.double().sum()))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/small_band.py > $O/c23_memcheck.log 2>&1; grep -E "ERROR SUMMARY|Invalid|memcheck run|at gdn" $O/c23_memcheck.log | head -20
timeout 1500 python tools/prof_run.py pr --kind g --scale 26 --reps 2 \
  --sweep "GDN_PR_BANDS=64,GDN_PR_OVERLAP=0;GDN_PR_OVERLAP=1;GDN_PR_OVERLAP=1,GDN_PR_CO_HOT=2048;GDN_PR_OVERLAP=1,GDN_PR_CO_HOT=8192,GDN_PR_BAND_PD=8;GDN_PR_OVERLAP=1,GDN_PR_BAND_PD=4,GDN_PR_WARM_MB=32;GDN_PR_OVERLAP=0,GDN_PR_WARM_MB=32;GDN_PR_OVERLAP=0,GDN_PR_WARM_MB=64" > $O/c23_pr26.json 2> $O/c23_pr26.err
tail -3 $O/c23_pr26.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/c23_pr26.json'))
for r in d['runs']: print(f"{r['env']:70s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} chk {r['checksum']!r}")
PY
