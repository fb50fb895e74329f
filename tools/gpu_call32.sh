#!/bin/bash
# fixed-point band finalize (pr_band_kernel<.,.,true> + pr_band_finalize_fix): parity, memcheck, timing A/B vs the slot finalize
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or plain_layout or golden or directed or resident" > $O/c32_pytest.log 2>&1; tail -3 $O/c32_pytest.log
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
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/small_band.py > $O/c32_memcheck.log 2>&1; grep -E "ERROR SUMMARY|Invalid|memcheck run|at gdn" $O/c32_memcheck.log | head -10
timeout 900 python tools/prof_run.py pr --kind g --scale 26 --reps 3 --sweep "GDN_PR_BAND_FIN=2;GDN_PR_BAND_FIN=0" > $O/c32_pr26.json 2> $O/c32_pr26.err
tail -2 $O/c32_pr26.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/c32_pr26.json'))
for r in d['runs']: print(f"{r['env']:40s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} chk {r['checksum']!r}")
PY
