#!/bin/bash
# CTA-cooperative band finalize; (cmin, dmin, B) sweep of the banded layout (one process per setting: the layout is
# built once per graph); stage trace of the one-shot entry point
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or plain_layout or golden or directed" > $O/c26_pytest.log 2>&1; tail -5 $O/c26_pytest.log
run() {  # env...
  env "$@" GDN_TRACE=1 timeout 600 python tools/prof_run.py pr --kind g --scale 26 --reps 2 > $O/c26_tmp.json 2> $O/c26_tmp.err
  grep "band layout" $O/c26_tmp.err | cut -c1-260
  python - "$*" <<'PY'
import json, sys
d = json.load(open('gpurun_out/c26_tmp.json'))
for r in d['runs'][-1:]: print(f"  {sys.argv[1]:60s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} chk {r['checksum']!r}")
PY
}
run GDN_PR_BANDS=64
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=3 GDN_PR_BAND_DMIN=32
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=6 GDN_PR_BAND_DMIN=64
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=4 GDN_PR_BAND_DMIN=24
run GDN_PR_BANDS=32 GDN_PR_BAND_CMIN=4 GDN_PR_BAND_DMIN=64
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=8 GDN_PR_BAND_DMIN=128
timeout 600 python tools/e2e_trace.py 26 2>&1 | grep -E "rep |\+ +[0-9]{2,}\." | head -40
