#!/bin/bash
# thread-per-row finalize restored: timing; cmin 4/5/6; ncu --set full of the kernels of one banded iteration
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or plain_layout or golden or directed or resident" > $O/c28_pytest.log 2>&1; tail -3 $O/c28_pytest.log
run() {
  env "$@" GDN_TRACE=1 timeout 600 python tools/prof_run.py pr --kind g --scale 26 --reps 3 > $O/c28_tmp.json 2> $O/c28_tmp.err
  grep "band layout" $O/c28_tmp.err | cut -c1-200
  python - "$*" <<'PY'
import json, sys
d = json.load(open('gpurun_out/c28_tmp.json'))
for r in d['runs'][-2:]: print(f"  {sys.argv[1]:60s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} chk {r['checksum']!r}")
PY
}
run GDN_PR_BANDS=64
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=5
run GDN_PR_BANDS=64 GDN_PR_BAND_CMIN=6
run GDN_PR_BANDS=64 GDN_PR_BAND_FIN=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pr_band_kernel|pr_sell_pipe|pr_band_finalize" -s 6 -c 3 -f -o $O/c28_pr_band \
    python tools/prof_run.py pr --kind g --scale 26 --reps 1 > $O/c28_ncu.log 2>&1
python tools/ncu_summary.py $O/c28_pr_band.ncu-rep > $O/c28_ncu_summary.txt 2>&1; head -120 $O/c28_ncu_summary.txt
