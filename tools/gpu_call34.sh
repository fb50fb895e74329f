#!/bin/bash
# segmented mode (L2-sized bands, one launch per band) for graphs without a hot set: parity on small graphs, urand-25/26 A/B
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "segmented or banded" > $O/c34_pytest.log 2>&1; tail -12 $O/c34_pytest.log
run() {  # kind scale env...
  k=$1; sc=$2; shift 2
  env "$@" GDN_TRACE=1 timeout 900 python tools/prof_run.py pr --kind $k --scale $sc --reps 3 > $O/c34_tmp.json 2> $O/c34_tmp.err
  grep -E "layout:|band_build: (begin|done|host|counted)" $O/c34_tmp.err | cut -c1-230
  python - "$k$sc $*" <<'PY'
import json, sys
d = json.load(open('gpurun_out/c34_tmp.json'))
for r in d['runs'][-2:]: print(f"  {sys.argv[1]:40s} it {r['iterations']} kernel_ms/iter {r['kernel_ms']/r['kernel_calls']:.3f} solve {r['solve_ms']:.2f} launches {r['launches']} chk {r['checksum']!r}")
PY
}

run u 25 GDN_PR_SEGMENT=-1
run u 26 GDN_PR_SEGMENT=-1
run u 26 GDN_PR_SEGMENT=-1 GDN_PR_SEG_IDS=8000000
run g 26 GDN_PR_SEGMENT=-1
