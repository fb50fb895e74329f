#!/bin/bash
export PYTHONPATH=$PWD
run() { echo "== $*"; env "$@" GDN_PR_KTIME=1 timeout 600 python tools/prof_run.py pr --kind g --scale 26 --reps 2 2>&1 | grep -E "per launch" | cut -c1-200 | tail -1; }
for pf in 8 12 20 24; do run GDN_PR_BAND_PF=$pf; done
sp() { echo "== spmv $*"; env "${@:3}" timeout 600 python tools/prof_run.py spmv --kind $1 --scale $2 --reps 4 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print([ (round(r['solve_ms'],3), round(r['kernel_ms'],3)) for r in d['runs']])"; }
sp u 24 GDN_SPMV_PF=0
sp u 24 GDN_SPMV_PF=1
sp g 26 GDN_SPMV_PF=0
sp g 26 GDN_SPMV_PF=1
for xl in 8192 16384 32768 0; do echo "== sweep spmv exact len $xl"; GDN_SPMV_EXACT_LEN=$xl timeout 900 python tools/sweep.py --scales 26 --kinds g --only spmv --keep 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l)['spmv']; print({k: r[k] for k in ('ms','main_kernel_ms','launches','gflops','maxrel_vs_reference','rows_bit_identical') if k in r})"; done
