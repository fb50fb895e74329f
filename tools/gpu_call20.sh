#!/bin/bash
# split column upload + pinned scratch (e2e), pipelined SpMV kernel: parity, sweep, bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c20_pytest.log 2>&1; tail -15 $O/c20_pytest.log
SW="GDN_SPMV_PIPE=0;GDN_SPMV_PIPE=5122;GDN_SPMV_PIPE=10241;GDN_SPMV_PIPE=3842;GDN_SPMV_PIPE=2563;GDN_SPMV_PIPE=2562;GDN_SPMV_PIPE=2564"
timeout 600 python tools/prof_run.py spmv --kind u --scale 24 --reps 3 --sweep "$SW" > $O/c20_spmv_u24.json 2> $O/c20_spmv.err
timeout 600 python tools/prof_run.py spmv --kind g --scale 24 --reps 3 --sweep "$SW" > $O/c20_spmv_g24.json 2>> $O/c20_spmv.err
python - <<'PY'
import json
for f in ('u24','g24'):
    d=json.load(open(f'gpurun_out/c20_spmv_{f}.json'))
    for r in d['runs']: print(f, r['env'], 'kernel_ms', round(r['kernel_ms'],3), 'solve', round(r['solve_ms'],3), 'sum', r['checksum'])
PY
tail -3 $O/c20_spmv.err
timeout 900 python tools/e2e_trace.py 26 2> $O/c20_e2e_trace.txt; grep -v "^\[bench\]" $O/c20_e2e_trace.txt | tail -32
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu > $O/c20_bench.json 2> $O/c20_bench.err; tail -5 $O/c20_bench.err; cut -c1-1800 $O/c20_bench.json
