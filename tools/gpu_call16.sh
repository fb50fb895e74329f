#!/bin/bash
# pipelined PR kernel: parity + variant sweep on Kron-26
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/c16_pytest.log 2>&1; tail -5 $O/c16_pytest.log
timeout 900 python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep "GDN_PR_PIPE=0;GDN_PR_PIPE=4512;GDN_PR_PIPE=2512;GDN_PR_PIPE=2768;GDN_PR_PIPE=21024;GDN_PR_PIPE=4768;GDN_PR_PIPE=4384;GDN_PR_PIPE=4512,GDN_PR_WARM_MB=32;GDN_PR_PIPE=4512,GDN_PR_WARM_MB=64;GDN_PR_PIPE=4512,GDN_PR_WARM_MB=96" > $O/c16_pr_pipe.json 2> $O/c16_pr_pipe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c16_pr_pipe.json'))
for r in d['runs']: print(r['env'], 'iters', r['iterations'], 'kernel_ms/iter', round(r['kernel_ms']/max(r['kernel_calls'],1),3), 'solve', round(r['solve_ms'],1), 'sum', r['checksum'])
PY
tail -3 $O/c16_pr_pipe.err
timeout 600 python tools/prof_run.py pr --kind u --scale 24 --reps 2 --sweep "GDN_PR_PIPE=0;GDN_PR_PIPE=4512" > $O/c16_pr_pipe_u24.json 2>> $O/c16_pr_pipe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c16_pr_pipe_u24.json'))
for r in d['runs']: print('urand24', r['env'], 'iters', r['iterations'], 'kernel_ms/iter', round(r['kernel_ms']/max(r['kernel_calls'],1),3), 'solve', round(r['solve_ms'],1), 'sum', r['checksum'])
PY
python tools/e2e_trace.py 26 2> $O/c16_e2e_trace.txt; cat $O/c16_e2e_trace.txt | tail -40
