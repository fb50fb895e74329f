#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t9_pytest.log 2>&1; tail -5 gpurun_out/t9_pytest.log
python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep 'GDN_PR_POLICY=1;GDN_PR_POLICY=2;GDN_PR_POLICY=1,GDN_PR_PERSIST=48;GDN_PR_POLICY=2,GDN_PR_PERSIST=48;GDN_PR_POLICY=2,GDN_PR_PERSIST=64,GDN_PR_WARM_MB=64;GDN_PR_POLICY=2,GDN_PR_PERSIST=32,GDN_PR_WARM_MB=48' > gpurun_out/p9_pr26.json 2> gpurun_out/p9_pr26.err
python -c "
import json
d=json.load(open('gpurun_out/p9_pr26.json'))
for r in d['runs']: print(r)
"
python tools/prof_run.py bfs --kind g --scale 26 --reps 16 > gpurun_out/p9_bfs26.json 2> gpurun_out/p9_bfs26.err
python -c "
import json
d=json.load(open('gpurun_out/p9_bfs26.json'))
for r in d['runs']:
    print(r['source'], round(r['solve_ms'],3), r['iterations'], round(r['gteps'],1), round(r['kernel_ms'],3), r['launches'])
"
python tools/prof_run.py spmv --kind u --scale 24 --reps 4 > gpurun_out/p9_spmv24.json 2> gpurun_out/p9_spmv24.err
cat gpurun_out/p9_spmv24.json
