#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t8_pytest.log 2>&1; tail -5 gpurun_out/t8_pytest.log
python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep 'GDN_PR_PUSH=1;GDN_PR_PUSH=0;GDN_PR_PUSH=1,GDN_PR_WARM_MB=32;GDN_PR_PUSH=1,GDN_PR_WARM_MB=24;GDN_PR_PUSH=1,GDN_PR_WARM_MB=16;GDN_PR_PUSH=1,GDN_PR_WARM_MB=64' > gpurun_out/p8_pr26.json 2> gpurun_out/p8_pr26.err
python -c "
import json
d=json.load(open('gpurun_out/p8_pr26.json'))
for r in d['runs']: print(r)
"
python tools/prof_run.py bfs --kind g --scale 26 --reps 16 > gpurun_out/p8_bfs26.json 2> gpurun_out/p8_bfs26.err
python -c "
import json
d=json.load(open('gpurun_out/p8_bfs26.json'))
for r in d['runs']:
    print(r['source'], round(r['solve_ms'],3), r['iterations'], round(r['gteps'],1), round(r['kernel_ms'],3), r['launches'])
    for s in r['steps']: print('     ', s)
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pr26_v3.csv python tools/prof_run.py pr --kind g --scale 26 --reps 1 > /dev/null 2>&1
