#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t5_pytest.log 2>&1; tail -5 gpurun_out/t5_pytest.log
python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep 'GDN_PR_POLICY=1,GDN_PR_WARM_MB=48;GDN_PR_POLICY=1,GDN_PR_WARM_MB=64;GDN_PR_POLICY=1,GDN_PR_WARM_MB=80' > gpurun_out/p5_pr26.json 2> gpurun_out/p5_pr26.err
python -c "
import json
d=json.load(open('gpurun_out/p5_pr26.json'))
for r in d['runs']: print(r)
"
python tools/prof_run.py bfs --kind g --scale 26 --reps 8 > gpurun_out/p5_bfs26.json 2> gpurun_out/p5_bfs26.err
GDN_BFS_NO_REORDER=1 python tools/prof_run.py bfs --kind g --scale 26 --reps 8 > gpurun_out/p5_bfs26_noreorder.json 2> gpurun_out/p5_bfs26_noreorder.err
python -c "
import json
for f in ['gpurun_out/p5_bfs26.json','gpurun_out/p5_bfs26_noreorder.json']:
    d=json.load(open(f))
    print(f)
    for r in d['runs']: print(r['source'], round(r['solve_ms'],3), r['iterations'], round(r['gteps'],1), round(r['kernel_ms'],3), r['launches'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bfs26_v2.csv python tools/prof_run.py bfs --kind g --scale 26 --reps 2 > /dev/null 2>&1
