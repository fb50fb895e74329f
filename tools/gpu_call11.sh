#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t11_pytest.log 2>&1; tail -4 gpurun_out/t11_pytest.log
python tools/prof_run.py pr --kind g --scale 26 --reps 2 > gpurun_out/p11_pr26.json 2> gpurun_out/p11_pr26.err
python -c "
import json
d=json.load(open('gpurun_out/p11_pr26.json'))
for r in d['runs']: print(r)
"
python tools/prof_run.py bfs --kind g --scale 26 --reps 16 > gpurun_out/p11_bfs26.json 2> gpurun_out/p11_bfs26.err
python -c "
import json
d=json.load(open('gpurun_out/p11_bfs26.json'))
tot=0
for r in d['runs']:
    tot+=r['solve_ms']
    print(r['source'], round(r['solve_ms'],3), r['iterations'], round(r['gteps'],1), round(r['kernel_ms'],3), r['launches'])
print('mean ms', tot/len(d['runs']))
"
python tools/e2e_trace.py 26 2> gpurun_out/p11_e2e_trace.txt; cat gpurun_out/p11_e2e_trace.txt
