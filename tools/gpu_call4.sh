#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "pr" > gpurun_out/t4_pytest.log 2>&1; tail -3 gpurun_out/t4_pytest.log
python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep 'GDN_PR_POLICY=1,GDN_PR_WARM_MB=48;GDN_PR_POLICY=1,GDN_PR_WARM_MB=64;GDN_PR_POLICY=0' > gpurun_out/p4_pr26.json 2> gpurun_out/p4_pr26.err
python -c "
import json
d=json.load(open('gpurun_out/p4_pr26.json'))
for r in d['runs']: print(r)
"
ncu --set full --clock-control none --import-source on -k regex:pr_sell_kernel -s 1 -c 1 -o gpurun_out/prof_pr_sell26_v2 -f python tools/prof_run.py pr --kind g --scale 26 --reps 1 > gpurun_out/ncu_pr2.log 2>&1
