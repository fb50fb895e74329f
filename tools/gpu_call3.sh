#!/bin/bash
mkdir -p gpurun_out
./tools/l2_policy_microbench > gpurun_out/l2policy.txt 2>&1
python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep 'GDN_PR_POLICY=0;GDN_PR_POLICY=1,GDN_PR_WARM_MB=32;GDN_PR_POLICY=1,GDN_PR_WARM_MB=64;GDN_PR_POLICY=1,GDN_PR_WARM_MB=96;GDN_PR_POLICY=2,GDN_PR_WARM_MB=32;GDN_PR_POLICY=2,GDN_PR_WARM_MB=64;GDN_PR_POLICY=2,GDN_PR_WARM_MB=96;GDN_PR_POLICY=1,GDN_PR_WARM_MB=16;GDN_PR_POLICY=1,GDN_PR_WARM_MB=48' > gpurun_out/p3_pr26.json 2> gpurun_out/p3_pr26.err
cat gpurun_out/l2policy.txt
python -c "
import json
d=json.load(open('gpurun_out/p3_pr26.json'))
for r in d['runs']: print(r)
"
