#!/bin/bash
# pipelined PR kernel: depth / width / warps sweep on Kron-26, tier probe, ncu of the default
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/c17_pytest.log 2>&1; tail -3 $O/c17_pytest.log
SW=""
for v in 22768 22640 22896 221024 22512 42512 23640 23768 24512 121024 131024 141024 13896 14768; do SW="$SW;GDN_PR_PIPE=$v"; done
timeout 900 python tools/prof_run.py pr --kind g --scale 26 --reps 1 --sweep "${SW:1};GDN_PR_PIPE=22768,GDN_PR_SKIP_FROM_MB=0,GDN_PR_WARM_MB=0;GDN_PR_PIPE=22768,GDN_PR_SKIP_FROM_MB=4,GDN_PR_WARM_MB=4;GDN_PR_PIPE=22768,GDN_PR_SKIP_FROM_MB=48,GDN_PR_WARM_MB=48" > $O/c17_pr_pipe.json 2> $O/c17_pr_pipe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c17_pr_pipe.json'))
for r in d['runs']: print(r['env'], 'iters', r['iterations'], 'kernel_ms/iter', round(r['kernel_ms']/max(r['kernel_calls'],1),3), 'solve', round(r['solve_ms'],1), 'sum', r['checksum'])
PY
tail -3 $O/c17_pr_pipe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pr_sell_pipe -s 3 -c 1 -f -o $O/c17_pr_pipe \
    python tools/prof_run.py pr --kind g --scale 26 --reps 1 > $O/c17_ncu_pr.log 2>&1
