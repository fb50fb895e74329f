#!/bin/bash
# experiments: (1) what the cold tier of the PR gather costs (timing only, wrong results), (2) SpMV column passes
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/prof_run.py pr --kind g --scale 26 --reps 2 --sweep "GDN_PR_SKIP_FROM_MB=100000;GDN_PR_SKIP_FROM_MB=48;GDN_PR_WARM_MB=32,GDN_PR_SKIP_FROM_MB=32;GDN_PR_WARM_MB=16,GDN_PR_SKIP_FROM_MB=16;GDN_PR_WARM_MB=4,GDN_PR_SKIP_FROM_MB=4;GDN_PR_WARM_MB=0,GDN_PR_SKIP_FROM_MB=0;GDN_PR_WARM_MB=48,GDN_PR_SKIP_FROM_MB=80;GDN_PR_WARM_MB=48,GDN_PR_SKIP_FROM_MB=128" > $O/c15_pr_skip.json 2> $O/c15_pr_skip.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c15_pr_skip.json'))
for r in d['runs']: print(r['env'], 'iters', r['iterations'], 'kernel_ms/iter', round(r['kernel_ms']/max(r['kernel_calls'],1),3), 'solve', round(r['solve_ms'],1))
PY
for P in 1 2 3 4; do
  GDN_SPMV_PASSES=$P timeout 300 python tools/prof_run.py spmv --kind u --scale 24 --reps 6 > $O/c15_spmv_p$P.json 2>> $O/c15_spmv.err
  python -c "
import json
d=json.load(open('$O/c15_spmv_p$P.json'))
print('urand24 passes $P', [round(r['kernel_ms'],3) for r in d['runs']])"
done
for P in 1 2 4; do
  GDN_SPMV_PASSES=$P timeout 300 python tools/prof_run.py spmv --kind g --scale 24 --reps 4 > $O/c15_spmv_g_p$P.json 2>> $O/c15_spmv.err
  python -c "
import json
d=json.load(open('$O/c15_spmv_g_p$P.json'))
print('kron24 passes $P', [round(r['kernel_ms'],3) for r in d['runs']])"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/c15_pytest.log 2>&1; tail -3 $O/c15_pytest.log
GDN_SPMV_PASSES=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/c15_pytest_p3.log 2>&1; tail -3 $O/c15_pytest_p3.log
