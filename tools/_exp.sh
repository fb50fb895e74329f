#!/bin/bash
export PYTHONPATH=$PWD
echo "== e2e trace, 10 one-shot PRSolver calls"
timeout 600 python tools/e2e_trace.py 26 10 2>&1 | grep -E "^rep|pr_oneshot|graph_create|pull_prepare" | awk '/^rep/{print; next} {if ($NF+0 > 0 && $(NF-1)+0 > 150) print}' | head -60
echo "== sweep spmv, Kronecker 22-27, exact rows > 8192"
timeout 1200 python tools/sweep.py --scales 22,23,24,25,26,27 --kinds g --only spmv 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['spmv']; print(d['scale'], {k: r[k] for k in ('ms','main_kernel_ms','launches','gflops','roofline_frac','cpu_gflops','maxrel_vs_reference','rows_bit_identical') if k in r})"
