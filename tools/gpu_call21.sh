#!/bin/bash
# re-entry check of HEAD (pr_sell_pipe default, spmv_pipe, streaming one-shot build): parity suite, full bench line,
# ncu launch list of the same bench command (per-launch times are cold-cache/serialised: shares only)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c21_pytest.log 2>&1; tail -4 $O/c21_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/c21_bench.json 2> $O/c21_bench.err; tail -5 $O/c21_bench.err; cat $O/c21_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c21_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-also > $O/c21_bench_under_ncu.log 2>&1
grep -c gdn:: $O/c21_launches.csv
