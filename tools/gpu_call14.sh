#!/bin/bash
# tests + full bench + launch lists + ncu --set full of the three dominant kernels (evidence for profiles/)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c14_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/c14_pytest.log 2>&1; tail -4 $O/c14_pytest.log
timeout 1200 python bench.py > $O/c14_bench.json 2> $O/c14_bench.err; cat $O/c14_bench.json; tail -5 $O/c14_bench.err
# launch list of the bench command (PR only: the headline)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c14_launches_pr26.csv \
    python bench.py --steps 2 --warmup 3 --no-also --no-cpu > $O/c14_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/c14_launches_bfs26.csv \
    python tools/prof_run.py bfs --kind g --scale 26 --reps 6 > $O/c14_ncu_bfs.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pr_sell_kernel -s 3 -c 2 -f -o $O/c14_pr_sell \
    python tools/prof_run.py pr --kind g --scale 26 --reps 1 > $O/c14_ncu_pr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 1 -c 1 -f -o $O/c14_spmv \
    python tools/prof_run.py spmv --kind u --scale 24 --reps 2 > $O/c14_ncu_spmv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:bu_sweep|td_heavy|td_expand" -s 20 -c 12 -f -o $O/c14_bfs \
    python tools/prof_run.py bfs --kind g --scale 26 --reps 3 > $O/c14_ncu_bfs_full.log 2>&1
ls -la $O | tail -20
timeout 120 ./tools/l2_policy_microbench > $O/c14_l2_policy_microbench.txt 2>&1
timeout 120 ./tools/dsmem_microbench > $O/c14_dsmem_microbench.txt 2>&1
