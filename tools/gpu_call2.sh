#!/bin/bash
# round-1 profiling pass: microbench + launch list + ncu --set full of the dominant kernels
mkdir -p gpurun_out
./tools/dsmem_microbench > gpurun_out/dsmem.txt 2>&1
# generate + cache graphs once
python tools/prof_run.py pr --kind g --scale 26 --reps 2 > gpurun_out/p_pr26.json 2> gpurun_out/p_pr26.err
python tools/prof_run.py bfs --kind g --scale 26 --reps 4 > gpurun_out/p_bfs26.json 2> gpurun_out/p_bfs26.err
python tools/prof_run.py spmv --kind u --scale 24 --reps 3 > gpurun_out/p_spmv24.json 2> gpurun_out/p_spmv24.err
python tools/prof_run.py spmv --kind g --scale 26 --reps 3 > gpurun_out/p_spmv26.json 2> gpurun_out/p_spmv26.err
# launch lists
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pr26.csv python tools/prof_run.py pr --kind g --scale 26 --reps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bfs26.csv python tools/prof_run.py bfs --kind g --scale 26 --reps 2 > /dev/null 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:pr_sell_kernel -s 1 -c 1 -o gpurun_out/prof_pr_sell26 -f python tools/prof_run.py pr --kind g --scale 26 --reps 1 > gpurun_out/ncu_pr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bu_sweep -s 0 -c 2 -o gpurun_out/prof_bu26 -f python tools/prof_run.py bfs --kind g --scale 26 --reps 1 > gpurun_out/ncu_bu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:td_expand -s 2 -c 1 -o gpurun_out/prof_td26 -f python tools/prof_run.py bfs --kind g --scale 26 --reps 1 > gpurun_out/ncu_td.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 1 -c 1 -o gpurun_out/prof_spmv24 -f python tools/prof_run.py spmv --kind u --scale 24 --reps 2 > gpurun_out/ncu_spmv.log 2>&1
ls -la gpurun_out
cat gpurun_out/dsmem.txt
