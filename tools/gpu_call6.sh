#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t6_pytest.log 2>&1; tail -5 gpurun_out/t6_pytest.log
python tools/prof_run.py bfs --kind g --scale 26 --reps 8 > gpurun_out/p6_bfs26.json 2> gpurun_out/p6_bfs26.err
python -c "
import json
for f in ['gpurun_out/p6_bfs26.json']:
    d=json.load(open(f))
    print(f)
    for r in d['runs']: print(r['source'], round(r['solve_ms'],3), r['iterations'], round(r['gteps'],1), round(r['kernel_ms'],3), r['launches'])
"
python tools/prof_run.py spmv --kind u --scale 24 --reps 4 > gpurun_out/p6_spmv24.json 2> gpurun_out/p6_spmv24.err
GDN_SPMV_LEGACY=1 python tools/prof_run.py spmv --kind u --scale 24 --reps 4 > gpurun_out/p6_spmv24_legacy.json 2> gpurun_out/p6_spmv24_legacy.err
python tools/prof_run.py spmv --kind g --scale 26 --reps 3 > gpurun_out/p6_spmv26.json 2> gpurun_out/p6_spmv26.err
cat gpurun_out/p6_spmv24.json gpurun_out/p6_spmv24_legacy.json gpurun_out/p6_spmv26.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bfs26_v3.csv python tools/prof_run.py bfs --kind g --scale 26 --reps 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_tma -s 1 -c 1 -o gpurun_out/prof_spmv24_tma -f python tools/prof_run.py spmv --kind u --scale 24 --reps 2 > gpurun_out/ncu_spmv2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bu_sweep -s 0 -c 3 -o gpurun_out/prof_bu26_v3 -f python tools/prof_run.py bfs --kind g --scale 26 --reps 1 > gpurun_out/ncu_bu3.log 2>&1
