#!/bin/bash
# HEAD (call 35): full GPU suite, warp-cooperative finalize timing, full bench line, ncu launch list of the same bench command
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c35_pytest.log 2>&1; tail -8 $O/c35_pytest.log
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/c35_bench.json 2> $O/c35_bench.err; tail -4 $O/c35_bench.err; cat $O/c35_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c35_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-also > $O/c35_bench_under_ncu.log 2>&1
grep -c gdn:: $O/c35_launches.csv
