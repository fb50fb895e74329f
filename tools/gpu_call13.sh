#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t13_pytest.log 2>&1; tail -4 gpurun_out/t13_pytest.log
python tools/e2e_trace.py 26 2> gpurun_out/p13_e2e_trace.txt; tail -16 gpurun_out/p13_e2e_trace.txt
python bench.py --no-also --no-cpu > gpurun_out/t13_bench.json 2> gpurun_out/t13_bench.err; cat gpurun_out/t13_bench.json
