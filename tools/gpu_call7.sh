#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t7_pytest.log 2>&1; tail -3 gpurun_out/t7_pytest.log
(time python bench.py) > gpurun_out/t7_bench.json 2> gpurun_out/t7_bench.err
tail -5 gpurun_out/t7_bench.err
cat gpurun_out/t7_bench.json
