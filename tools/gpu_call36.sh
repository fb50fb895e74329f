#!/bin/bash
# two GPUs: partitioned PR (banded, segmented, fixed-point accumulators) + BFS parity at HEAD
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k 2 > gpurun_out/c36_pytest.log 2>&1; grep -E "\[multi\].*PR|passed|failed|Error|error|assert" gpurun_out/c36_pytest.log | head -20
